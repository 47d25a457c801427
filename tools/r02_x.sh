#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_x_tests.log 2>&1; tail -3 gpurun_out/r02_x_tests.log
bash tools/r02_v.sh variants/libbotgat_noep.so
