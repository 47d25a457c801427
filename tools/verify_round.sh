#!/bin/bash
# what the driver runs at round end, on one GPU: smoke(), the GPU test suite, the default bench line
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/verify_smoke.log 2>&1; tail -2 gpurun_out/verify_smoke.log
python -m pytest tests -m gpu -x -q > gpurun_out/verify_tests.log 2>&1; tail -2 gpurun_out/verify_tests.log
python bench.py > gpurun_out/verify_bench.json 2> gpurun_out/verify_bench.err; tail -c 600 gpurun_out/verify_bench.json
