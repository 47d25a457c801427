#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'gat_bwd_src' --launch-skip 3 --launch-count 1 -f -o gpurun_out/r02_y_rank8 \
  python tools/rank_bench.py --world 8 --iters 2 > gpurun_out/r02_y_ncu.log 2>&1
ncu -i gpurun_out/r02_y_rank8.ncu-rep --page raw --csv > gpurun_out/r02_y_raw.csv
python tools/ncu_raw.py gpurun_out/r02_y_raw.csv > gpurun_out/r02_y_rank8_src_summary.txt
ncu -i gpurun_out/r02_y_rank8.ncu-rep --page source --csv > gpurun_out/r02_y_source.csv 2>/dev/null
python tools/ncu_src.py gpurun_out/r02_y_source.csv 40 > gpurun_out/r02_y_rank8_src_stalls.txt 2>&1
rm -f gpurun_out/r02_y_rank8.ncu-rep gpurun_out/r02_y_source.csv gpurun_out/r02_y_raw.csv
cat gpurun_out/r02_y_rank8_src_summary.txt; head -55 gpurun_out/r02_y_rank8_src_stalls.txt
