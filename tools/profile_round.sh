#!/bin/bash
# Round profile: bench JSON, ncu launch list of the same command, one ncu full capture of our kernels.
# usage: bash tools/profile_round.sh <tag>     (run under gpurun; outputs in gpurun_out/)
set -x
tag=${1:-x}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_${tag}.json 2>> gpurun_out/bench_${tag}.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-skew > gpurun_out/ncu_launch_${tag}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'gat_|k_edge_|k_drop_' -c 22 -f -o gpurun_out/prof_${tag} \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-skew > gpurun_out/ncu_full_${tag}.log 2>&1
ncu -i gpurun_out/prof_${tag}.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv
python tools/ncu_raw.py gpurun_out/prof_${tag}_raw.csv > gpurun_out/ncu_full_summary_${tag}.txt
tail -3 gpurun_out/bench_${tag}.err
