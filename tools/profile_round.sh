#!/bin/bash
# Round profile: bench JSON (both arms), ncu launch list of the same command, one ncu full capture of our kernels,
# DRAM bytes per launch merged into profiles/dram_traffic.json.
# usage: bash tools/profile_round.sh <tag> [shape]     (run under gpurun; outputs in gpurun_out/)
set -x
tag=${1:-x}; shape=${2:-proteins}
mkdir -p gpurun_out
# one whole step (11 kernels at the proteins shape): the second of warm-up / timed / instrumented
ncu --set full --clock-control none -k regex:'gat_|k_edge_|k_drop_|k_fwd_|k_bwd_' --launch-skip ${SKIP:-11} --launch-count ${COUNT:-11} -f -o gpurun_out/prof_${tag} \
  python bench.py --shape $shape --steps 1 --warmup 1 --no-cpu-baseline --no-skew --no-e2e --no-parity > gpurun_out/ncu_full_${tag}.log 2>&1
ncu -i gpurun_out/prof_${tag}.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv
python tools/ncu_raw.py gpurun_out/prof_${tag}_raw.csv > gpurun_out/ncu_full_summary_${tag}.txt
rm -f gpurun_out/prof_${tag}.ncu-rep   # gpurun_out/ travels back only below 64 MiB
python tools/dram_traffic.py gpurun_out/prof_${tag}_raw.csv $shape 1 "ncu --set full of bench.py --shape $shape --steps 1 --warmup 1 (profiles/${tag}_ncu_full_summary.txt)" 1 > gpurun_out/dram_traffic_${tag}.json
cp profiles/dram_traffic.json gpurun_out/dram_traffic_merged_${tag}.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv \
  python bench.py --shape $shape --steps 2 --warmup 1 --no-cpu-baseline --no-skew --no-e2e --no-parity > gpurun_out/ncu_launch_${tag}.log 2>&1
python bench.py --shape $shape > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
tail -3 gpurun_out/bench_${tag}.err
