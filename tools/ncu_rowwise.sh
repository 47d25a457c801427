#!/bin/bash
# ncu of the two rowwise kernels at the products shape
mkdir -p gpurun_out
tag=${1:-r}
BOTGAT_ROWWISE=1 ncu --set full --clock-control none --import-source on -k regex:'rowwise' --launch-skip 2 --launch-count 2 -f -o gpurun_out/r02_${tag}_rw \
  python bench.py --shape products --steps 1 --warmup 1 --no-cpu-baseline --no-skew --no-e2e --no-parity > gpurun_out/r02_${tag}_ncu.log 2>&1
ncu -i gpurun_out/r02_${tag}_rw.ncu-rep --page raw --csv > gpurun_out/r02_${tag}_rw_raw.csv
python tools/ncu_raw.py gpurun_out/r02_${tag}_rw_raw.csv > gpurun_out/r02_${tag}_rw_summary.txt
ncu -i gpurun_out/r02_${tag}_rw.ncu-rep --page source --csv -k regex:'gat_fwd' > gpurun_out/r02_${tag}_fwd_source.csv 2>/dev/null
ncu -i gpurun_out/r02_${tag}_rw.ncu-rep --page source --csv -k regex:'gat_bwd' > gpurun_out/r02_${tag}_bwd_source.csv 2>/dev/null
python tools/ncu_src.py gpurun_out/r02_${tag}_fwd_source.csv 30 > gpurun_out/r02_${tag}_fwd_stalls.txt 2>&1
python tools/ncu_src.py gpurun_out/r02_${tag}_bwd_source.csv 30 > gpurun_out/r02_${tag}_bwd_stalls.txt 2>&1
rm -f gpurun_out/r02_${tag}_rw.ncu-rep gpurun_out/*_source.csv
cat gpurun_out/r02_${tag}_rw_summary.txt; head -45 gpurun_out/r02_${tag}_fwd_stalls.txt; head -45 gpurun_out/r02_${tag}_bwd_stalls.txt
