#!/bin/bash
# Round-2 src-pass comparison: parity of the variants, bench with each, ncu full capture of one of them.
# usage (under gpurun): bash tools/r02_tma.sh <tag> "<variants to bench>" <variant to profile>
tag=${1:-h}; variants=${2:-"1 0"}; prof=${3:-1}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -k "tma or parity or families or philox" > gpurun_out/r02_${tag}_tests.log 2>&1
tail -3 gpurun_out/r02_${tag}_tests.log
for v in $variants; do
  BOTGAT_BWD_TMA=$v python bench.py --no-cpu-baseline --no-skew --no-e2e > gpurun_out/r02_${tag}_bench_tma$v.json 2> gpurun_out/r02_${tag}_bench_tma$v.err
  python - <<PY
import json
for l in open("gpurun_out/r02_${tag}_bench_tma$v.json"):
    if l.startswith("{"):
        d = json.loads(l); print("tma=$v", d["ms_per_step"], {k: v["avg_ms"] for k, v in d["kernels"].items()}, d.get("parity", {}).get("parity_max_rel"))
PY
done
BOTGAT_BWD_TMA=$prof ncu --set full --clock-control none --import-source on -k regex:'gat_bwd_src' -c 1 -f -o gpurun_out/r02_${tag}_src_tma \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-skew --no-e2e --no-parity > gpurun_out/r02_${tag}_ncu.log 2>&1
ncu -i gpurun_out/r02_${tag}_src_tma.ncu-rep --page raw --csv > gpurun_out/r02_${tag}_src_tma_raw.csv
python tools/ncu_raw.py gpurun_out/r02_${tag}_src_tma_raw.csv > gpurun_out/r02_${tag}_src_tma_summary.txt
ncu -i gpurun_out/r02_${tag}_src_tma.ncu-rep --page source --csv > gpurun_out/r02_${tag}_src_tma_source.csv
python tools/ncu_src.py gpurun_out/r02_${tag}_src_tma_source.csv 25 > gpurun_out/r02_${tag}_src_tma_stalls.txt
cat gpurun_out/r02_${tag}_src_tma_summary.txt gpurun_out/r02_${tag}_src_tma_stalls.txt
