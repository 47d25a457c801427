#!/bin/bash
# products shape: A/B of the forward occupancy, ncu of the two gather kernels
mkdir -p gpurun_out
for lib in "" variants/libbotgat_mb6.so; do
  BOTGAT_LIB=${lib:+$PWD/$lib} python bench.py --shape products --no-cpu-baseline --no-skew --no-e2e --no-parity --steps 5 > gpurun_out/r02_o_products.json 2> gpurun_out/r02_o_products.err
  python - "$lib" <<'PY'
import json, sys
for l in open("gpurun_out/r02_o_products.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print(sys.argv[1] or "default", round(d["ms_per_step"], 3), {k: v["avg_ms"] for k, v in d["kernels"].items()})
PY
done
ncu --set full --clock-control none --import-source on -k regex:'gat_fwd|gat_bwd_src' --launch-skip 2 --launch-count 2 -f -o gpurun_out/r02_o_products \
  python bench.py --shape products --steps 1 --warmup 1 --no-cpu-baseline --no-skew --no-e2e --no-parity > gpurun_out/r02_o_ncu.log 2>&1
ncu -i gpurun_out/r02_o_products.ncu-rep --page raw --csv > gpurun_out/r02_o_products_raw.csv
python tools/ncu_raw.py gpurun_out/r02_o_products_raw.csv > gpurun_out/r02_o_products_summary.txt
ncu -i gpurun_out/r02_o_products.ncu-rep --page source --csv > gpurun_out/r02_o_products_source.csv 2>/dev/null
python tools/ncu_src.py gpurun_out/r02_o_products_source.csv 25 > gpurun_out/r02_o_products_stalls.txt 2>&1
rm -f gpurun_out/r02_o_products.ncu-rep gpurun_out/r02_o_products_source.csv
cat gpurun_out/r02_o_products_summary.txt; head -60 gpurun_out/r02_o_products_stalls.txt
