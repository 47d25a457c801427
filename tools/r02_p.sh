#!/bin/bash
mkdir -p gpurun_out
for thr in 0 ; do
  BOTGAT_LOWDEG_BWD=$thr python bench.py --shape products --no-cpu-baseline --no-skew --no-e2e --no-parity --steps 5 > gpurun_out/r02_p_products.json 2> gpurun_out/r02_p_products.err
  python - "$thr" <<'PY'
import json, sys
for l in open("gpurun_out/r02_p_products.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("products lowdeg_bwd thr", sys.argv[1], round(d["ms_per_step"], 3), {k: v["avg_ms"] for k, v in d["kernels"].items()})
PY
done
for w in 4 8; do BOTGAT_LOWDEG_BWD=0 python tools/rank_bench.py --world $w; done
for w in 8; do BOTGAT_LOWDEG=64 BOTGAT_LOWDEG_BWD=96 python tools/rank_bench.py --world $w; done
