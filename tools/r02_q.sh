#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -k "rowwise or tail" > gpurun_out/r02_q_tests.log 2>&1; tail -15 gpurun_out/r02_q_tests.log
for rw in 0 1; do
  BOTGAT_ROWWISE=$rw python bench.py --shape products --no-cpu-baseline --no-skew --no-e2e --steps 5 > gpurun_out/r02_q_products_rw$rw.json 2> gpurun_out/r02_q_products_rw$rw.err
  python - "$rw" <<'PY'
import json, sys
for l in open(f"gpurun_out/r02_q_products_rw{sys.argv[1]}.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("products rowwise", sys.argv[1], round(d["ms_per_step"], 3), {k: v["avg_ms"] for k, v in d["kernels"].items()}, d.get("parity", {}).get("parity_max_rel"))
PY
done
