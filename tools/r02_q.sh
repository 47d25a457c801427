#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -x -q -k "rowwise or tail" > gpurun_out/r02_q_tests.log 2>&1; tail -15 gpurun_out/r02_q_tests.log
for bulk in 1 0; do
  BOTGAT_RW_BULK=$bulk python bench.py --shape products --no-cpu-baseline --no-skew --no-e2e --steps 5 > gpurun_out/r02_q_products_bulk$bulk.json 2> gpurun_out/r02_q_products_bulk$bulk.err
  python - "$bulk" <<'PY'
import json, sys
for l in open(f"gpurun_out/r02_q_products_bulk{sys.argv[1]}.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("products bulk", sys.argv[1], round(d["ms_per_step"], 3), {k: v["avg_ms"] for k, v in d["kernels"].items()}, d.get("parity", {}).get("parity_max_rel"))
PY
done
