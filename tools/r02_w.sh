#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_w_tests.log 2>&1; tail -4 gpurun_out/r02_w_tests.log
for s in proteins products; do
  python bench.py --shape $s --no-cpu-baseline --no-skew --no-e2e --steps 10 > gpurun_out/r02_w_$s.json 2> gpurun_out/r02_w_$s.err
  python - "$s" <<'PY'
import json, sys
for l in open(f"gpurun_out/r02_w_{sys.argv[1]}.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print(sys.argv[1], round(d["ms_per_step"], 3), {k: v["avg_ms"] for k, v in d["kernels"].items()}, d.get("parity", {}).get("parity_max_rel"))
PY
done
