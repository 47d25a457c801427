#!/bin/bash
# what the driver runs at round end, on one GPU: smoke(), the GPU test suite, the default bench line (+ the products line)
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/verify_smoke.log 2>&1; tail -2 gpurun_out/verify_smoke.log
python -m pytest tests -m gpu -x -q > gpurun_out/verify_tests.log 2>&1; tail -2 gpurun_out/verify_tests.log
python bench.py > gpurun_out/verify_bench.json 2> gpurun_out/verify_bench.err
python bench.py --shape products --no-cpu-baseline --no-e2e --steps 5 > gpurun_out/verify_bench_products.json 2> gpurun_out/verify_bench_products.err
python - <<'PY'
import json
for f in ("verify_bench", "verify_bench_products"):
    for l in open(f"gpurun_out/{f}.json"):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, round(d["ms_per_step"], 3), f'{d["value"]/1e9:.3f}G', {k: v["avg_ms"] for k, v in d["kernels"].items()}, "parity", d["parity"]["parity_max_rel"], "e2e", (d.get("e2e") or {}).get("ms_per_step"))
PY
