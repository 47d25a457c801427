#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02_m_tests.log 2>&1; tail -4 gpurun_out/r02_m_tests.log
python bench.py --no-cpu-baseline > gpurun_out/r02_m_bench.json 2> gpurun_out/r02_m_bench.err; tail -2 gpurun_out/r02_m_bench.err
python bench.py --shape arxiv --no-cpu-baseline --no-skew --no-e2e > gpurun_out/r02_m_arxiv.json 2> gpurun_out/r02_m_arxiv.err
python - <<'PY'
import json
for f in ("r02_m_bench", "r02_m_arxiv"):
    for l in open(f"gpurun_out/{f}.json"):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, round(d["ms_per_step"], 3), {k: v["avg_ms"] for k, v in d["kernels"].items()})
            if d.get("e2e"):
                print("  e2e", d["e2e"]["ms_per_step"], "model", d["e2e"].get("model"))
            if d.get("skew_variant"):
                print("  skew", d["skew_variant"])
PY
