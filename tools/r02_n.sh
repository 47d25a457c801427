#!/bin/bash
# per-rank kernel workloads on one GPU + every BASELINE shape at its own flags (N = 1)
mkdir -p gpurun_out
for w in 2 4 8; do python tools/rank_bench.py --world $w > gpurun_out/r02_n_rank_w$w.json 2> gpurun_out/r02_n_rank_w$w.err; cat gpurun_out/r02_n_rank_w$w.json; done
for s in reddit products arxiv cora; do
  timeout 600 python bench.py --shape $s --no-cpu-baseline --no-skew --no-e2e > gpurun_out/r02_n_$s.json 2> gpurun_out/r02_n_$s.err
  python - $s <<'PY'
import json, sys
s = sys.argv[1]
for l in open(f"gpurun_out/r02_n_{s}.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print(s, round(d["ms_per_step"], 3), f'{d["value"]/1e9:.3f} G/s', {k: v["avg_ms"] for k, v in d["kernels"].items()}, d.get("parity", {}).get("parity_max_rel"))
PY
done
