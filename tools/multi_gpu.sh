#!/bin/bash
# N GPUs of one box: the 2-GPU exchange test in both kernel families, then the partitioned bench lines
n=${1:-2}; shapes=${2:-"products proteins"}
mkdir -p gpurun_out
if [ "$n" = "2" ]; then
  python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/r02_n${n}_multi_tests.log 2>&1; tail -2 gpurun_out/r02_n${n}_multi_tests.log
  BOTGAT_ROWWISE=1 python -m pytest tests/test_gpu_multi.py -q -x > gpurun_out/r02_n${n}_multi_tests_rowwise.log 2>&1; tail -2 gpurun_out/r02_n${n}_multi_tests_rowwise.log
fi
for s in $shapes; do
  extra="--no-e2e"; [ "$s" = "proteins" ] && extra=""
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --shape $s --no-cpu-baseline --no-skew $extra \
    > gpurun_out/r02_zz_n${n}_$s.json 2> gpurun_out/r02_zz_n${n}_$s.err
  python - $n $s <<'PY'
import json, sys
n, s = sys.argv[1], sys.argv[2]
for l in open(f"gpurun_out/r02_zz_n{n}_{s}.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print(s, "N", d["n_gpus"], round(d["ms_per_step"], 3), f'{d["value"]/1e9:.3f}G', {k: (v["launches"], v["avg_ms"]) for k, v in d["kernels"].items()}, (d.get("graph") or {}).get("halo_exchange"), "e2e", (d.get("e2e") or {}).get("ms_per_step"))
PY
done
