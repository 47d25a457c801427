#!/bin/bash
# Round-2 closing evidence on one GPU: proteins profile (ncu full set + launch list + full bench line, both arms),
# every other BASELINE shape at its own flags, ncu of the products-shape kernels.
mkdir -p gpurun_out
bash tools/profile_round.sh r02_z proteins > gpurun_out/r02_z_profile.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference_r02_z.json 2> gpurun_out/bench_reference_r02_z.err
for s in reddit arxiv products; do
  timeout 600 python bench.py --shape $s --no-cpu-baseline --no-e2e > gpurun_out/r02_z_$s.json 2> gpurun_out/r02_z_$s.err
done
python tools/quick_bench.py --shape cora --iters 20 --warmup 5 > gpurun_out/r02_z_cora.txt 2>&1
python tools/quick_bench.py --shape arxiv_last --iters 20 --warmup 5 > gpurun_out/r02_z_arxiv_last.txt 2>&1
ncu --set full --clock-control none -k regex:'rowwise|rowbulk|gat_bwd_node|k_edge' --launch-skip 7 --launch-count 7 -f -o gpurun_out/r02_z_products \
  python bench.py --shape products --steps 1 --warmup 1 --no-cpu-baseline --no-skew --no-e2e --no-parity > gpurun_out/r02_z_products_ncu.log 2>&1
ncu -i gpurun_out/r02_z_products.ncu-rep --page raw --csv > gpurun_out/r02_z_products_raw.csv
python tools/ncu_raw.py gpurun_out/r02_z_products_raw.csv > gpurun_out/r02_z_products_ncu_summary.txt
python tools/dram_traffic.py gpurun_out/r02_z_products_raw.csv products 1 "ncu --set full of bench.py --shape products --steps 1 --warmup 1 (profiles/r02_z_products_ncu_summary.txt)" 1 > gpurun_out/dram_traffic_r02_z_products.json 2> gpurun_out/dram_traffic_r02_z_products.err
cp profiles/dram_traffic.json gpurun_out/dram_traffic_merged_r02_z.json
rm -f gpurun_out/*.ncu-rep gpurun_out/r02_z_products_raw.csv gpurun_out/prof_r02_z_raw.csv
python - <<'PY'
import json
for s in ("bench_r02_z", "r02_z_reddit", "r02_z_arxiv", "r02_z_products"):
    try:
        for l in open(f"gpurun_out/{s}.json"):
            if l.startswith("{"):
                d = json.loads(l)
                print(s, round(d["ms_per_step"], 3), f'{d["value"]/1e9:.3f}G', {k: v["avg_ms"] for k, v in d["kernels"].items()},
                      "parity", d.get("parity", {}).get("parity_max_rel"), "e2e", (d.get("e2e") or {}).get("ms_per_step"), "skew", (d.get("skew_variant") or {}).get("ms_per_step"))
    except Exception as e:
        print(s, "failed", e)
PY
tail -3 gpurun_out/r02_z_cora.txt gpurun_out/r02_z_arxiv_last.txt
