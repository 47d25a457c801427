"""Developer tool: DRAM bytes per launch of every kernel of one bench step, from `ncu -i X.ncu-rep --page raw --csv` of
`bench.py --steps 1 --warmup 1` -> merged into profiles/dram_traffic.json under [shape][n_gpus] (what bench.py's
`roofline.traffic` / `whole_step.dram` read).  usage: python tools/dram_traffic.py raw.csv <shape> <n_gpus> <source tag>"""
import csv
import json
import os
import sys

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
# bench.py kernel-timer name <- kernel-name substrings
GROUPS = {"gat_fwd": ["gat_fwd_kernel", "gat_fwd_lowdeg", "gat_fwd_rowwise", "k_fwd_combine"], "gat_bwd_node": ["gat_bwd_node"],
          "gat_bwd_src": ["gat_bwd_src", "k_bwd_combine"], "edge_stage": ["k_edge_stage"],
          "gat_bwd_edge": ["k_edge_unstage", "k_edge_reduce_dst"], "edge_drop_draw": ["k_drop_"]}

path, shape, world, tag = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
rows = list(csv.reader(open(path)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}


def val(r, name):
    return float(r[idx[name]].replace(",", "")) * UNIT.get(units[idx[name]], 1)


per_kernel = {}
for r in rows[2:]:
    name = r[idx["Kernel Name"]]
    b = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
    per_kernel.setdefault(name, []).append(b)
# the capture holds warm-up + timed step(s): per-launch average, per-step total = sum over kernels of (avg x launches per step)
steps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
out = {}
total = 0.0
for g, subs in GROUPS.items():
    launches = [(n, v) for n, v in per_kernel.items() if any(s in n for s in subs)]
    if not launches:
        continue
    nl = sum(len(v) for _, v in launches)
    tot = sum(sum(v) for _, v in launches)
    out[g] = int(tot / nl * (1 if g != "gat_bwd_edge" else len(launches)))  # edge phase = unstage + reduce per step
    total += tot / steps
out["_step_total"] = int(total)
out["_source"] = tag
dst = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "dram_traffic.json")
cur = json.load(open(dst)) if os.path.exists(dst) else {}
if "gat_fwd" in cur:   # the flat round-1 layout
    cur = {"_round1_proteins_1gpu": cur}
cur.setdefault(shape, {})[str(world)] = out
json.dump(cur, open(dst, "w"), indent=1)
print(json.dumps(out, indent=1))
