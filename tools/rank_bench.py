"""Developer tool: the per-rank kernel workload of an N-way partitioned proteins-shape layer, on ONE GPU
(no collectives): times gat_fused forward/backward on rank 0's local block graph."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from bot_b200 import functional, partition  # noqa: E402
from bot_b200.functional import gat_fused  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--world", type=int, default=8)
ap.add_argument("--iters", type=int, default=5)
args = ap.parse_args()
dev = torch.device("cuda", 0)
src, dst = bench.synth_edges(bench.N_NODES, bench.N_EDGES, dev)
pg = partition.PartitionedGraph(src, dst, bench.N_NODES, world=args.world, rank=0, plan="dense")
g = pg.local
g.create_formats_()
E, ns, nd = g.number_of_edges(), g.number_of_src_nodes(), g.number_of_dst_nodes()
H, D = bench.HEADS, bench.HID
gen = torch.Generator(device=dev).manual_seed(1)
ft = torch.randn(ns, H, D, device=dev, generator=gen).requires_grad_(True)
el = torch.randn(ns, H, device=dev, generator=gen).requires_grad_(True)
er = torch.randn(nd, H, device=dev, generator=gen).requires_grad_(True)
ee = torch.randn(E, 8, device=dev, generator=gen).requires_grad_(True)
gout = torch.randn(nd, H, D, device=dev, generator=gen)
keep = (torch.rand(E, device=dev, generator=gen) >= 0.1).to(torch.uint8)


def step():
    out = gat_fused(g, ft, el, er, ee, keep)
    out.backward(gout)
    for t in (ft, el, er, ee):
        t.grad = None


for _ in range(2):
    step()
torch.cuda.synchronize()
kt = functional.KernelTimer()
functional.timer = kt
for _ in range(args.iters):
    step()
torch.cuda.synchronize()
functional.timer = None
print(json.dumps({"world": args.world, "E_local": E, "n_src": ns, "n_dst": nd,
                  "kernels_ms": {k: round(t / n, 4) for k, (n, t) in kt.totals().items()},
                  "env": {k: v for k, v in os.environ.items() if k.startswith("BOTGAT_")}}))
