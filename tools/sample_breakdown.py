"""Developer tool: where a sampled batch's preparation time goes (proteins training configuration)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import bot_b200  # noqa: E402
from bot_b200 import sampling  # noqa: E402

dev = torch.device("cuda", 0)
src, dst = bench.synth_edges(bench.N_NODES, bench.N_EDGES, dev)
g = bot_b200.Graph(src, dst, bench.N_NODES)
g.create_formats_()
seeds0 = torch.randperm(bench.N_NODES, device=dev)[:8662]


def tick():
    torch.cuda.synchronize()
    return time.perf_counter()


for it in range(3):
    t = {"sample": 0.0, "compact": 0.0, "csr": 0.0}
    seeds = seeds0
    for layer in range(6):
        a = tick()
        s, d, e, _ = sampling.sample_neighbors(g, seeds, 32, seed=it * 10 + layer)
        b = tick()
        blk = sampling.to_block(g, seeds, s, d, e)
        c = tick()
        blk.create_formats_()
        dd = tick()
        t["sample"] += b - a
        t["compact"] += c - b
        t["csr"] += dd - c
        seeds = blk.srcdata[sampling.NID]
    print({k: round(v * 1e3, 2) for k, v in t.items()}, "ms for 6 layers; last block edges", blk.number_of_edges())
