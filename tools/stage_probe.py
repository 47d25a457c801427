"""Developer tool: cost of the edge staging pass with and without the keep mask (proteins shape)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import bot_b200  # noqa: E402
from bot_b200 import _lib, functional  # noqa: E402

dev = torch.device("cuda", 0)
src, dst = bench.synth_edges(bench.N_NODES, bench.N_EDGES, dev)
g = bot_b200.Graph(src, dst, bench.N_NODES)
g.create_formats_()
E, H = bench.N_EDGES, bench.HEADS
ee = torch.randn(E, 8, device=dev)
keep = functional.edge_drop_keep(E, E // 10, 1, dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, kw in (("ee + keep", dict(ee=ee, keep=keep)), ("ee only", dict(ee=ee)), ("keep only", dict(keep=keep))):
    for order in (_lib.ORDER_IN, _lib.ORDER_OUT):
        for _ in range(3):
            e0.record()
            functional.edge_stage(g, order, H, **kw)
            e1.record()
            torch.cuda.synchronize()
        print("%-10s order %d: %.3f ms" % (name, order, e0.elapsed_time(e1)))
