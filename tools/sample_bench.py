"""Developer tool: device neighbour sampling at the proteins training configuration
(src/ogbn-proteins/gat.py:177-188: 6 layers, fanout 32, batch = |train| / 10 seeds) and at the evaluation one
(fanout 100, batch 65536) — time per batch for sampling + block construction, and for one training step on the blocks."""
import os
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import bot_b200  # noqa: E402
from bot_b200 import sampling  # noqa: E402
from bot_b200.ogbn_proteins import GAT  # noqa: E402

dev = torch.device("cuda", 0)
src, dst = bench.synth_edges(bench.N_NODES, bench.N_EDGES, dev)
g = bot_b200.Graph(src, dst, bench.N_NODES)
g.create_formats_()
del src, dst
g.ndata["feat"] = torch.randn(bench.N_NODES, 8, device=dev)
g.ndata["labels"] = (torch.rand(bench.N_NODES, 112, device=dev) > 0.5).float()
g.edata["feat"] = torch.rand(bench.N_EDGES, 8, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fan, batch in (("train", 32, 8662), ("eval", 100, 65536)):
    sampler = sampling.MultiLayerNeighborSampler([fan] * 6)
    seeds = torch.randperm(bench.N_NODES, device=dev)[:batch]
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        blocks = sampler.sample_blocks(g, seeds, seed=it)
        for b in blocks:
            b.create_formats_()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
    print("%s: fanout %d, %d seeds: %.2f ms per batch (sample + compact + CSR build, 6 layers); block edges %s, src nodes %s" % (
        name, fan, batch, dt, [b.number_of_edges() for b in blocks], [b.number_of_src_nodes() for b in blocks]))
    if name == "train":
        torch.manual_seed(0)
        model = GAT(8, 8, 112, 6, bench.HEADS, bench.HID, bench.EDGE_EMB, F.relu, 0.25, 0.1, 0.0, 0.1).to(dev).train()
        for it in range(3):
            e0.record()
            pred = model(blocks)
            loss = F.binary_cross_entropy_with_logits(pred, blocks[-1].dstdata["labels"])
            loss.backward()
            e1.record()
            torch.cuda.synchronize()
            model.zero_grad(set_to_none=True)
        print("train step on these blocks (fwd + bwd): %.2f ms" % e0.elapsed_time(e1))
