"""Developer tool: one full-graph training step (forward + backward) of the 6-layer proteins model
(bot_b200.ogbn_proteins.GAT, gat.py:320-327 defaults) at the proteins shape, CUDA events.
BOTGAT_NO_EDGE_MLP=1 materialises the per-layer edge embedding as the reference does."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import bot_b200  # noqa: E402
from bot_b200.ogbn_proteins import GAT  # noqa: E402

dev = torch.device("cuda", 0)
n_layers = int(os.environ.get("LAYERS", 6))
src, dst = bench.synth_edges(bench.N_NODES, bench.N_EDGES, dev)
graph = bot_b200.Graph(src, dst, bench.N_NODES)
graph.create_formats_()
del src, dst
torch.manual_seed(0)
model = GAT(8, 8, 112, n_layers, bench.HEADS, bench.HID, bench.EDGE_EMB, F.relu, 0.25, 0.1, 0.0, 0.1).to(dev)
model.train()
graph.srcdata["feat"] = torch.randn(bench.N_NODES, 8, device=dev)
graph.edata["feat"] = torch.rand(bench.N_EDGES, 8, device=dev)
labels = (torch.rand(bench.N_NODES, 112, device=dev) > 0.5).float()
e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
for it in range(4):
    torch.cuda.reset_peak_memory_stats()
    e0.record()
    loss = F.binary_cross_entropy_with_logits(model(graph), labels)
    e1.record()
    loss.backward()
    e2.record()
    torch.cuda.synchronize()
    model.zero_grad(set_to_none=True)
    print("fwd %.2f ms  bwd %.2f ms  total %.2f ms  peak mem %.1f GB" % (
        e0.elapsed_time(e1), e1.elapsed_time(e2), e0.elapsed_time(e2), torch.cuda.max_memory_allocated() / 2**30))
if os.environ.get("PROFILE"):
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        F.binary_cross_entropy_with_logits(model(graph), labels).backward()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=24, max_name_column_width=60))
