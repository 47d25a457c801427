"""Developer tool: kernel-level breakdown of the end-to-end GATConv step (torch.profiler)."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import bot_b200  # noqa: E402
from bot_b200.ogbn_proteins import GATConv  # noqa: E402

dev = torch.device("cuda", 0)
src, dst = bench.synth_edges(bench.N_NODES, bench.N_EDGES, dev)
graph = bot_b200.Graph(src, dst, bench.N_NODES)
graph.create_formats_()
torch.manual_seed(0)
conv = GATConv(bench.HEADS * bench.HID, bench.EDGE_EMB, bench.HID, n_heads=bench.HEADS, edge_drop=bench.EDGE_DROP).to(dev)
h_host = torch.randn(bench.N_NODES, bench.HEADS * bench.HID).pin_memory()
fe_host = torch.randn(bench.N_EDGES, bench.EDGE_EMB).pin_memory()


def step():
    h = h_host.to(dev, non_blocking=True).requires_grad_(True)
    fe = fe_host.to(dev, non_blocking=True).requires_grad_(True)
    y = conv(graph, h, fe)
    loss = y.square().mean()
    loss.backward()
    conv.zero_grad(set_to_none=True)
    return float(loss.item())


for _ in range(2):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
