"""Developer tool: where the end-to-end GATConv step spends its time (CUDA events):
host->device copy alone, the layer alone on resident inputs, and torch.profiler's top kernels of the layer."""
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
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(fn, k=5, w=2):
    for _ in range(w):
        fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(k):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


hd = torch.empty_like(h_host, device=dev)
fd = torch.empty_like(fe_host, device=dev)


def copy_only():
    hd.copy_(h_host, non_blocking=True)
    fd.copy_(fe_host, non_blocking=True)


def layer_only(grad_fe=True):
    h = hd.detach().requires_grad_(True)
    fe = fd.detach().requires_grad_(grad_fe)
    loss = conv(graph, h, fe).square().mean()
    loss.backward()
    conv.zero_grad(set_to_none=True)
    return float(loss.item())


print("copy only      %.2f ms" % timed(copy_only))
print("layer only     %.2f ms (feat_edge requires grad)" % timed(layer_only))
print("layer only     %.2f ms (feat_edge is a plain input)" % timed(lambda: layer_only(False)))
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    layer_only()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=60))
