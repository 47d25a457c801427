"""Developer tool: the per-edge logit projection kernels at the proteins shape (E = 39.6 M, C = 16, H = 6)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bot_b200.functional import edge_logits, timer  # noqa: E402,F401
from bot_b200 import functional  # noqa: E402

E, C, H = int(os.environ.get("E", 39561252)), int(os.environ.get("C", 16)), int(os.environ.get("H", 6))
dev = torch.device("cuda", 0)
x = torch.randn(E, C, device=dev).requires_grad_(True)
w = torch.randn(H, C, device=dev).requires_grad_(True)
gy = torch.randn(E, functional.pad_heads(H), device=dev)
e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for it in range(4):
    e[0].record()
    y = edge_logits(x, w)
    e[1].record()
    y.backward(gy)
    e[2].record()
    torch.cuda.synchronize()
    x.grad = w.grad = None
    print("fwd %.3f ms  bwd (gx + gw) %.3f ms" % (e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))
gb = E * 4 * (C + functional.pad_heads(H)) / 1e9
print("ideal at 6555 GB/s: fwd %.3f ms, gx %.3f ms, gw %.3f ms" % (gb / 6.555, gb / 6.555, gb / 6.555))
