"""Developer tool: one forward + backward of the per-edge projection and of the fused edge MLP (for ncu)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bot_b200 import functional  # noqa: E402

E = int(os.environ.get("E", 39561252))
dev = torch.device("cuda", 0)
gy = torch.randn(E, 8, device=dev)
x = torch.randn(E, 16, device=dev).requires_grad_(True)
w = torch.randn(6, 16, device=dev).requires_grad_(True)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
raw = torch.randn(E, 8, device=dev)
w1 = torch.randn(16, 8, device=dev).requires_grad_(True)
b1 = torch.randn(16, device=dev).requires_grad_(True)
w2 = torch.randn(6, 16, device=dev).requires_grad_(True)
for _ in range(int(os.environ.get("ITERS", 1))):
    ev[0].record()
    y = functional.edge_logits(x, w)
    ev[1].record()
    y.backward(gy)
    ev[2].record()
    y2 = functional.EdgeMLPLogits.apply(raw, w1, b1, w2)
    ev[3].record()
    y2.backward(gy)
    ev[4].record()
    torch.cuda.synchronize()
    print("proj fwd %.3f bwd %.3f | mlp fwd %.3f bwd %.3f ms" % (
        ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3]), ev[3].elapsed_time(ev[4])))
