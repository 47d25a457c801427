"""Developer tool: full-graph INFERENCE of the reference models (model.eval(), torch.no_grad()) with the layer tail
(residual adds, eval-mode BatchNorm / bias, ReLU) fused into the forward kernel's epilogue against the op-by-op tail
(BOTGAT_FUSE_TAIL=0).  Shapes: products (BASELINE config 5: 3-layer GAT, 4 heads x 120), Reddit (3 x 64, V1 GAT), proteins."""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import bot_b200  # noqa: E402

dev = torch.device("cuda", 0)
res = {}
for shape in sys.argv[1:] or ["products", "reddit", "proteins"]:
    c = bench.SHAPES[shape]
    n, e = c["n"], c["e"]
    src, dst = bench.synth_edges(n, e, dev)
    g = bot_b200.Graph(src, dst, n)
    g.create_formats_()
    del src, dst
    torch.manual_seed(0)
    if shape == "products":
        from bot_b200.ogbn_products import GAT
        model = GAT(100, 0, 47, 3, c["H"], c["D"], 0, F.relu, 0.5, 0.1, 0.0, 0.1, allow_zero_in_degree=True, residual=True).to(dev)
        g.srcdata["feat"] = torch.randn(n, 100, device=dev)
        call = lambda: model(g)
    elif shape == "proteins":
        from bot_b200.ogbn_proteins import GAT
        model = GAT(8, 8, 112, 6, c["H"], c["D"], 16, F.relu, 0.25, 0.1, 0.0, 0.1, allow_zero_in_degree=True).to(dev)
        g.srcdata["feat"] = torch.randn(n, 8, device=dev)
        g.edata["feat"] = torch.rand(e, 8, device=dev)
        call = lambda: model(g)
    else:
        from bot_b200.no_sampling import GAT
        model = GAT(602, 0, 41, c["D"], 3, c["H"], F.relu, norm="batch", dropout=0.5, use_symmetric_norm=True, residual=True).to(dev)
        for conv in model.convs:
            conv.set_allow_zero_in_degree(True)
        x = torch.randn(n, 602, device=dev)
        call = lambda: model(g, x)
    model.eval()
    out = {}
    for fuse in ("1", "0"):
        os.environ["BOTGAT_FUSE_TAIL"] = fuse
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.no_grad():
            for _ in range(2):
                y = call()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(5):
                y = call()
            e1.record()
            torch.cuda.synchronize()
        out["fused_ms" if fuse == "1" else "unfused_ms"] = round(e0.elapsed_time(e1) / 5, 3)
        out["y" + fuse] = y
    out["max_abs_diff"] = float((out.pop("y1") - out.pop("y0")).abs().max())
    res[shape] = out
    print(shape, json.dumps(out), flush=True)
    del g, model, y
    torch.cuda.empty_cache()
print(json.dumps(res))
