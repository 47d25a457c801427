"""Developer tool: kernel-only timing of the fused forward / backward at a named shape.
Not the contract benchmark (that is bench.py); used for parameter sweeps under gpurun."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bot_b200  # noqa: E402
from bot_b200.functional import gat_fused  # noqa: E402

SHAPES = {
    # name: N, E, H, D, er, ee, symm, edge_drop
    "proteins": (132534, 39561252, 6, 80, True, True, False, 0.1),
    "products": (2449029, 61859140, 4, 120, True, False, False, 0.1),
    "reddit": (232965, 114615892, 4, 64, False, False, True, 0.0),
    "arxiv": (169343, 2484941, 3, 250, False, False, True, 0.0),
    "arxiv_last": (169343, 2484941, 1, 40, False, False, True, 0.0),
    "cora": (2708, 13264, 8, 8, False, False, False, 0.5),
    "small": (20000, 2000000, 6, 80, True, True, False, 0.1),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="proteins")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--no-bwd", action="store_true")
    ap.add_argument("--edge-drop", type=float, default=None)
    ap.add_argument("--power-law", type=float, default=0.0)
    ap.add_argument("--pad-ee", type=int, default=0, help="row width of ee (0 = H, unpadded)")
    args = ap.parse_args()
    N, E, H, D, has_er, has_ee, symm, edrop = SHAPES[args.shape]
    if args.edge_drop is not None:
        edrop = args.edge_drop
    dev = torch.device("cuda", 0)
    if os.environ.get("L2_FETCH"):
        import ctypes
        torch.cuda.init()
        rt = ctypes.CDLL("libcudart.so.12")
        rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(int(os.environ["L2_FETCH"])))  # cudaLimitMaxL2FetchGranularity
        v = ctypes.c_size_t()
        rt.cudaDeviceGetLimit(ctypes.byref(v), 5)
        print("L2 fetch granularity limit ->", rc, v.value, file=sys.stderr)
    g = torch.Generator(device=dev).manual_seed(0)
    src = torch.randint(0, N, (E,), device=dev, generator=g)
    if args.power_law > 0:
        w = torch.arange(1, N + 1, device=dev, dtype=torch.float64) ** (-args.power_law)
        dst = torch.multinomial((w / w.sum()).float(), E, replacement=True, generator=g)
    else:
        dst = torch.randint(0, N, (E,), device=dev, generator=g)
    t0 = time.time()
    gr = bot_b200.Graph(src, dst, N)
    gr.create_formats_()
    torch.cuda.synchronize()
    t_ingest = time.time() - t0
    ft = torch.randn(N, H, D, device=dev, generator=g).requires_grad_(True)
    el = torch.randn(N, H, device=dev, generator=g).requires_grad_(True)
    er = torch.randn(N, H, device=dev, generator=g).requires_grad_(True) if has_er else None
    ee = torch.randn(E, args.pad_ee or H, device=dev, generator=g).requires_grad_(True) if has_ee else None
    keep = None
    if edrop > 0:
        keep = (torch.rand(E, device=dev, generator=g) >= edrop).to(torch.uint8)
    cs = gr.deg_scale("out", -0.5) if symm else None
    ds = gr.deg_scale("in", 0.5) if symm else None
    gout = torch.randn(N, H, D, device=dev, generator=g)

    def step():
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        out = gat_fused(gr, ft, el, er, ee, keep, None, cs, ds, 0.2, 0.0, 0)
        e1.record()
        if not args.no_bwd:
            out.backward(gout)
        e2.record()
        torch.cuda.synchronize()
        ft.grad = el.grad = None
        if er is not None:
            er.grad = None
        if ee is not None:
            ee.grad = None
        return e0.elapsed_time(e1), e1.elapsed_time(e2)

    for _ in range(args.warmup):
        step()
    ts = [step() for _ in range(args.iters)]
    f = sorted(t[0] for t in ts)[len(ts) // 2]
    b = sorted(t[1] for t in ts)[len(ts) // 2]
    R = 4 * H * D
    res = {
        "shape": args.shape, "N": N, "E": E, "H": H, "D": D, "ingest_s": round(t_ingest, 3),
        "fwd_ms": round(f, 3), "bwd_ms": round(b, 3),
        "edges_per_s_fwd_bwd": E / ((f + b) * 1e-3),
        "fwd_gather_GBs": E * R / (f * 1e-3) / 1e9,
        "env": {k: v for k, v in os.environ.items() if k.startswith("BOTGAT_")},
    }
    print(json.dumps(res))


if __name__ == "__main__":
    main()
