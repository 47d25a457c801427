"""Developer tool: cost of the per-edge record passes (stage in / stage out / unstage / reduce) at the proteins shape,
for a sweep of the cache-blocked out-CSR traversal (BOTGAT_TILES = "src blocks,dst blocks"; "1,1" = plain order)."""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import bot_b200  # noqa: E402
from bot_b200 import _lib, functional  # noqa: E402
from bot_b200.graph import _stream  # noqa: E402

dev = torch.device("cuda", 0)
N, E, H = bench.N_NODES, bench.N_EDGES, bench.HEADS
src, dst = bench.synth_edges(N, E, dev)
ee = torch.randn(E, 8, device=dev)
keep = functional.edge_drop_keep(E, E // 10, 1, dev)
gz = torch.randn(H, E, device=dev)
gee = torch.empty(E, 8, device=dev)
ger = torch.empty(N, H, device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
lib = _lib.load()


def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return round(best, 3)


res = {}
for tiles in sys.argv[1:] or ["1,1", "4,4", "8,8", "16,16", "12,12", "8,4", "4,8", "16,8", "8,16", "32,32"]:
    os.environ["BOTGAT_TILES"] = tiles
    g = bot_b200.Graph(src, dst, N)
    h = g._ensure()
    r = {"tiles": [g._info.tiles_src, g._info.tiles_dst]}
    r["stage_in_ee_keep"] = timed(lambda: functional.edge_stage(g, _lib.ORDER_IN, H, ee=ee, keep=keep))
    r["stage_out_ee_keep"] = timed(lambda: functional.edge_stage(g, _lib.ORDER_OUT, H, ee=ee, keep=keep))
    r["stage_out_ee"] = timed(lambda: functional.edge_stage(g, _lib.ORDER_OUT, H, ee=ee))
    r["unstage_out"] = timed(lambda: lib.botgat_edge_unstage(h, _lib.ORDER_OUT, H, gz.data_ptr(), gee.data_ptr(), 8, _stream()))
    r["reduce_dst"] = timed(lambda: lib.botgat_edge_reduce_dst(h, H, gee.data_ptr(), 8, ger.data_ptr(), None, _stream()))
    res[tiles] = r
    print(tiles, json.dumps(r), flush=True)
    del g
