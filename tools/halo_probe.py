"""Developer tool (torchrun, >= 2 GPUs): bandwidth of the peer-memory halo kernels (csrc/halo.cu) against NCCL's
all-gather / reduce-scatter of the same table, for several grid sizes."""
import ctypes as C
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bot_b200 import _lib  # noqa: E402
from bot_b200.graph import _stream  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
import torch.distributed._symmetric_memory as symm_mem  # noqa: E402

lib = _lib.load()
rows_total, P = int(sys.argv[1]) if len(sys.argv) > 1 else 2449029, 512
rows = (rows_total + world - 1) // world
table = symm_mem.empty((world * rows, P), dtype=torch.float32, device=dev)
gtable = symm_mem.empty((world * rows, P), dtype=torch.float32, device=dev)
hs, hg = symm_mem.rendezvous(table, dist.group.WORLD), symm_mem.rendezvous(gtable, dist.group.WORLD)
table.normal_()
gtable.normal_()
shard = table[rank * rows:(rank + 1) * rows]
sp = (C.c_void_p * world)(*[int(p) for p in hs.buffer_ptrs])
gp = (C.c_void_p * world)(*[int(p) for p in hg.buffer_ptrs])
gshard = torch.empty((rows, P), device=dev)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)


def timed(fn, reps=4):
    best = 1e9
    for _ in range(reps):
        dist.barrier()
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    t = torch.tensor([best], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


res = {"world": world, "rows_per_rank": rows, "row_bytes": P * 4, "remote_GB_per_rank": (world - 1) * rows * P * 4 / 1e9}
for blocks in (148, 296, 592, 1184, 2368):
    ms = timed(lambda: lib.botgat_halo_pull(world, rank, sp, rows, P, 0, P, blocks, _stream()))
    ms2 = timed(lambda: lib.botgat_halo_pull_reduce(world, rank, gp, rows, P, 0, P, gshard.data_ptr(), P, blocks, _stream()))
    res[f"blocks_{blocks}"] = {"pull_ms": round(ms, 3), "pull_remote_GBs": round(res["remote_GB_per_rank"] / ms * 1e3, 1),
                               "pull_reduce_ms": round(ms2, 3), "reduce_remote_GBs": round(res["remote_GB_per_rank"] / ms2 * 1e3, 1)}
out = torch.empty((world * rows, P), device=dev)
ms = timed(lambda: dist.all_gather_into_tensor(out, shard))
rs = torch.empty((rows, P), device=dev)
ms2 = timed(lambda: dist.reduce_scatter_tensor(rs, gtable))
res["nccl"] = {"all_gather_ms": round(ms, 3), "ag_remote_GBs": round(res["remote_GB_per_rank"] / ms * 1e3, 1),
               "reduce_scatter_ms": round(ms2, 3), "rs_remote_GBs": round(res["remote_GB_per_rank"] / ms2 * 1e3, 1)}
ms = timed(lambda: hs.barrier(channel=0))
res["symm_barrier_ms"] = round(ms, 4)
if rank == 0:
    print(json.dumps(res, indent=1))
dist.destroy_process_group()
