"""Developer tool: LDG row gather vs cp.async.bulk row gather from an L2-resident slab (tools/bulk_gather.cu)."""
import ctypes
import json
import os

import torch

here = os.path.dirname(os.path.abspath(__file__))
lib = ctypes.CDLL(os.path.join(here, "libbulkgather.so"))
lib.gather_ms.restype = ctypes.c_double
lib.gather_ms.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_int]
dev = torch.device("cuda", 0)
N, LD, D, E = 132534, 480, 80, 39561252
table = torch.randn(N, LD, device=dev)
idx = torch.randint(0, N, (E,), device=dev, dtype=torch.int32)
torch.cuda.synchronize()
res = {}
names = {0: "ldg_8lanes_4steps", 1: "bulk_16rows_x3", 2: "bulk_32rows_x2", 3: "bulk_8rows_x4", 4: "ldg256_4lanes_2steps"}
for variant, bps_list in ((0, (4, 8)), (4, (4, 8)), (1, (3,))):
    for bps in bps_list:
        ms = min(lib.gather_ms(table.data_ptr(), LD, D, idx.data_ptr(), E, variant, 148 * bps) for _ in range(3))
        res[f"{names[variant]}_{bps}blk"] = {"ms": round(ms, 3), "GBs": round(E * D * 4 / (ms * 1e-3) / 1e9, 1) if ms > 0 else None}
print(json.dumps(res, indent=1))
