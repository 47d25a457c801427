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

# tile::gather4 (UTMALDG): four rows per request through a tensor map over the table
lib.gather4_ms.restype = ctypes.c_double
lib.gather4_ms.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
                           ctypes.c_int, ctypes.c_int]
lib.gather4_checksum.restype = ctypes.c_double
lib.gather4_checksum.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64,
                                 ctypes.c_int, ctypes.c_int]
g4 = {}
n_small = 1 << 20
want = float(table[idx[:n_small].long(), :D].double().sum())
for box_rows in (1,):   # a box of 4 rows is rejected by the hardware (illegal instruction): gather4 wants box = (cols, 1)
    got = lib.gather4_checksum(table.data_ptr(), LD, D, N, idx.data_ptr(), n_small, 148 * 3, box_rows)
    g4[f"checksum_box_rows_{box_rows}"] = {"got": got, "want": want, "ok": abs(got - want) <= 1e-3 * max(1.0, abs(want)) + 50.0}
ok_box = [b for b in (1,) if g4[f"checksum_box_rows_{b}"]["ok"]]
names4 = {5: "gather4_16rows_x3", 6: "gather4_32rows_x2", 7: "gather4_8rows_x4", 8: "gather4_16rows_x2"}
for box_rows in ok_box[:1]:
    for variant, bps_list in ((5, (3, 4)), (6, (2, 3)), (7, (4, 6)), (8, (4, 5))):
        for bps in bps_list:
            ms = min(lib.gather4_ms(table.data_ptr(), LD, D, N, idx.data_ptr(), E, variant, 148 * bps, box_rows) for _ in range(3))
            g4[f"{names4[variant]}_{bps}blk"] = {"ms": round(ms, 3), "GBs": round(E * D * 4 / (ms * 1e-3) / 1e9, 1) if ms > 0 else None}
print(json.dumps(g4, indent=1))
