"""Developer tool: measure the L2->SM read bandwidth for an L2-resident working set (tools/l2peak.cu)."""
import ctypes
import json
import os
import sys

import torch

torch.cuda.init()
lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "libl2peak.so"))
lib.l2_read_gbs.restype = ctypes.c_double
lib.l2_read_gbs.argtypes = [ctypes.c_int64, ctypes.c_int, ctypes.c_int]
res = {}
for mb in (16, 32, 64):
    for bps in (2, 4, 8):
        res[f"{mb}MB_x{bps}"] = round(max(lib.l2_read_gbs(mb << 20, 200, bps) for _ in range(3)), 1)
res["dram_1GB"] = round(lib.l2_read_gbs(1 << 30, 10, 8), 1)
print(json.dumps(res))
