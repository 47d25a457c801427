"""Developer tool: cost of the edge-drop draw at several edge counts."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bot_b200.functional import edge_drop_keep  # noqa: E402

dev = torch.device("cuda", 0)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for E in (13264, 2484941, 4945269, 9890313, 39561252, 114615892):
    for it in range(4):
        e0.record()
        k = edge_drop_keep(E, E // 10, it, dev)
        e1.record()
        torch.cuda.synchronize()
    print("E = %11d: %.3f ms" % (E, e0.elapsed_time(e1)))
