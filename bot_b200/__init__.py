"""bot_b200 — B200-native GAT message-passing engine, drop-in for the GATConv hot path
of AiRyunn/BoT (src/no-sampling, src/ogbn-proteins, src/ogbn-products).

Layout: ``csrc/`` hand-written sm_100a kernels + the C ABI (``include/botgat.h``),
``_lib`` ctypes binding, ``graph`` the DGLGraph-subset object, ``functional`` the
autograd seam, ``no_sampling`` / ``ogbn_proteins`` / ``ogbn_products`` the mirrors of
the reference's ``models.py`` files, ``partition`` the multi-GPU path, ``feed`` the double-buffered host-to-device input feed,
``sampling`` device neighbour sampling / block construction (the reference's NodeDataLoader path).
There is no CPU fallback: every compute entry point raises if libbotgat.so is absent
or the tensors are not on a CUDA device.
"""
from . import _lib  # noqa: F401
from .graph import Graph, add_self_loop, create_block, graph, remove_self_loop, to_bidirected  # noqa: F401
from .functional import Deferred, EdgeEmbedding, EdgeMLPLogits, GATFusedFn, edge_logits, gat_fused  # noqa: F401
from .feed import HostFeed  # noqa: F401
from . import sampling  # noqa: F401

__all__ = ["Graph", "graph", "create_block", "to_bidirected", "remove_self_loop", "add_self_loop",
           "GATFusedFn", "gat_fused", "edge_logits", "EdgeMLPLogits", "EdgeEmbedding", "Deferred", "HostFeed"]
