"""``from bot_b200.ogbn_proteins import GAT`` replaces ``from models import GAT`` in
src/ogbn-proteins/gat.py:23 (reference module: src/ogbn-proteins/models.py)."""
from .sampled import GATConv, ProteinsGAT as GAT  # noqa: F401
