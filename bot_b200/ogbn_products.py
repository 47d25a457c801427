"""``from bot_b200.ogbn_products import GAT`` replaces ``from models import GAT`` in
src/ogbn-products/gat.py (reference module: src/ogbn-products/models.py)."""
from .sampled import GATConv, ProductsGAT as GAT  # noqa: F401
