"""Module-level oracle: the reference's two GATConv.forward bodies restated on top of
``oracle.gat_ref.gat_sparse`` — TEST INFRASTRUCTURE ONLY.

Checked against golden vectors recorded from the UNMODIFIED reference modules
(tests/golden/*.pt, made by tests/golden/make_golden.py through oracle/dgl_shim.py), so the
orchestration around the sparse section — where the degree scaling is applied, which tensors
are aliased before scaling, the +0.5 exponent, the residual — is pinned to the reference's own
code: src/no-sampling/models.py:475-566 (V1) and src/ogbn-proteins/models.py:87-168 (V2).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import gat_ref


def keep_from_perm(perm, n_edges, edge_drop):
    """models.py:529-532: eids = perm[int(E*p):] are kept."""
    bound = int(n_edges * edge_drop)
    keep = torch.ones(n_edges, dtype=torch.bool)
    keep[perm[:bound]] = False
    return keep, perm[bound:]


def attn_mul_full(mul_rows, n_edges, n_heads, eids):
    """Recorded dropout multiplier (rows in `eids` order when edge-drop is on) -> (E,H) edge-id order."""
    if mul_rows is None:
        return None
    m = mul_rows.reshape(-1, n_heads)
    if eids is None:
        return m
    full = torch.zeros(n_edges, n_heads, dtype=m.dtype)
    full[eids] = m
    return full


def gatconv_v1(sd, src, dst, n_src, n_dst, feat, *, num_heads, out_feats, negative_slope=0.2,
               use_symmetric_norm=False, is_block=False, keep=None, attn_mul=None):
    """src/no-sampling/models.py:475-566 with parameters from a reference state_dict `sd`."""
    H, D = num_heads, out_feats
    ft = F.linear(feat, sd["fc.weight"]).view(-1, H, D)                       # :492
    h_dst, ft_dst = (feat[:n_dst], ft[:n_dst]) if is_block else (feat, ft)    # :493-498
    src_scale = dst_scale = None
    if use_symmetric_norm:                                                    # :500-505, :550-555
        src_scale = torch.bincount(src, minlength=n_src).to(ft.dtype).clamp(min=1).pow(-0.5)
        dst_scale = torch.bincount(dst, minlength=n_dst).to(ft.dtype).clamp(min=1).pow(0.5)
    ft_s = ft if src_scale is None else ft * src_scale.view(-1, 1, 1)
    el = (ft_s * sd["attn_l"]).sum(-1)                                        # :517 (scaled ft)
    er = (ft_dst * sd["attn_r"]).sum(-1) if "attn_r" in sd else None          # :521 (unscaled)
    rst = gat_ref.gat_sparse(src, dst, n_dst, ft, el, er, None, keep, attn_mul, negative_slope, src_scale, dst_scale)
    if "res_fc.weight" in sd:                                                 # :557-560
        rst = rst + F.linear(h_dst, sd["res_fc.weight"]).view(h_dst.shape[0], -1, D)
    return rst


def gatconv_v2(sd, src, dst, n_src, n_dst, feat_src, feat_edge=None, *, n_heads, out_feats, negative_slope=0.2,
               use_symmetric_norm=False, is_block=False, deg=None, keep=None, attn_mul=None):
    """src/ogbn-proteins/models.py:87-168 with parameters from a reference state_dict `sd`."""
    H, D = n_heads, out_feats
    feat_dst = feat_src[:n_dst] if is_block else feat_src                     # :93-96
    dst_scale = None
    if use_symmetric_norm:                                                    # :98-104, :150-156
        feat_src = feat_src * deg.pow(-0.5).view(-1, 1)
        dst_scale = deg[:n_dst].pow(0.5)
    ft = F.linear(feat_src, sd["src_fc.weight"]).view(-1, H, D)               # :106
    resid = F.linear(feat_dst, sd["dst_fc.weight"], sd["dst_fc.bias"]).view(-1, H, D)   # :107
    el = F.linear(feat_src, sd["attn_src_fc.weight"])                         # :108
    er = F.linear(feat_dst, sd["attn_dst_fc.weight"]) if "attn_dst_fc.weight" in sd else None   # :122-124
    ee = F.linear(feat_edge, sd["attn_edge_fc.weight"]) if feat_edge is not None else None      # :130-131
    rst = gat_ref.gat_sparse(src, dst, n_dst, ft, el, er, ee, keep, attn_mul, negative_slope, None, dst_scale)
    return rst + resid                                                        # :159-160


def edge_mlp_logits(efeat, w1, b1, w2):
    """Per-layer edge term of the proteins model, caller + layer side in one expression:
    ``efeat_emb = relu(edge_encoder[i](efeat))`` (src/ogbn-proteins/models.py:245-247) then
    ``attn_edge_fc(efeat_emb)`` (models.py:131).  (E, C) -> (E, H)."""
    return F.linear(F.relu(F.linear(efeat, w1, b1)), w2)
