"""GATConv math oracle (pure torch, CPU, fp32 or fp64) — TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED (see oracle/__init__.py).  Restates, with DGL 0.5 op semantics
(SURVEY.md Appendix B), the sparse section of

* V1  src/no-sampling/models.py:517-555    (el/er -> u_add_v -> leaky_relu ->
      [edge-drop] edge_softmax -> attn_drop -> u_mul_e/sum -> dst scale)
* V2  src/ogbn-proteins/models.py:120-156  (same, plus the per-edge logit term
      ``attn_edge`` :130-133)

``gat_sparse`` is the materialising form (index_select / scatter_reduce /
index_add_), differentiable by autograd: it is the parity reference for both the
forward values and the gradients.  ``gat_sparse_big_*`` is a non-materialising
form (segment_reduce + per-head sparse CSR matmul) with an explicit backward,
used only as the timed CPU baseline at shapes where E*H*D does not fit in RAM.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def edge_softmax(dst, e, n_dst):
    """DGL ``edge_softmax(graph, e)`` with norm_by='dst' (SURVEY.md App. B):
    m = max over in-edges; p = exp(e - m[dst]); S = sum; a = p / S[dst]."""
    shape = (n_dst,) + tuple(e.shape[1:])
    idx = dst.view(-1, *([1] * (e.dim() - 1))).expand_as(e)
    m = torch.full(shape, float("-inf"), dtype=e.dtype).scatter_reduce(0, idx, e, "amax", include_self=True)
    p = torch.exp(e - m.index_select(0, dst))
    s = torch.zeros(shape, dtype=e.dtype).index_add_(0, dst, p)
    return p / s.index_select(0, dst)


def gat_sparse(src, dst, n_dst, ft, el, er=None, ee=None, keep=None, attn_mul=None,
               slope=0.2, src_scale=None, dst_scale=None):
    """Sparse section of GATConv.forward.

    src, dst   (E,) int64 COO in edge-id order
    ft         (N_s,H,D) projected source features, NOT yet degree-scaled
    el         (N_s,H)   source logit term (already includes any scaling the caller applied)
    er         (N_d,H)   or None                       models.py:521-523 / proteins :122-125
    ee         (E,H)     or None, edge-id order        proteins models.py:130-133
    keep       (E,) bool or None — edge-drop keep set  models.py:528-537
    attn_mul   (E,H) or None — attention-dropout multiplier m/(1-p), edge-id order
               (rows of dropped edges are ignored)     models.py:537/544
    src_scale  (N_s,) or None — clamp(out_deg,1)^-0.5  models.py:500-505
    dst_scale  (N_d,) or None — clamp(in_deg,1)^+0.5   models.py:550-555
    returns    (N_d,H,D)
    """
    H = ft.shape[1]
    if src_scale is not None:
        ft = ft * src_scale.view(-1, 1, 1)
    e = el.index_select(0, src)                                   # copy_u        :525
    if er is not None:
        e = e + er.index_select(0, dst)                           # u_add_v       :523
    if ee is not None:
        e = e + ee                                                # proteins      :133
    e = F.leaky_relu(e, slope)                                    #               :526
    if keep is not None:
        eids = torch.nonzero(keep).flatten()
        a_kept = edge_softmax(dst.index_select(0, eids), e.index_select(0, eids), n_dst)
        if attn_mul is not None:
            a_kept = a_kept * attn_mul.index_select(0, eids)
        a = torch.zeros_like(e).index_copy(0, eids, a_kept)       # a[eids] = ... :534-537
    else:
        a = edge_softmax(dst, e, n_dst)                           #               :544
        if attn_mul is not None:
            a = a * attn_mul
    msg = ft.index_select(0, src) * a.view(-1, H, 1)              # u_mul_e       :547
    out = torch.zeros((n_dst,) + tuple(ft.shape[1:]), dtype=ft.dtype).index_add_(0, dst, msg)
    if dst_scale is not None:
        out = out * dst_scale.view(-1, 1, 1)                      #               :550-555
    return out


# --------------------------------------------------------------------------
# non-materialising form (CPU baseline at large shapes)
# --------------------------------------------------------------------------

class BigGraph:
    """CSR views needed by ``gat_sparse_big_*`` (built once, outside the timed region)."""

    def __init__(self, src, dst, n_src, n_dst):
        self.n_src, self.n_dst = n_src, n_dst
        order = torch.sort(dst, stable=True).indices
        self.in_eid = order
        self.in_src = src.index_select(0, order)
        self.in_dst = dst.index_select(0, order)
        self.in_counts = torch.bincount(dst, minlength=n_dst)
        self.in_indptr = torch.zeros(n_dst + 1, dtype=torch.int64)
        self.in_indptr[1:] = torch.cumsum(self.in_counts, 0)
        order_o = torch.sort(src, stable=True).indices
        self.out_eid = order_o
        # position (in the in-CSR order) of every out-CSR entry
        inv = torch.empty_like(order)
        inv[order] = torch.arange(order.numel())
        self.out_pos_in = inv.index_select(0, order_o)
        self.out_dst = dst.index_select(0, order_o)
        self.out_counts = torch.bincount(src, minlength=n_src)
        self.out_indptr = torch.zeros(n_src + 1, dtype=torch.int64)
        self.out_indptr[1:] = torch.cumsum(self.out_counts, 0)


def gat_sparse_big_forward(g: BigGraph, ft, el, er=None, ee=None, slope=0.2, keep=None, attn_mul=None, src_scale=None,
                           dst_scale=None):
    """Forward without E*H*D temporaries; same arguments as ``gat_sparse`` (``keep`` (E,) bool and ``attn_mul``
    (E,H) in edge-id order).  Returns (out, saved) with ``saved`` = what the explicit backward needs."""
    H = ft.shape[1]
    z = el.index_select(0, g.in_src)
    if er is not None:
        z = z + er.index_select(0, g.in_dst)
    if ee is not None:
        z = z + ee.index_select(0, g.in_eid)
    s = F.leaky_relu(z, slope)
    if keep is not None:                                           # softmax over the kept edges only, models.py:534-537
        s = torch.where(keep.index_select(0, g.in_eid).view(-1, 1), s, torch.full_like(s, float("-inf")))
    m = torch.segment_reduce(s, "max", lengths=g.in_counts, initial=float("-inf"))
    m = torch.where(torch.isinf(m), torch.zeros_like(m), m)        # rows without a kept edge
    p = torch.exp(s - m.index_select(0, g.in_dst))
    ssum = torch.segment_reduce(p, "sum", lengths=g.in_counts, initial=0.0)
    a = p / ssum.index_select(0, g.in_dst).clamp(min=torch.finfo(p.dtype).tiny)
    am = attn_mul.index_select(0, g.in_eid) if attn_mul is not None else None
    at = a if am is None else a * am                               # attention dropout, models.py:537/544
    fts = ft if src_scale is None else ft * src_scale.view(-1, 1, 1)   # models.py:500-505
    acc = torch.empty((g.n_dst,) + tuple(ft.shape[1:]), dtype=ft.dtype)
    for h in range(H):
        A = torch.sparse_csr_tensor(g.in_indptr, g.in_src, at[:, h].contiguous(), size=(g.n_dst, g.n_src))
        acc[:, h, :] = A @ fts[:, h, :]
    out = acc if dst_scale is None else acc * dst_scale.view(-1, 1, 1)  # models.py:550-555
    return out, (a, am, z, acc, fts)


def gat_sparse_big_backward(g: BigGraph, saved, gout, slope=0.2, need_er=True, need_ee=True, keep=None, src_scale=None,
                            dst_scale=None):
    """Explicit adjoint.  Returns grad_ft, grad_el, grad_er, grad_ee."""
    a, am, z, acc, fts = saved
    H = fts.shape[1]
    gp = gout if dst_scale is None else gout * dst_scale.view(-1, 1, 1)
    # d_k = <src_scale * ft[src_k], g'[dst_k]> without materialising E*H*D: chunk over edges
    E = a.shape[0]
    d = torch.empty_like(a)
    chunk = 1 << 20
    for lo in range(0, E, chunk):
        hi = min(E, lo + chunk)
        d[lo:hi] = (fts.index_select(0, g.in_src[lo:hi]) * gp.index_select(0, g.in_dst[lo:hi])).sum(-1)
    t = (acc * gp).sum(-1)                                     # (N_d,H)
    if am is not None:
        d = d * am
    gs = a * (d - t.index_select(0, g.in_dst))
    gz = gs * torch.where(z > 0, torch.ones_like(z), torch.full_like(z, slope))
    if keep is not None:
        gz = torch.where(keep.index_select(0, g.in_eid).view(-1, 1), gz, torch.zeros_like(gz))
    grad_er = torch.segment_reduce(gz, "sum", lengths=g.in_counts, initial=0.0) if need_er else None
    grad_ee = None
    if need_ee:
        grad_ee = torch.empty_like(gz)
        grad_ee[g.in_eid] = gz
    gz_out = gz.index_select(0, g.out_pos_in)
    at = a if am is None else a * am
    a_out = at.index_select(0, g.out_pos_in)
    grad_el = torch.segment_reduce(gz_out, "sum", lengths=g.out_counts, initial=0.0)
    grad_ft = torch.empty_like(fts)
    for h in range(H):
        AT = torch.sparse_csr_tensor(g.out_indptr, g.out_dst, a_out[:, h].contiguous(), size=(g.n_src, g.n_dst))
        grad_ft[:, h, :] = AT @ gp[:, h, :]
    if src_scale is not None:
        grad_ft = grad_ft * src_scale.view(-1, 1, 1)
    return grad_ft, grad_el, grad_er, grad_ee
