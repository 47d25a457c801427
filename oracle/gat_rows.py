"""Row-subsample oracle (fp64) — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Value parity at the full BASELINE sizes: ``gat_ref.gat_sparse`` materialises E x H x D
messages (76 GB at the proteins shape), so the full-size checks evaluate the SAME math on a
SUB-PROBLEM: all in-edges of a chosen set W of destination rows.  Because softmax and sum are
per destination (src/no-sampling/models.py:537-548, src/ogbn-proteins/models.py:141-148) the
sub-problem reproduces, exactly,

* ``out[v]``, ``grad_er[v]`` for every v in W, ``grad_ee[k]`` for every in-edge k of W;
* ``grad_ft[u]``, ``grad_el[u]`` for every source u ALL of whose out-edges end in W
  (``complete`` below; pick W = destinations adjacent to a few sampled sources).

Two evaluators of a sub-problem, both fp64 on the CPU:

* ``eval_autograd``  — ``gat_ref.gat_sparse`` + autograd, head slice by head slice (the
  restatement itself; memory = edges x D x 8 B x a few temporaries);
* ``eval_explicit``  — the adjoint of SURVEY.md Appendix A.3 written out and chunked over
  edges (memory-bounded for any sub-problem, e.g. the 800 K-edge rows of the skewed graph).
  ``tests/test_gat_rows_cpu.py`` pins it to ``eval_autograd`` on small cases with every flag.

``build_sub`` only moves data (mask / nonzero / index_select, on whatever device the full
tensors live on) and converts to fp64 CPU tensors; it uses no structure computed by libbotgat.
"""
from __future__ import annotations

import torch

from . import gat_ref


def build_sub(src, dst, n_src, n_dst, W, *, ft, el, er=None, ee=None, keep=None, attn_mul=None, src_scale=None,
              dst_scale=None, gout=None):
    """Sub-problem of the destination rows ``W`` (1-D int64, unique).  ``src``/``dst`` (E,) int64 COO in edge-id
    order; ``ft`` (N_s, H, D), ``gout`` (N_d, H, D); per-edge operands (E, >= H) in edge-id order (padding columns
    are ignored).  Returns a dict of CPU tensors (float data in fp64)."""
    dev = src.device
    W = torch.as_tensor(W, dtype=torch.int64, device=dev)
    lut = torch.full((n_dst,), -1, dtype=torch.int64, device=dev)
    lut[W] = torch.arange(W.numel(), device=dev)
    ld = lut.index_select(0, dst)
    sel = torch.nonzero(ld >= 0).flatten()                 # edge ids, ascending
    e_dst = ld.index_select(0, sel)
    gsrc = src.index_select(0, sel)
    U, e_src = torch.unique(gsrc, return_inverse=True)     # referenced sources (sorted global ids)
    out_deg_full = torch.bincount(src, minlength=n_src).index_select(0, U)
    complete = out_deg_full == torch.bincount(e_src, minlength=U.numel())
    H = ft.shape[1]

    def f64(t):
        return None if t is None else t.detach().to("cpu", torch.float64)

    def rows(t, idx, width=None):
        if t is None:
            return None
        r = t.index_select(0, idx)
        if width is not None:
            r = r.reshape(r.shape[0], -1)[:, :width]
        return f64(r)

    sub = {
        "n_w": int(W.numel()), "n_u": int(U.numel()), "H": H,
        "W": W.cpu(), "U": U.cpu(), "eid": sel.cpu(), "e_src": e_src.cpu(), "e_dst": e_dst.cpu(), "complete": complete.cpu(),
        "ft": rows(ft, U), "el": rows(el, U, H), "src_scale": rows(src_scale, U),
        "er": rows(er, W, H), "dst_scale": rows(dst_scale, W), "gout": rows(gout, W),
        "ee": rows(ee, sel, H), "attn_mul": rows(attn_mul, sel, H),
        "keep": None if keep is None else keep.index_select(0, sel).to("cpu", torch.bool),
    }
    return sub


def eval_autograd(sub, slope=0.2):
    """``gat_ref.gat_sparse`` on the sub-problem, fp64, autograd gradients; one head at a time (heads are
    independent everywhere in the layer) to bound the E' x D message temporary."""
    H = sub["H"]
    n_w, n_u = sub["n_w"], sub["n_u"]
    D = sub["ft"].shape[2]
    res = {"out": torch.empty(n_w, H, D, dtype=torch.float64), "grad_ft": torch.empty(n_u, H, D, dtype=torch.float64),
           "grad_el": torch.empty(n_u, H, dtype=torch.float64),
           "grad_er": None if sub["er"] is None else torch.empty(n_w, H, dtype=torch.float64),
           "grad_ee": None if sub["ee"] is None else torch.empty(sub["e_src"].numel(), H, dtype=torch.float64)}
    for h in range(H):
        def sl(t):
            return None if t is None else t[:, h:h + 1].clone()

        ft, el = sl(sub["ft"]).requires_grad_(True), sl(sub["el"]).requires_grad_(True)
        er, ee = sl(sub["er"]), sl(sub["ee"])
        for t in (er, ee):
            if t is not None:
                t.requires_grad_(True)
        out = gat_ref.gat_sparse(sub["e_src"], sub["e_dst"], n_w, ft, el, er, ee, sub["keep"], sl(sub["attn_mul"]), slope,
                                 sub["src_scale"], sub["dst_scale"])
        res["out"][:, h:h + 1] = out.detach()
        if sub["gout"] is not None:
            out.backward(sub["gout"][:, h:h + 1])
            res["grad_ft"][:, h:h + 1] = ft.grad
            res["grad_el"][:, h:h + 1] = el.grad
            if er is not None:
                res["grad_er"][:, h:h + 1] = er.grad
            if ee is not None:
                res["grad_ee"][:, h:h + 1] = ee.grad
    return res


def eval_explicit(sub, slope=0.2, chunk=1 << 17):
    """Forward and the adjoint of SURVEY.md Appendix A.3 written out, fp64, chunked over edges.

    z = el[u] + er[v] + ee[k]; s = leaky_relu(z); softmax over the KEPT in-edges of v
    (models.py:534-544); a~ = a * attn_mul; acc[v] = sum a~ * c_u * ft[u]; out = acc * ds_v.
    g' = gout * ds_v; d_k = <c_u ft[u], g'[v]>; t_v = <acc_v, g'_v>;
    gz_k = a_k (d_k * attn_mul_k - t_v) * (z_k > 0 ? 1 : slope);
    grad_ft[u] = c_u sum_k a~_k g'[v_k]; grad_el[u] = sum_k gz_k; grad_er[v] = sum_k gz_k; grad_ee[k] = gz_k."""
    H, n_w, n_u = sub["H"], sub["n_w"], sub["n_u"]
    e_src, e_dst = sub["e_src"], sub["e_dst"]
    S = e_src.numel()
    ft = sub["ft"]
    D = ft.shape[2]
    f64 = torch.float64
    cs = sub["src_scale"] if sub["src_scale"] is not None else torch.ones(n_u, dtype=f64)
    ds = sub["dst_scale"] if sub["dst_scale"] is not None else torch.ones(n_w, dtype=f64)
    z = sub["el"].index_select(0, e_src)
    if sub["er"] is not None:
        z = z + sub["er"].index_select(0, e_dst)
    if sub["ee"] is not None:
        z = z + sub["ee"]
    s = torch.where(z > 0, z, z * slope)
    kept = torch.ones(S, dtype=torch.bool) if sub["keep"] is None else sub["keep"]
    s = torch.where(kept.view(-1, 1), s, torch.full_like(s, float("-inf")))
    idx = e_dst.view(-1, 1).expand(S, H)
    m = torch.full((n_w, H), float("-inf"), dtype=f64).scatter_reduce(0, idx, s, "amax", include_self=True)
    m_e = m.index_select(0, e_dst)
    p = torch.where(torch.isinf(s), torch.zeros_like(s), torch.exp(s - torch.where(torch.isinf(m_e), torch.zeros_like(m_e), m_e)))
    l = torch.zeros((n_w, H), dtype=f64).index_add_(0, e_dst, p)
    l_e = l.index_select(0, e_dst)
    a = torch.where(l_e > 0, p / torch.where(l_e > 0, l_e, torch.ones_like(l_e)), torch.zeros_like(p))
    am = sub["attn_mul"] if sub["attn_mul"] is not None else torch.ones_like(a)
    at = a * am
    res = {"out": None, "alpha": at}
    acc = torch.zeros((n_w, H, D), dtype=f64)
    have_g = sub["gout"] is not None
    if have_g:
        gp = sub["gout"] * ds.view(-1, 1, 1)
        d = torch.empty((S, H), dtype=f64)
        gft = torch.zeros((n_u, H, D), dtype=f64)
    for lo in range(0, S, chunk):
        hi = min(S, lo + chunk)
        us, vs = e_src[lo:hi], e_dst[lo:hi]
        F = ft.index_select(0, us) * cs.index_select(0, us).view(-1, 1, 1)
        acc.index_add_(0, vs, F * at[lo:hi].unsqueeze(-1))
        if have_g:
            G = gp.index_select(0, vs)
            d[lo:hi] = (F * G).sum(-1)
            gft.index_add_(0, us, G * at[lo:hi].unsqueeze(-1))
    res["out"] = acc * ds.view(-1, 1, 1)
    if have_g:
        t = (acc * gp).sum(-1)
        gs = a * (d * am - t.index_select(0, e_dst))
        gz = gs * torch.where(z > 0, torch.ones_like(z), torch.full_like(z, slope))
        gz = torch.where(kept.view(-1, 1), gz, torch.zeros_like(gz))
        res["grad_ft"] = gft * cs.view(-1, 1, 1)
        res["grad_el"] = torch.zeros((n_u, H), dtype=f64).index_add_(0, e_src, gz)
        res["grad_er"] = None if sub["er"] is None else torch.zeros((n_w, H), dtype=f64).index_add_(0, e_dst, gz)
        res["grad_ee"] = None if sub["ee"] is None else gz
    return res


def adjacent_dst(src, dst, n_dst, U):
    """Destination rows reached by the out-edges of the sources ``U`` (sorted, unique)."""
    dev = src.device
    U = torch.as_tensor(U, dtype=torch.int64, device=dev)
    mark = torch.zeros(int(max(int(src.max().item()) + 1, int(U.max().item()) + 1)), dtype=torch.bool, device=dev)
    mark[U] = True
    return torch.unique(dst[mark.index_select(0, src)])


def row_rel_err(x, ref, floor_frac=1e-3):
    """Row-normalised error: max over rows r of  max|x_r - ref_r| / max(max|ref_r|, floor_frac * max|ref|)
    (rows = leading index; an elementwise relative error is ill-defined at near-zero entries, so each row is
    normalised by its own magnitude, floored at a small fraction of the global one)."""
    x = x.detach().to("cpu", torch.float64)
    ref = ref.detach().to("cpu", torch.float64)
    assert x.shape == ref.shape, (x.shape, ref.shape)
    if ref.numel() == 0:
        return 0.0
    x2, r2 = x.reshape(x.shape[0], -1), ref.reshape(ref.shape[0], -1)
    gmax = r2.abs().max().item()
    if gmax == 0:
        return float((x2 - r2).abs().max().item())
    denom = torch.clamp(r2.abs().amax(1), min=floor_frac * gmax)
    return float(((x2 - r2).abs().amax(1) / denom).max().item())
