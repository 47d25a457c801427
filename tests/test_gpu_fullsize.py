"""VALUE parity at the real BASELINE sizes — `-m gpu`.

The 16 small cases of test_gpu_parity.py cannot see bugs that need > 2^31-byte offsets, 100 K-block grids,
multi-slab column parts, 800 K-edge rows or the per-rank block of the 8-GPU partition.  Here the layer runs ONCE
at each full size on the GPU and a random sample of destination rows (forward, grad_er, grad_ee) and of source
rows (grad_ft, grad_el) is compared with the fp64 row-subsample oracle (oracle/gat_rows.py: the same math as
oracle/gat_ref.py on all in-edges of the sampled rows; pinned to it by tests/test_gat_rows_cpu.py).

Tolerances (BASELINE.json north_star): forward <= 1e-5, gradients <= 1e-4, measured both norm-relative
(max|x-ref| / max|ref| over the sample) and row-normalised (``gat_rows.row_rel_err``: every row — one node's (H, D)
block, one edge's H logit gradients — by its OWN max, floored at a fraction of the global max: 1e-3 for the forward,
1e-2 for the gradients.  The floor is what fp32 allows: a logit gradient is alpha * (d - t) with d, t dot products
of O(sqrt(D)) magnitude, so its absolute rounding error is ~1e-6 * alpha whatever the size of the difference; rows
where d ~ t are small against the global max and cannot be resolved to 1e-4 of THEMSELVES by any fp32 evaluation,
the reference's included)."""
import numpy as np
import pytest
import torch

from oracle import gat_rows
from util import FWD_TOL, GRAD_TOL, philox_attn_mul, rel_err

pytestmark = pytest.mark.gpu
SLOPE = 0.2


def _synth(n, e, dev, seed, power_law=0.0):
    g = torch.Generator(device=dev).manual_seed(seed)
    src = torch.randint(0, n, (e,), device=dev, generator=g)
    if power_law > 0:
        w = torch.arange(1, n + 1, device=dev, dtype=torch.float64) ** (-power_law)
        cdf = torch.cumsum(w / w.sum(), 0)
        dst = torch.searchsorted(cdf, torch.rand(e, device=dev, dtype=torch.float64, generator=g)).clamp(max=n - 1)
    else:
        dst = torch.randint(0, n, (e,), device=dev, generator=g)
    return src, dst


def _check(name, dev, graph, src, dst, n_src, n_dst, H, D, *, er, ee, edge_drop, attn_p, symm, n_v, n_u, extra_rows=(), seed=0,
           canonical=False):
    """One full-size layer step on the GPU vs the row-subsample oracle.  Returns the error dict.
    ``canonical``: the per-edge operands are given in the graph's canonical order (the production path: no
    permutation pass); the oracle then sees the COO in that same order."""
    from bot_b200 import functional
    from bot_b200.functional import gat_fused

    E = src.numel()
    if canonical and graph.edge_perm() is not None:
        src, dst = src[graph.edge_perm()], dst[graph.edge_perm()]
    gen = torch.Generator(device=dev).manual_seed(seed + 100)
    ft = torch.randn(n_src, H, D, device=dev, generator=gen).requires_grad_(True)
    el = torch.randn(n_src, H, device=dev, generator=gen).requires_grad_(True)
    er_t = torch.randn(n_dst, H, device=dev, generator=gen).requires_grad_(True) if er else None
    ee_t = torch.randn(E, functional.pad_heads(H), device=dev, generator=gen).requires_grad_(True) if ee else None
    gout = torch.randn(n_dst, H, D, device=dev, generator=gen)
    keep = functional.edge_drop_keep(E, int(E * edge_drop), 1234 + seed, dev) if edge_drop > 0 else None
    cs = graph.deg_scale("out", -0.5) if symm else None
    ds = graph.deg_scale("in", 0.5) if symm else None
    pseed = 987654321 + seed
    out = gat_fused(graph, ft, el, er_t, ee_t, keep, None, cs, ds, SLOPE, attn_p, pseed,
                    edge_order="canonical" if canonical else "eid")
    out.backward(gout)
    torch.cuda.synchronize()
    assert torch.isfinite(out).all() and torch.isfinite(ft.grad).all()

    # sampled rows: n_v destinations (+ the rows the caller wants, e.g. the heaviest), n_u sources with all their out-edges
    rs = np.random.default_rng(seed)
    V = torch.from_numpy(rs.choice(n_dst, size=min(n_v, n_dst), replace=False)).to(dev)
    has_out = torch.bincount(src, minlength=n_src) > 0
    cand = torch.nonzero(has_out).flatten()
    U = cand[torch.from_numpy(rs.choice(cand.numel(), size=min(n_u, cand.numel()), replace=False)).to(dev)]
    W = torch.unique(torch.cat([V, gat_rows.adjacent_dst(src, dst, n_dst, U),
                                torch.as_tensor(list(extra_rows), dtype=torch.int64, device=dev)]))
    if symm:   # degree tables for the oracle come from the COO, not from libbotgat
        cs_ref = torch.bincount(src, minlength=n_src).float().clamp(min=1).pow(-0.5)
        ds_ref = torch.bincount(dst, minlength=n_dst).float().clamp(min=1).pow(0.5)
        assert torch.equal(cs_ref, cs) and torch.equal(ds_ref, ds)
    sub = gat_rows.build_sub(src, dst, n_src, n_dst, W, ft=ft, el=el, er=er_t, ee=ee_t, keep=keep, src_scale=cs, dst_scale=ds,
                             gout=gout)
    if attn_p > 0:   # the in-kernel stream is keyed on the graph's canonical edge number
        cid = graph.canonical_edge_ids()
        ids = sub["eid"] if (cid is None or canonical) else cid[sub["eid"].to(dev)].cpu()
        sub["attn_mul"] = philox_attn_mul(pseed, 0, H, attn_p, eids=ids.numpy()).double()
    ref = gat_rows.eval_explicit(sub, SLOPE)
    Wd, Ud, comp = sub["W"].to(dev), sub["U"].to(dev), sub["complete"]
    assert int(comp.sum()) >= min(n_u, int(cand.numel())), "sampled sources must be complete in the sub-problem"
    got = {"out": out.detach()[Wd], "grad_ft": ft.grad[Ud][comp.to(dev)], "grad_el": el.grad[Ud][comp.to(dev)]}
    want = {"out": ref["out"], "grad_ft": ref["grad_ft"][comp], "grad_el": ref["grad_el"][comp]}
    if er:
        got["grad_er"], want["grad_er"] = er_t.grad[Wd], ref["grad_er"]
    if ee:
        got["grad_ee"], want["grad_ee"] = ee_t.grad[sub["eid"].to(dev)][:, :H], ref["grad_ee"]
        assert float(ee_t.grad[:, H:].abs().max()) == 0.0          # padding columns of the records receive zeros
    errs = {}
    for k in got:
        tol = FWD_TOL if k == "out" else GRAD_TOL
        errs[k] = (rel_err(got[k], want[k]), gat_rows.row_rel_err(got[k], want[k], 1e-3 if k == "out" else 1e-2))
    print(f"\n[fullsize {name}] E={E} rows checked: dst={W.numel()} src={int(comp.sum())} sub-edges={sub['e_src'].numel()} "
          + " ".join(f"{k}={a:.1e}/{b:.1e}" for k, (a, b) in errs.items()))
    for k, (a, b) in errs.items():
        tol = FWD_TOL if k == "out" else GRAD_TOL
        assert a <= tol, f"{name}: {k} norm-relative error {a:.3e} > {tol}"
        assert b <= tol, f"{name}: {k} row-normalised error {b:.3e} > {tol}"
    if keep is not None and ee:   # dropped edges: exact zeros
        dropped = torch.nonzero(keep == 0).flatten()[:100000]
        assert float(ee_t.grad[dropped].abs().max()) == 0.0
    return errs


def test_proteins_full_size_values(cuda):
    """BASELINE config 4 (north star): N=132,534, E=39,561,252, H=6, D=80, attn_dst + edge logits, edge_drop 0.1
    (src/ogbn-proteins/models.py:125-156)."""
    import bot_b200

    n, e = 132534, 39561252
    src, dst = _synth(n, e, cuda, 0)
    g = bot_b200.Graph(src, dst, n)
    _check("proteins", cuda, g, src, dst, n, n, 6, 80, er=True, ee=True, edge_drop=0.1, attn_p=0.0, symm=False, n_v=1000, n_u=16)
    # the production path: operands already in canonical order, blocked in <-> out transposes in the backward
    assert g._info.in_eid_identity == 1 and g._info.tiles_src * g._info.tiles_dst > 1
    _check("proteins-canonical", cuda, g, src, dst, n, n, 6, 80, er=True, ee=True, edge_drop=0.1, attn_p=0.2, symm=False, n_v=1000,
           n_u=16, seed=7, canonical=True)


def test_proteins_skewed_full_size_values(cuda):
    """Same shape, destinations ~ rank^-0.8 (hottest row ~800 K in-edges: the row-splitting path at scale)."""
    import bot_b200

    n, e = 132534, 39561252
    src, dst = _synth(n, e, cuda, 0, power_law=0.8)
    g = bot_b200.Graph(src, dst, n)
    assert g._ensure() and g._info.n_slots_in > 0
    _check("proteins-skew", cuda, g, src, dst, n, n, 6, 80, er=True, ee=True, edge_drop=0.1, attn_p=0.0, symm=False, n_v=300,
           n_u=2, extra_rows=(0, 1, 2, 3, n - 1), seed=1)


def test_reddit_full_size_values(cuda):
    """BASELINE config 3: N=232,965, E=114,615,892, H=4, D=64, source-only logits, symmetric norm, attention dropout
    0.1 drawn in-kernel (src/no-sampling/models.py:500-555, run.py:978)."""
    import bot_b200

    n, e = 232965, 114615892
    src, dst = _synth(n, e, cuda, 2)
    g = bot_b200.Graph(src, dst, n)
    _check("reddit", cuda, g, src, dst, n, n, 4, 64, er=False, ee=False, edge_drop=0.0, attn_p=0.1, symm=True, n_v=400, n_u=6, seed=2)


def test_products_full_size_values(cuda):
    """BASELINE config 5: N=2,449,029, E=61,859,140, H=4, D=120, attn_dst, edge_drop 0.1
    (src/ogbn-products/models.py:211-223): one head slab (1.18 GB) exceeds L2 -> column parts."""
    import bot_b200

    n, e = 2449029, 61859140
    src, dst = _synth(n, e, cuda, 3)
    g = bot_b200.Graph(src, dst, n)
    _check("products", cuda, g, src, dst, n, n, 4, 120, er=True, ee=False, edge_drop=0.1, attn_p=0.0, symm=False, n_v=3000, n_u=300,
           seed=3)


def test_arxiv_full_size_values(cuda):
    """BASELINE config 2 at the real preprocessed size: 1,166,243 raw edges -> to_bidirected -> self loops
    (src/no-sampling/run.py:133-148), H=3, D=250 (169 MB head slabs), symmetric norm, attention dropout 0.1."""
    import bot_b200

    n, e = 169343, 1166243
    src, dst = _synth(n, e, cuda, 4)
    g = bot_b200.Graph(src, dst, n).to_bidirected().remove_self_loop().add_self_loop()
    s2, d2 = g.edges()
    assert 2.2e6 < s2.numel() < 2.6e6
    _check("arxiv", cuda, g, s2, d2, n, n, 3, 250, er=False, ee=False, edge_drop=0.0, attn_p=0.1, symm=True, n_v=3000, n_u=500, seed=4)
    # last layer of the same model: H=1, D=40 (160-byte rows)
    _check("arxiv-last", cuda, g, s2, d2, n, n, 1, 40, er=False, ee=False, edge_drop=0.0, attn_p=0.1, symm=True, n_v=3000, n_u=500,
           seed=5)


def test_rank_block_of_8_full_size_values(cuda):
    """The per-rank block of the 8-GPU partition at the proteins shape (n_dst ~ N/8 rows, n_src = 8 * max_own, ~37
    local out-edges per source: the group-per-row backward) — rank 3's local problem, run in this one process."""
    from bot_b200 import partition

    n, e = 132534, 39561252
    src, dst = _synth(n, e, cuda, 0)
    part = partition.PartitionedGraph(src, dst, n, world=8, rank=3, plan="dense")
    del src, dst
    g = part.local
    _check("proteins-rank3of8", cuda, g, part.lsrc, part.ldst, part.n_src_local, part.n_own, 6, 80, er=True, ee=True, edge_drop=0.1,
           attn_p=0.0, symm=False, n_v=500, n_u=32, seed=6, canonical=True)
