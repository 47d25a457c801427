"""Self-validation of the oracle (CPU): it is 'parity unpinned' against real DGL (not installable offline), so it
is pinned by the golden vectors of the reference's own modules (test_golden_cpu.py) plus the invariants here."""
import numpy as np
import pytest
import torch

from oracle import gat_ref, graph_ref
from util import make_case, oracle_run


def test_hand_computed_case():
    """H = D = 1, two edges into node 0 with logits ln(1), ln(3) after leaky_relu -> alpha = 1/4, 3/4."""
    src = torch.tensor([0, 1, 1])
    dst = torch.tensor([0, 0, 1])
    ft = torch.tensor([[[2.0]], [[10.0]]])
    el = torch.log(torch.tensor([[1.0], [3.0]]))
    out = gat_ref.gat_sparse(src, dst, 2, ft, el)
    assert torch.allclose(out.flatten(), torch.tensor([0.25 * 2 + 0.75 * 10, 10.0]))
    # negative logits go through the 0.2 slope: el = -5 -> -1
    el2 = torch.tensor([[-5.0], [0.0]])
    a0 = torch.exp(torch.tensor(-1.0)) / (torch.exp(torch.tensor(-1.0)) + 1.0)
    out2 = gat_ref.gat_sparse(src, dst, 2, ft, el2)
    assert torch.allclose(out2[0, 0, 0], a0 * 2 + (1 - a0) * 10)


def test_rows_sum_to_one_and_zero_in_degree():
    c = make_case(40, 40, 100, 3, 5, ee=True, seed=1)
    c["ft"] = torch.ones_like(c["ft"])
    out, _ = oracle_run(c, torch.float64)
    indeg = torch.bincount(c["dst"], minlength=40)
    assert torch.allclose(out[indeg > 0], torch.ones_like(out[indeg > 0]))
    assert float(out[indeg == 0].abs().sum()) == 0.0


def test_edge_permutation_invariance():
    c = make_case(50, 50, 400, 2, 6, ee=True, seed=2)
    out, _ = oracle_run(c, torch.float64)
    perm = torch.randperm(400, generator=torch.Generator().manual_seed(0))
    c2 = dict(c, src=c["src"][perm], dst=c["dst"][perm], ee=c["ee"][perm])
    out2, _ = oracle_run(c2, torch.float64)
    assert torch.allclose(out, out2, rtol=1e-12, atol=1e-12)


def test_edge_drop_equals_subgraph():
    """Softmax over the kept edges only == running on the edge-induced subgraph (DGL edge_softmax(eids=...))."""
    c = make_case(50, 50, 400, 2, 6, ee=True, keep_p=0.3, seed=3)
    out, _ = oracle_run(c, torch.float64)
    k = c["keep"]
    c2 = dict(c, src=c["src"][k], dst=c["dst"][k], ee=c["ee"][k], keep=None)
    out2, _ = oracle_run(c2, torch.float64)
    assert torch.allclose(out, out2, rtol=1e-12, atol=1e-12)


def test_gradcheck_fp64():
    c = make_case(12, 12, 40, 2, 3, ee=True, symm=True, seed=4)
    ft, el, er, ee = (c[k].double().requires_grad_(True) for k in ("ft", "el", "er", "ee"))
    cs, ds = c["src_scale"].double(), c["dst_scale"].double()

    def f(ft, el, er, ee):
        return gat_ref.gat_sparse(c["src"], c["dst"], 12, ft, el, er, ee, None, None, 0.2, cs, ds)

    assert torch.autograd.gradcheck(f, (ft, el, er, ee), eps=1e-6, atol=1e-6)


@pytest.mark.parametrize("kw", [dict(ee=True), dict(ee=True, keep_p=0.3), dict(er=False, symm=True, attn_p=0.2, self_loops=True),
                                dict(ee=True, keep_p=0.2, attn_p=0.1, symm=True)])
def test_big_form_equals_materialising_form(kw):
    """The non-materialising form timed as the CPU arm (bench.py) == the parity reference, every flag."""
    c = make_case(80, 80, 900, 3, 8, seed=5, **kw)
    out, g = oracle_run(c, torch.float64)
    bg = gat_ref.BigGraph(c["src"], c["dst"], 80, 80)
    f64 = lambda t: None if t is None else t.double()  # noqa: E731
    o2, saved = gat_ref.gat_sparse_big_forward(bg, f64(c["ft"]), f64(c["el"]), f64(c["er"]), f64(c["ee"]), 0.2, c["keep"],
                                               f64(c["attn_mul"]), f64(c["src_scale"]), f64(c["dst_scale"]))
    gft, gel, ger, gee = gat_ref.gat_sparse_big_backward(bg, saved, c["gout"].double(), 0.2, c["er"] is not None, c["ee"] is not None,
                                                         c["keep"], f64(c["src_scale"]), f64(c["dst_scale"]))
    assert torch.allclose(out, o2, rtol=1e-10, atol=1e-12)
    for got, want in ((gft, g["ft"]), (gel, g["el"]), (ger, g["er"]), (gee, g["ee"])):
        if want is None:
            continue
        assert torch.allclose(got, want, rtol=1e-9, atol=1e-11)


def test_graph_oracle_known_answers():
    src = np.array([2, 0, 1, 0, 2, 2])
    dst = np.array([0, 1, 1, 2, 2, 1])
    f = graph_ref.build_formats(src, dst, 3, 3)
    assert f["in_indptr"].tolist() == [0, 1, 4, 6]
    assert f["in_indices"].tolist() == [2, 0, 1, 2, 0, 2]   # sources, row by row, edge-id order inside a row
    assert f["in_eid"].tolist() == [0, 1, 2, 5, 3, 4]
    assert f["out_indptr"].tolist() == [0, 2, 3, 6]
    assert f["out_indices"].tolist() == [1, 2, 1, 0, 2, 1]
    assert f["out_eid"].tolist() == [1, 3, 2, 0, 4, 5]
    assert f["in_deg"].tolist() == [1, 3, 2] and f["out_deg"].tolist() == [2, 1, 3]
    s, d = graph_ref.to_bidirected(np.array([0, 0, 2]), np.array([1, 1, 2]), 3)
    assert s.tolist() == [0, 1, 2] and d.tolist() == [1, 0, 2]          # dedup, sorted by (src,dst)
    s, d = graph_ref.add_self_loop(*graph_ref.remove_self_loop(s, d), 3)
    assert s.tolist() == [0, 1, 0, 1, 2] and d.tolist() == [1, 0, 0, 1, 2]  # loops appended last, i = 0..N-1
    assert np.allclose(graph_ref.deg_scale(np.array([0, 1, 4]), -0.5), [1.0, 1.0, 0.5])
    assert np.allclose(graph_ref.deg_scale(np.array([0, 1, 4]), 0.5), [1.0, 1.0, 2.0])


@pytest.mark.parametrize("parts", [1, 2, 4, 7])
def test_partition_oracle_properties(parts):
    n, e = 200, 3000
    src, dst = graph_ref.synthetic_coo(n, e, 6, power_law=0.7)
    f = graph_ref.build_formats(src, dst, n, n)
    b = graph_ref.partition_bounds(f["in_indptr"], parts)
    assert b[0] == 0 and b[-1] == n and (np.diff(b) >= 0).all()
    seen = np.zeros(e, dtype=int)
    for r in range(parts):
        loc = graph_ref.partition_local(src, dst, n, b, r)
        seen[loc["edge_gid"]] += 1
        gsrc = np.where(loc["lsrc"] < loc["hi"] - loc["lo"], loc["lsrc"] + loc["lo"],
                        loc["halo_gid"][np.maximum(loc["lsrc"] - (loc["hi"] - loc["lo"]), 0)] if len(loc["halo_gid"]) else 0)
        assert np.array_equal(gsrc, src[loc["edge_gid"]])
        assert np.array_equal(loc["ldst"] + loc["lo"], dst[loc["edge_gid"]])
        assert loc["recv_counts"].sum() == len(loc["halo_gid"])
    assert (seen == 1).all()  # every edge belongs to exactly one rank


def _scalar_loop_gat(c, slope=0.2):
    """The sparse section of GATConv.forward written as plain Python loops over destinations, heads and edges, reading
    the reference line by line (src/no-sampling/models.py:500-555, src/ogbn-proteins/models.py:125-156) and sharing NO
    code with oracle/gat_ref.py (no index_select / scatter / index_add): an independent evaluation of the same formulas
    on tiny graphs.  Returns out (N_d, H, D) as nested float64 numpy."""
    import math

    src, dst = c["src"].tolist(), c["dst"].tolist()
    n_dst, H, D = c["n_dst"], c["H"], c["D"]
    ft, el = c["ft"].double().numpy(), c["el"].double().numpy()
    er = None if c["er"] is None else c["er"].double().numpy()
    ee = None if c["ee"] is None else c["ee"].double().numpy()
    keep = None if c["keep"] is None else c["keep"].tolist()
    am = None if c["attn_mul"] is None else c["attn_mul"].double().numpy()
    cs = None if c["src_scale"] is None else c["src_scale"].double().numpy()
    ds = None if c["dst_scale"] is None else c["dst_scale"].double().numpy()
    out = np.zeros((n_dst, H, D))
    in_edges = [[] for _ in range(n_dst)]
    for k, v in enumerate(dst):
        if keep is None or keep[k]:          # edge_softmax(graph, e[eids], eids=eids): dropped edges take no part
            in_edges[v].append(k)
    for v in range(n_dst):
        for h in range(H):
            logits = []
            for k in in_edges[v]:
                z = el[src[k], h] + (er[v, h] if er is not None else 0.0) + (ee[k, h] if ee is not None else 0.0)
                logits.append(z if z > 0 else slope * z)
            if not logits:
                continue
            m = max(logits)
            w = [math.exp(s - m) for s in logits]
            tot = sum(w)
            for k, wk in zip(in_edges[v], w):
                a = wk / tot
                if am is not None:
                    a *= am[k, h]
                u = src[k]
                scale = cs[u] if cs is not None else 1.0
                for d in range(D):
                    out[v, h, d] += a * scale * ft[u, h, d]
            if ds is not None:
                for d in range(D):
                    out[v, h, d] *= ds[v]
    return out


@pytest.mark.parametrize("kw", [dict(), dict(ee=True), dict(ee=True, keep_p=0.4), dict(er=False, symm=True, attn_p=0.3, self_loops=True),
                                dict(ee=True, keep_p=0.3, attn_p=0.2, symm=True), dict(er=False, keep_p=0.6)])
def test_scalar_loop_restatement(kw):
    """The vectorised oracle against an independent scalar-loop evaluation (forward), and the oracle's autograd
    gradients against central differences of that scalar-loop evaluation (backward) — on graphs with zero in-degree
    rows, parallel edges and rows that lose every edge to edge-drop."""
    n_src = n_dst = 9
    c = make_case(n_src, n_dst, 30, 2, 3, seed=11 + len(kw), **kw)
    ref = _scalar_loop_gat(c)
    out, grads = oracle_run(c, torch.float64)
    assert np.allclose(out.numpy(), ref, rtol=1e-12, atol=1e-12)
    # d<gout, out>/d(input) by central differences of the scalar-loop form, a few entries per input
    gout = c["gout"].double().numpy()
    rng = np.random.default_rng(0)
    for name, key in (("ft", "ft"), ("el", "el"), ("er", "er"), ("ee", "ee")):
        if c[key] is None:
            continue
        g = grads[name].numpy()
        flat = c[key].double().numpy().reshape(-1)
        for idx in rng.choice(flat.size, size=min(6, flat.size), replace=False):
            vals = []
            for sgn in (+1.0, -1.0):
                pert = flat.copy()
                pert[idx] += sgn * 1e-6
                c2 = dict(c)
                c2[key] = torch.from_numpy(pert.reshape(c[key].shape))
                vals.append(float((_scalar_loop_gat(c2) * gout).sum()))
            fd = (vals[0] - vals[1]) / 2e-6
            assert abs(fd - g.reshape(-1)[idx]) <= 1e-6 * max(1.0, abs(fd)), (name, idx, fd, g.reshape(-1)[idx])
