"""The row-subsample oracle (oracle/gat_rows.py) against the materialising restatement (oracle/gat_ref.py), CPU.

The full-size GPU parity tests (tests/test_gpu_fullsize.py) trust ``gat_rows.eval_explicit`` on sub-problems of
the BASELINE-size graphs; here it is pinned (a) to autograd of ``gat_ref.gat_sparse`` on the same sub-problem and
(b) to the full-graph oracle restricted to the sampled rows."""
import pytest
import torch

from oracle import gat_rows
from util import make_case, oracle_run

FLAGS = {
    "plain": dict(),
    "src_only": dict(er=False),
    "ee_keep": dict(ee=True, keep_p=0.3),
    "attn_drop_symm": dict(er=False, symm=True, attn_p=0.2, self_loops=True),
    "everything": dict(ee=True, keep_p=0.2, attn_p=0.1, symm=True),
    "skew": dict(ee=True, power_law=1.0, keep_p=0.1),
}


def _sub(c, W):
    return gat_rows.build_sub(c["src"], c["dst"], c["n_src"], c["n_dst"], W, ft=c["ft"], el=c["el"], er=c["er"], ee=c["ee"],
                              keep=c["keep"], attn_mul=c["attn_mul"], src_scale=c["src_scale"], dst_scale=c["dst_scale"],
                              gout=c["gout"])


@pytest.mark.parametrize("name", list(FLAGS))
def test_explicit_adjoint_equals_autograd(name):
    c = make_case(60, 60, 700, 3, 5, seed=11, **FLAGS[name])
    W = torch.randperm(60, generator=torch.Generator().manual_seed(1))[:25].sort().values
    sub = _sub(c, W)
    a, b = gat_rows.eval_autograd(sub), gat_rows.eval_explicit(sub, chunk=97)
    for k in ("out", "grad_ft", "grad_el", "grad_er", "grad_ee"):
        if a[k] is None:
            assert b[k] is None
            continue
        assert torch.allclose(a[k], b[k], rtol=1e-11, atol=1e-12), k


@pytest.mark.parametrize("name", list(FLAGS))
def test_sub_problem_reproduces_the_full_oracle(name):
    c = make_case(80, 80, 900, 2, 6, seed=12, **FLAGS[name])
    ref_out, ref_g = oracle_run(c, torch.float64)
    U = torch.tensor([3, 17, 40])
    W = torch.unique(torch.cat([gat_rows.adjacent_dst(c["src"], c["dst"], 80, U), torch.tensor([0, 5, 79])]))
    sub = _sub(c, W)
    r = gat_rows.eval_explicit(sub)
    assert torch.allclose(r["out"], ref_out[sub["W"]], rtol=1e-11, atol=1e-12)
    if ref_g["er"] is not None:
        assert torch.allclose(r["grad_er"], ref_g["er"][sub["W"]], rtol=1e-11, atol=1e-12)
    if ref_g["ee"] is not None:
        assert torch.allclose(r["grad_ee"], ref_g["ee"][sub["eid"]], rtol=1e-11, atol=1e-12)
    comp = sub["complete"]
    # every sampled source (with at least one out-edge) is complete in the sub-problem
    have = torch.isin(U, sub["U"][comp])
    assert bool(have[torch.isin(U, c["src"])].all())
    assert torch.allclose(r["grad_ft"][comp], ref_g["ft"][sub["U"][comp]], rtol=1e-11, atol=1e-12)
    assert torch.allclose(r["grad_el"][comp], ref_g["el"][sub["U"][comp]], rtol=1e-11, atol=1e-12)


def test_row_rel_err():
    ref = torch.tensor([[1.0, 2.0], [1e-6, 0.0], [100.0, -50.0]], dtype=torch.float64)
    x = ref.clone()
    x[0, 1] += 2e-3           # row max 2 -> 1e-3
    assert abs(gat_rows.row_rel_err(x, ref) - 1e-3) < 1e-12
    x = ref.clone()
    x[1, 0] += 1e-4           # tiny row: normalised by the floor 1e-3 * 100 = 0.1
    assert abs(gat_rows.row_rel_err(x, ref) - 1e-3) < 1e-9
