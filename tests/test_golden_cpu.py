"""The oracle against the golden vectors recorded from the unmodified reference modules (CPU)."""
import glob
import os

import pytest
import torch

from oracle import modules_ref

_ALL = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.pt")))
GOLDEN = [p for p in _ALL if not os.path.basename(p).startswith(("model_", "gcn_"))]   # single GATConv layers
GOLDEN_GCN = [p for p in _ALL if os.path.basename(p).startswith("gcn_")]               # single GraphConv layers
GOLDEN_MODELS = [p for p in _ALL if os.path.basename(p).startswith("model_")]     # whole GAT models


def load(path):
    return torch.load(path, weights_only=False)


def replay(case):
    """keep set and (E,H) attention multiplier the reference drew, in edge-id order."""
    E = case["src"].numel()
    ctor = case["ctor"]
    H = ctor.get("num_heads", ctor.get("n_heads", 1))
    keep = eids = None
    if case["perm"] is not None:
        keep, eids = modules_ref.keep_from_perm(case["perm"], E, ctor["edge_drop"])
    mul = modules_ref.attn_mul_full(case["attn_mul"], E, H, eids)
    return keep, mul


def oracle_forward(case, dtype=torch.float32):
    ctor, sd = case["ctor"], {k: v.to(dtype) for k, v in case["state_dict"].items()}
    keep, mul = replay(case)
    mul = None if mul is None else mul.to(dtype)
    x = case["x"].to(dtype).clone().requires_grad_(True)
    fe = None if case["fe"] is None else case["fe"].to(dtype).clone().requires_grad_(True)
    n, n_dst = case["n"], case["n_dst"]
    if case["kind"] == "v1":
        y = modules_ref.gatconv_v1(sd, case["src"], case["dst"], n, n_dst, x, num_heads=ctor.get("num_heads", 1),
                                   out_feats=ctor["out_feats"], use_symmetric_norm=ctor.get("use_symmetric_norm", False),
                                   is_block=case["block"], keep=keep, attn_mul=mul)
    else:
        deg = torch.bincount(case["src"], minlength=n).to(dtype).clamp(min=1)
        y = modules_ref.gatconv_v2(sd, case["src"], case["dst"], n, n_dst, x, fe, n_heads=ctor.get("n_heads", 1),
                                   out_feats=ctor["out_feats"], use_symmetric_norm=ctor.get("use_symmetric_norm", False),
                                   is_block=case["block"], deg=deg, keep=keep, attn_mul=mul)
    return x, fe, y


def test_golden_files_present():
    assert len(GOLDEN) >= 12 and len(GOLDEN_MODELS) >= 4


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-3] for p in GOLDEN])
def test_oracle_matches_reference(path):
    case = load(path)
    x, fe, y = oracle_forward(case)
    assert y.shape == case["y"].shape
    assert torch.allclose(y, case["y"], rtol=2e-5, atol=2e-6), float((y - case["y"]).abs().max())
    (y * case["w"]).sum().backward()
    assert torch.allclose(x.grad, case["gx"], rtol=1e-4, atol=1e-5)
    if fe is not None:
        assert torch.allclose(fe.grad, case["gfe"], rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("path", GOLDEN[:4], ids=[os.path.basename(p)[:-3] for p in GOLDEN[:4]])
def test_oracle_fp64_agrees_with_fp32_reference(path):
    """fp64 oracle (the arbiter of the GPU parity tests) vs the fp32 reference run."""
    case = load(path)
    _, _, y = oracle_forward(case, torch.float64)
    assert torch.allclose(y.float(), case["y"], rtol=2e-5, atol=2e-6)
