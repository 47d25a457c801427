"""The drop-in modules (CUDA, through the C ABI) against golden vectors recorded from the unmodified
reference modules: same state_dict, same inputs, same random draws — `-m gpu`."""
import glob
import os

import pytest
import torch

from test_golden_cpu import GOLDEN, load, replay
from util import FWD_TOL, GRAD_TOL, rel_err

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-3] for p in GOLDEN])
def test_module_matches_reference(cuda, path, monkeypatch):
    import bot_b200
    from bot_b200 import no_sampling, sampled

    case = load(path)
    ctor = case["ctor"]
    torch.manual_seed(0)
    cls = no_sampling.GATConv if case["kind"] == "v1" else sampled.GATConv
    conv = cls(**ctor).to(cuda)
    missing = conv.load_state_dict({k: v.to(cuda) for k, v in case["state_dict"].items()}, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    conv.train(case["train"])

    keep, mul = replay(case)
    if keep is not None:
        eids = torch.nonzero(keep).flatten()
        monkeypatch.setattr(no_sampling, "draw_edge_keep", lambda E, p, dev: (keep.to(dev).to(torch.uint8), eids.to(dev)))
        monkeypatch.setattr(sampled, "draw_edge_keep", lambda E, p, dev: (keep.to(dev).to(torch.uint8), eids.to(dev)))
    if mul is not None:
        monkeypatch.setattr(no_sampling, "draw_attn_mul", lambda mod, E, H, dev, e=None: mul.to(dev))
        monkeypatch.setattr(sampled, "draw_attn_mul", lambda mod, E, H, dev, e=None: mul.to(dev))

    g = bot_b200.Graph(case["src"].to(cuda), case["dst"].to(cuda), case["n"], case["n_dst"], is_block=case["block"])
    if case["kind"] == "v2" and ctor.get("use_symmetric_norm"):
        deg = g.out_degrees().float().clamp(min=1)
        g.srcdata["deg"], g.dstdata["deg"] = deg, deg[: case["n_dst"]]
    x = case["x"].to(cuda).requires_grad_(True)
    fe = None if case["fe"] is None else case["fe"].to(cuda).requires_grad_(True)
    y = conv(g, x) if case["kind"] == "v1" else conv(g, x, fe)
    assert y.shape == case["y"].shape
    assert rel_err(y, case["y"]) <= FWD_TOL * 3  # the golden run itself is fp32
    (y * case["w"].to(cuda)).sum().backward()
    assert rel_err(x.grad, case["gx"]) <= GRAD_TOL
    if fe is not None:
        assert rel_err(fe.grad, case["gfe"]) <= GRAD_TOL
    for k, p in conv.named_parameters():
        if k in case["gparams"]:
            assert rel_err(p.grad, case["gparams"][k]) <= GRAD_TOL, k
