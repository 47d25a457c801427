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
    # replay the reference's own nn.Dropout draw (the default is the in-kernel Philox stream)
    monkeypatch.setattr(no_sampling.GATConv, "attn_dropout_mode", "exact")
    monkeypatch.setattr(sampled.GATConv, "attn_dropout_mode", "exact")
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


def _build_model(name):
    import torch.nn.functional as F

    from bot_b200.no_sampling import GAT
    from bot_b200.ogbn_products import GAT as ProductsGAT
    from bot_b200.ogbn_proteins import GAT as ProteinsGAT

    if name == "model_v1_gat_bn":
        return GAT(20, 0, 5, 8, 3, 2, F.relu, norm="batch", dropout=0.5, attn_drop=0.1, use_symmetric_norm=True,
                   linear=True, residual=True)
    if name == "model_v1_gat_bias":
        return GAT(20, 0, 5, 8, 2, 3, F.relu, norm="none", non_interactive_attn=True)
    if name == "model_proteins_gat":
        return ProteinsGAT(8, 8, 11, 2, 3, 10, 16, F.relu, 0.25, 0.1, 0.0, 0.1)
    if name == "model_v1_gcn":
        from bot_b200.no_sampling import GCN

        return GCN(20, 5, 16, 3, F.relu, norm="batch", norm_adj="symm", dropout=0.5, residual=True, use_linear=True)
    if name == "model_products_gat":
        return ProductsGAT(9, 0, 7, 2, 2, 6, 0, F.relu, 0.5, 0.1, 0.0, 0.1, allow_zero_in_degree=True, residual=True)
    raise KeyError(name)


from test_golden_cpu import GOLDEN_MODELS  # noqa: E402


@pytest.mark.parametrize("path", GOLDEN_MODELS, ids=[os.path.basename(p)[:-3] for p in GOLDEN_MODELS])
def test_whole_model_matches_reference(cuda, path):
    """The GAT wrappers (callers of GATConv) end to end against the unmodified reference models."""
    import bot_b200

    case = load(path)
    name = os.path.basename(path)[:-3]
    model = _build_model(name).to(cuda)
    model.load_state_dict({k: v.to(cuda) for k, v in case["state_dict"].items()}, strict=True)
    model.eval()
    g = bot_b200.Graph(case["src"].to(cuda), case["dst"].to(cuda), case["n"])
    t = {k: v.to(cuda) for k, v in case["tensors"].items()}
    with torch.no_grad():
        if name.startswith("model_v1"):
            y = model(g, t["feat"])
        else:
            g.srcdata["feat"] = t["feat"]
            if "efeat" in t:
                g.edata["feat"] = t["efeat"]
            y = model(g)
    assert y.shape == case["y"].shape
    assert rel_err(y, case["y"]) <= 5e-5  # several layers deep; the golden run is fp32 itself


from test_golden_cpu import GOLDEN_GCN  # noqa: E402


@pytest.mark.parametrize("path", GOLDEN_GCN, ids=[os.path.basename(p)[:-3] for p in GOLDEN_GCN])
def test_graphconv_matches_reference(cuda, path):
    """GraphConv on the GAT gather kernel (uniform attention) against the unmodified reference layer."""
    import bot_b200
    from bot_b200.no_sampling import GraphConv

    case = load(path)
    conv = GraphConv(**case["ctor"]).to(cuda)
    conv.load_state_dict({k: v.to(cuda) for k, v in case["state_dict"].items()}, strict=True)
    g = bot_b200.Graph(case["src"].to(cuda), case["dst"].to(cuda), case["n"])
    x = case["x"].to(cuda).requires_grad_(True)
    y = conv(g, x)
    assert rel_err(y, case["y"]) <= 3e-5
    (y * case["w"].to(cuda)).sum().backward()
    assert rel_err(x.grad, case["gx"]) <= GRAD_TOL
    for k, p in conv.named_parameters():
        assert rel_err(p.grad, case["gparams"][k]) <= GRAD_TOL, k
