"""Device neighbour sampling / block construction (bot_b200.sampling) — `-m gpu`.
Integer work: bit-exact against the numpy restatement in tests/util.py."""
import numpy as np
import pytest
import torch

from oracle import graph_ref
from util import FWD_TOL, rel_err, sample_neighbors_ref, to_block_ref

pytestmark = pytest.mark.gpu


def _graph(cuda, n=600, e=30000, seed=0, power_law=0.9):
    import bot_b200

    src, dst = graph_ref.synthetic_coo(n, e, seed, power_law=power_law)
    g = bot_b200.Graph(torch.from_numpy(src).to(cuda), torch.from_numpy(dst).to(cuda), n)
    ref = graph_ref.build_formats(src, dst, n, n, canonical=True)   # the sampler walks the graph's own (canonical) rows
    return g, ref


@pytest.mark.parametrize("fanout", [1, 5, 32, 100, 256, -1])
def test_sample_neighbors_bit_exact(cuda, fanout):
    from bot_b200 import sampling

    g, ref = _graph(cuda)
    rng = np.random.default_rng(fanout + 7)
    seeds = rng.permutation(600)[:257].astype(np.int64)
    src, dst, eid, off = sampling.sample_neighbors(g, torch.from_numpy(seeds).to(cuda), fanout, seed=1234 + fanout)
    rs, rd, re, ro = sample_neighbors_ref(ref["in_indptr"], ref["in_indices"], ref["in_eid"], seeds, fanout, 1234 + fanout)
    assert np.array_equal(off.cpu().numpy(), ro)
    assert np.array_equal(src.cpu().numpy(), rs) and np.array_equal(dst.cpu().numpy(), rd)
    assert np.array_equal(eid.cpu().numpy(), re)
    deg = ref["in_deg"][seeds]
    assert np.array_equal(np.diff(ro), deg if fanout <= 0 else np.minimum(deg, fanout))
    assert deg.max() > 2 * max(fanout, 1) or fanout >= 256 or fanout <= 0        # the draw path is exercised
    # every pick is an in-edge of its seed, no edge twice
    e_src, e_dst = g.edges()
    assert torch.equal(e_src[eid], src) and torch.equal(e_dst[eid], torch.from_numpy(seeds).to(cuda)[dst])
    assert eid.unique().numel() == eid.numel()


def test_sample_neighbors_is_uniform(cuda):
    """One row of 50 in-edges, 10 (direct draw) and 40 (complement draw) wanted: every edge equally likely."""
    import bot_b200
    from bot_b200 import sampling

    d = 50
    g = bot_b200.Graph(torch.arange(d, device=cuda), torch.zeros(d, dtype=torch.int64, device=cuda), d)
    seeds = torch.zeros(1, dtype=torch.int64, device=cuda)
    for k in (10, 40):
        draws = 3000
        freq = torch.zeros(d, device=cuda)
        for s in range(draws):
            _, _, eid, _ = sampling.sample_neighbors(g, seeds, k, seed=s)
            assert eid.numel() == k
            freq[eid] += 1
        p = k / d
        z = (freq / draws - p) / (p * (1 - p) / draws) ** 0.5
        assert float(z.abs().max()) < 4.5 and abs(float(z.std()) - 1.0) < 0.35


def test_block_compact_bit_exact(cuda):
    from bot_b200 import sampling

    g, ref = _graph(cuda, n=500, e=8000, seed=3)
    seeds = np.random.default_rng(1).permutation(500)[:64].astype(np.int64)
    tseeds = torch.from_numpy(seeds).to(cuda)
    src, dst, eid, _ = sampling.sample_neighbors(g, tseeds, 6, seed=5)
    blk = sampling.to_block(g, tseeds, src, dst, eid)
    nodes, local = to_block_ref(500, seeds, src.cpu().numpy())
    assert blk.is_block and blk.number_of_dst_nodes() == 64 and blk.number_of_src_nodes() == nodes.size
    assert np.array_equal(blk.srcdata[sampling.NID].cpu().numpy(), nodes)
    assert np.array_equal(blk.dstdata[sampling.NID].cpu().numpy(), seeds)
    bs, bd = blk.edges()
    assert np.array_equal(bs.cpu().numpy(), local) and torch.equal(bd, dst)
    assert torch.equal(blk.srcdata[sampling.NID][bs], src)
    # no sampled edges at all: the block is the seed set
    empty = torch.empty(0, dtype=torch.int64, device=cuda)
    blk0 = sampling.to_block(g, tseeds, empty, empty, empty)
    assert blk0.number_of_src_nodes() == 64 and blk0.number_of_edges() == 0


def test_multilayer_blocks_and_lazy_features(cuda):
    from bot_b200 import sampling

    g, _ = _graph(cuda, n=800, e=20000, seed=5)
    g.ndata["feat"] = torch.randn(800, 7, device=cuda)
    g.ndata["deg"] = g.out_degrees().float().clamp(min=1)
    g.ndata["labels"] = torch.randint(0, 2, (800, 3), device=cuda)
    g.edata["feat"] = torch.randn(20000, 4, device=cuda)
    seeds = torch.randperm(800, device=cuda)[:50]
    blocks = sampling.MultiLayerNeighborSampler([4, 3, 2]).sample_blocks(g, seeds, seed=11)
    assert len(blocks) == 3 and torch.equal(blocks[-1].dstdata[sampling.NID], seeds)
    for a, b in zip(blocks[:-1], blocks[1:]):
        assert torch.equal(a.dstdata[sampling.NID], b.srcdata[sampling.NID])
    for b, f in zip(blocks, [4, 3, 2]):
        n_dst = b.number_of_dst_nodes()
        assert torch.equal(b.srcdata[sampling.NID][:n_dst], b.dstdata[sampling.NID])     # dst = prefix of src
        assert int(b.in_degrees().max()) <= f
        assert torch.equal(b.srcdata["feat"], g.ndata["feat"][b.srcdata[sampling.NID]])
        assert torch.equal(b.dstdata["deg"], g.ndata["deg"][b.dstdata[sampling.NID]])
        assert torch.equal(b.edata["feat"], g.edata["feat"][b.edata[sampling.EID]])
        ps, pd = g.edges()
        bs, bd = b.edges()
        assert torch.equal(ps[b.edata[sampling.EID]], b.srcdata[sampling.NID][bs])
        assert torch.equal(pd[b.edata[sampling.EID]], b.dstdata[sampling.NID][bd])
    assert torch.equal(blocks[-1].dstdata["labels"], g.ndata["labels"][seeds])
    again = sampling.MultiLayerNeighborSampler([4, 3, 2]).sample_blocks(g, seeds, seed=11)
    assert all(torch.equal(x.edata[sampling.EID], y.edata[sampling.EID]) for x, y in zip(blocks, again))


def test_full_neighbor_blocks_equal_full_graph(cuda):
    """MultiLayerFullNeighborSampler over all nodes: the proteins model on the blocks == on the whole graph."""
    import torch.nn.functional as F

    from bot_b200 import sampling
    from bot_b200.ogbn_proteins import GAT

    g, _ = _graph(cuda, n=300, e=6000, seed=8, power_law=0.0)
    g = g.remove_self_loop().add_self_loop()
    g.ndata["feat"] = torch.randn(300, 8, device=cuda)
    g.edata["feat"] = torch.rand(g.number_of_edges(), 8, device=cuda)
    torch.manual_seed(0)
    model = GAT(8, 8, 5, 2, 3, 16, 16, F.relu, 0.0, 0.0, 0.0, 0.0).to(cuda).eval()
    want = model(g)
    blocks = sampling.MultiLayerFullNeighborSampler(2).sample_blocks(g, torch.arange(300, device=cuda))
    assert all(b.number_of_src_nodes() == 300 and b.number_of_edges() == g.number_of_edges() for b in blocks)
    got = model(blocks)
    assert rel_err(got, want) <= FWD_TOL


def test_node_dataloader(cuda):
    from bot_b200 import sampling

    g, _ = _graph(cuda, n=400, e=9000, seed=9)
    g.ndata["feat"] = torch.randn(400, 5, device=cuda)
    nids = torch.arange(0, 400, 2, device=cuda)
    batches = [list(range(0, 70)), list(range(70, 200))]
    loader = sampling.NodeDataLoader(g, nids, sampling.MultiLayerNeighborSampler([3, 3]), batch_sampler=batches, num_workers=10)
    seen = []
    for input_nodes, output_nodes, blocks in loader:
        assert torch.equal(input_nodes, blocks[0].srcdata[sampling.NID])
        assert torch.equal(output_nodes, blocks[-1].dstdata[sampling.NID])
        assert blocks[0].srcdata["feat"].shape == (input_nodes.numel(), 5)
        seen.append(output_nodes)
    assert len(loader) == 2 and torch.equal(torch.cat(seen), nids)
    loader2 = sampling.NodeDataLoader(g, nids, sampling.MultiLayerNeighborSampler([2]), batch_size=64, shuffle=True)
    got = torch.cat([o for _, o, _ in loader2])
    assert len(loader2) == 4 and torch.equal(got.sort().values, nids)
