"""Partition maps and halo exchange on CPU (gloo, world_size 2 and 3) — host logic of the multi-GPU path.
The oracle (oracle/graph_ref.py, oracle/gat_ref.py) is the checker and stands in for the CUDA compute."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import gat_ref, graph_ref


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _graph(seed=0, n=300, e=4000, power_law=0.0):
    src, dst = graph_ref.synthetic_coo(n, e, seed, power_law=power_law)
    return n, torch.from_numpy(src), torch.from_numpy(dst)


@pytest.mark.parametrize("parts", [1, 2, 3, 8])
@pytest.mark.parametrize("power_law", [0.0, 1.0])
def test_bounds_match_oracle(parts, power_law):
    from bot_b200.partition import partition_bounds

    n, src, dst = _graph(3, 500, 7000, power_law)
    f = graph_ref.build_formats(src.numpy(), dst.numpy(), n, n)
    ref = graph_ref.partition_bounds(f["in_indptr"], parts)
    got = partition_bounds(dst, n, parts).numpy()
    assert np.array_equal(got, ref)
    assert got[0] == 0 and got[-1] == n and (np.diff(got) >= 0).all()


def test_bounds_degenerate():
    from bot_b200.partition import partition_bounds

    # all edges into one row; more parts than rows with edges
    dst = torch.full((100,), 5, dtype=torch.int64)
    b = partition_bounds(dst, 10, 4).numpy()
    f = graph_ref.build_formats(np.zeros(100, dtype=np.int64), dst.numpy(), 10, 10)
    assert np.array_equal(b, graph_ref.partition_bounds(f["in_indptr"], 4))
    # empty graph
    b = partition_bounds(torch.zeros(0, dtype=torch.int64), 7, 3).numpy()
    assert b[0] == 0 and b[-1] == 7


def _worker(rank, world, port, plan, seed, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from bot_b200.partition import PartitionedGraph

        n, src, dst = _graph(seed)
        pg = PartitionedGraph(src, dst, n, plan=plan, build_graph=False)
        H, D = 2, 4
        g = torch.Generator().manual_seed(seed + 100)
        ft = torch.randn(n, H, D, generator=g, dtype=torch.float64)
        el = torch.randn(n, H, generator=g, dtype=torch.float64)
        er = torch.randn(n, H, generator=g, dtype=torch.float64)
        ee = torch.randn(src.numel(), H, generator=g, dtype=torch.float64)
        gout = torch.randn(n, H, D, generator=g, dtype=torch.float64)

        # ---- maps vs the partition oracle (sparse numbering) ----
        ref = graph_ref.partition_local(src.numpy(), dst.numpy(), n, pg.bounds.numpy(), rank)
        assert (pg.lo, pg.hi) == (ref["lo"], ref["hi"])
        assert np.array_equal(pg.edge_gid.numpy(), ref["edge_gid"])
        assert np.array_equal(pg.ldst.numpy(), ref["ldst"])
        assert np.array_equal(pg.halo_gid.numpy(), ref["halo_gid"])
        if pg.plan == "sparse":
            assert np.array_equal(pg.lsrc.numpy(), ref["lsrc"])
            assert pg.recv_counts == ref["recv_counts"].tolist()
            want = np.concatenate(ref["send_lists"]) - ref["lo"] if world > 1 else np.zeros(0, dtype=np.int64)
            assert np.array_equal(pg.send_idx.numpy(), want)

        # ---- forward + backward through the exchange, oracle as the local compute ----
        ft_o = pg.owned_slice(ft).clone().requires_grad_(True)
        el_o = pg.owned_slice(el).clone().requires_grad_(True)
        er_o = pg.owned_slice(er).clone().requires_grad_(True)
        ee_l = pg.local_edges(ee).clone().requires_grad_(True)
        ft_all, el_all = pg.halo_gather(ft_o, el_o)
        assert ft_all.shape[0] == pg.n_src_local
        out = gat_ref.gat_sparse(pg.lsrc, pg.ldst, pg.n_own, ft_all, el_all, er_o, ee_l)
        out.backward(pg.owned_slice(gout))

        # single-process reference on the whole graph
        ftr, elr = ft.clone().requires_grad_(True), el.clone().requires_grad_(True)
        err, eer = er.clone().requires_grad_(True), ee.clone().requires_grad_(True)
        full = gat_ref.gat_sparse(src, dst, n, ftr, elr, err, eer)
        full.backward(gout)
        for got, want in ((out, full[pg.lo:pg.hi]), (ft_o.grad, ftr.grad[pg.lo:pg.hi]), (el_o.grad, elr.grad[pg.lo:pg.hi]),
                          (er_o.grad, err.grad[pg.lo:pg.hi]), (ee_l.grad, eer.grad.index_select(0, pg.edge_gid))):
            assert torch.allclose(got, want.detach(), rtol=1e-10, atol=1e-12)
        ret[rank] = "ok"
    except Exception as ex:  # surface the failure in the parent
        import traceback

        ret[rank] = traceback.format_exc() + repr(ex)
    finally:
        dist.destroy_process_group()


def _fake_gat_fused(graph, ft, el, er=None, ee=None, keep=None, attn_mul=None, src_scale=None, dst_scale=None,
                    slope=0.2, attn_p=0.0, seed=0, hooks=None, edge_order="canonical"):
    """CPU stand-in for bot_b200.functional.gat_fused (oracle math) that honours the Hooks protocol, so that
    PartitionedGraph.gat's overlap logic (async all-gather / early reduce-scatter) can run under gloo."""
    lsrc, ldst, n_dst = graph

    class Fn(torch.autograd.Function):
        @staticmethod
        def forward(ctx, ft, el, er, ee):
            if hooks.head_chunks:
                for i in range(len(hooks.head_chunks)):   # the kernels would be launched range by range
                    hooks.pre_head(i)
            else:
                hooks.pre_kernel()
            with torch.enable_grad():
                ins = [t.detach().requires_grad_(True) for t in (ft, el, er, ee)]
                out = gat_ref.gat_sparse(lsrc, ldst, n_dst, *ins, keep, attn_mul, slope, src_scale, dst_scale)
            ctx.ins, ctx.out = ins, out
            return out.detach()

        @staticmethod
        def backward(ctx, g):
            gft, gel, ger, gee = torch.autograd.grad(ctx.out, ctx.ins, g)
            if hooks.head_chunks:
                for i in range(len(hooks.head_chunks)):
                    hooks.post_src_head(i, gft.contiguous(), gel.contiguous())
            else:
                hooks.post_src(gft.contiguous(), gel.contiguous())
            return gft, gel, ger, gee

    return Fn.apply(ft, el, er, ee)


def _worker_overlap(rank, world, port, seed, ret, pipeline_heads=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import bot_b200.functional as F_
        from bot_b200.partition import PartitionedGraph

        n, src, dst = _graph(seed)
        pg = PartitionedGraph(src, dst, n, plan="dense", build_graph=False)
        pg.local = (pg.lsrc, pg.ldst, pg.n_own)
        pg.pipeline_heads = pipeline_heads
        F_.gat_fused = _fake_gat_fused
        H, D = 2, 4
        g = torch.Generator().manual_seed(seed + 100)
        ft = torch.randn(n, H, D, generator=g, dtype=torch.float64)
        el = torch.randn(n, H, generator=g, dtype=torch.float64)
        er = torch.randn(n, H, generator=g, dtype=torch.float64)
        ee = torch.randn(src.numel(), H, generator=g, dtype=torch.float64)
        gout = torch.randn(n, H, D, generator=g, dtype=torch.float64)
        ft_o = pg.owned_slice(ft).clone().requires_grad_(True)
        el_o = pg.owned_slice(el).clone().requires_grad_(True)
        er_o = pg.owned_slice(er).clone().requires_grad_(True)
        ee_l = pg.local_edges(ee).clone().requires_grad_(True)
        out = pg.gat(ft_o, el_o, er_o, ee_l)
        out.backward(pg.owned_slice(gout))
        ftr, elr = ft.clone().requires_grad_(True), el.clone().requires_grad_(True)
        err, eer = er.clone().requires_grad_(True), ee.clone().requires_grad_(True)
        full = gat_ref.gat_sparse(src, dst, n, ftr, elr, err, eer)
        full.backward(gout)
        for got, want in ((out, full[pg.lo:pg.hi]), (ft_o.grad, ftr.grad[pg.lo:pg.hi]), (el_o.grad, elr.grad[pg.lo:pg.hi]),
                          (er_o.grad, err.grad[pg.lo:pg.hi]), (ee_l.grad, eer.grad.index_select(0, pg.edge_gid))):
            assert torch.allclose(got, want.detach(), rtol=1e-10, atol=1e-12)
        ret[rank] = "ok"
    except Exception as ex:
        import traceback

        ret[rank] = traceback.format_exc() + repr(ex)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,pipeline_heads", [(2, False), (3, False), (2, True), (3, True)])
def test_overlapped_layer_equals_single(world, pipeline_heads):
    """PartitionedGraph.gat (collectives started asynchronously around the kernels; optionally one head range at a
    time) == single-process result."""
    port = _free_port()
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker_overlap, args=(world, port, 11, ret, pipeline_heads), nprocs=world, join=True)
        assert all(ret.get(r) == "ok" for r in range(world)), dict(ret)


@pytest.mark.parametrize("world,plan", [(2, "sparse"), (2, "dense"), (3, "sparse"), (3, "dense"), (2, "auto")])
def test_partitioned_equals_single(world, plan):
    port = _free_port()
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(world, port, plan, 7, ret), nprocs=world, join=True)
        assert all(ret.get(r) == "ok" for r in range(world)), dict(ret)
