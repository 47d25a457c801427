"""Multi-GPU path on real devices — `-m gpu`; the multi-process tests need >= 2 GPUs (`gpurun --gpus 2`) and are
skipped on a single-GPU box (the CPU suite covers the same logic under gloo, tests/test_partition_cpu.py)."""
import ctypes as C
import os
import socket

import pytest
import torch

from oracle import gat_ref
from util import FWD_TOL, make_case, rel_err

pytestmark = pytest.mark.gpu


def test_halo_kernels_with_local_peers(cuda):
    """botgat_halo_pull / botgat_halo_pull_reduce with every "peer" buffer on this GPU: exact copies / fixed-order sums,
    column ranges, unaligned widths (scalar path), every specialised world size."""
    from bot_b200 import _lib
    from bot_b200.graph import _stream

    lib = _lib.load()
    torch.manual_seed(0)
    for world, rows, P, col0, width in ((3, 1000, 96, 0, 96), (4, 777, 512, 80, 160), (2, 50, 40, 3, 7), (8, 1200, 512, 480, 6),
                                        (2, 5000, 128, 32, 64), (8, 900, 64, 0, 64), (5, 300, 64, 16, 32)):
        tables = [torch.randn(world * rows, P, device=cuda) for _ in range(world)]
        ptrs = (C.c_void_p * world)(*[t.data_ptr() for t in tables])
        for me in (0, world - 1):
            before = tables[me].clone()
            _lib.check(lib.botgat_halo_pull(world, me, ptrs, rows, P, col0, width, 8, _stream()), "pull")
            want = before.clone()
            for r in range(world):
                if r != me:
                    want[r * rows:(r + 1) * rows, col0:col0 + width] = tables[r][r * rows:(r + 1) * rows, col0:col0 + width]
            assert torch.equal(tables[me], want)
            out = torch.full((rows, P + 8), 5.0, device=cuda)
            _lib.check(lib.botgat_halo_pull_reduce(world, me, ptrs, rows, P, col0, width, out.data_ptr(), P + 8, 8, _stream()), "reduce")
            acc = tables[0][me * rows:(me + 1) * rows, col0:col0 + width].clone()
            for r in range(1, world):
                acc += tables[r][me * rows:(me + 1) * rows, col0:col0 + width]      # the same fixed order
            want = torch.full_like(out, 5.0)
            want[:, col0:col0 + width] = acc
            assert torch.equal(out, want)
    with pytest.raises(RuntimeError, match="world"):
        _lib.check(lib.botgat_halo_pull(17, 0, ptrs, 1, 8, 0, 8, 0, _stream()), "pull")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from bot_b200.partition import PartitionedGraph

        c = make_case(1500, 1500, 90000, 6, 16, ee=True, keep_p=0.2, seed=5)
        out_ref, g_ref = None, None
        ft, el, er, ee = (c[k].double().requires_grad_(True) for k in ("ft", "el", "er", "ee"))
        full = gat_ref.gat_sparse(c["src"], c["dst"], 1500, ft, el, er, ee, c["keep"])
        full.backward(c["gout"].double())
        src, dst = c["src"].to(dev), c["dst"].to(dev)
        results = {}
        for exchange, chunks in (("nccl", 0), ("p2p", 1), ("p2p", 3), ("p2p", 6)):
            pg = PartitionedGraph(src, dst, 1500, plan="dense")
            pg.exchange, pg.halo_chunks = exchange, chunks
            ft_o = pg.owned_slice(c["ft"].to(dev)).clone().requires_grad_(True)
            el_o = pg.owned_slice(c["el"].to(dev)).clone().requires_grad_(True)
            er_o = pg.owned_slice(c["er"].to(dev)).clone().requires_grad_(True)
            ee_l = pg.local_edges(c["ee"].to(dev)).clone().requires_grad_(True)
            keep_l = pg.local_edges(c["keep"].to(dev))
            for _ in range(2):      # twice: the second step reuses the exchange buffers
                for t in (ft_o, el_o, er_o, ee_l):
                    t.grad = None
                out = pg.gat(ft_o, el_o, er_o, ee_l, keep_l)
                out.backward(pg.owned_slice(c["gout"].to(dev)))
            torch.cuda.synchronize()
            lo, hi = pg.lo, pg.hi
            assert rel_err(out, full[lo:hi]) <= FWD_TOL, (exchange, chunks)
            for got, want in ((ft_o.grad, ft.grad[lo:hi]), (el_o.grad, el.grad[lo:hi]), (er_o.grad, er.grad[lo:hi]),
                              (ee_l.grad, ee.grad.index_select(0, pg.edge_gid.cpu()))):
                assert rel_err(got, want) <= 1e-4, (exchange, chunks)
            results[(exchange, chunks)] = (out.clone(), ft_o.grad.clone(), el_o.grad.clone())
            if exchange == "p2p" and chunks == 3:
                # the producer writes its rows straight into the exchange buffers: no copy inside the layer
                hx = pg.halo_buffers(6, 16)
                hx.h_table.barrier(channel=0)
                with torch.no_grad():
                    hx.own_ft.copy_(ft_o)
                    hx.own_el.copy_(el_o)
                f2, e2 = hx.own_ft.detach().requires_grad_(True), hx.own_el.detach().requires_grad_(True)
                out2 = pg.gat(f2, e2, er_o, ee_l, keep_l)
                out2.backward(pg.owned_slice(c["gout"].to(dev)))
                assert torch.equal(out2, out) and torch.equal(f2.grad, ft_o.grad) and torch.equal(e2.grad, el_o.grad)
        # the peer-memory exchange moves the same numbers whatever the chunking; its fixed-order reduction makes the
        # gradients independent of it bit for bit
        a, b = results[("p2p", 1)], results[("p2p", 6)]
        if os.environ.get("BOTGAT_ROWWISE") == "1":
            # the all-heads-per-row kernels give a head range of another width another lane geometry (summation order)
            assert all(rel_err(y, x) <= 2e-6 for x, y in zip(a, b))
        else:
            assert all(torch.equal(x, y) for x, y in zip(a, b))
        ret[rank] = "ok"
    except Exception as ex:
        import traceback

        ret[rank] = traceback.format_exc() + repr(ex)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_p2p_halo_exchange_matches_single_gpu(world):
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = _free_port()
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert all(ret.get(r) == "ok" for r in range(world)), dict(ret)
