"""Neighbour sampling and mini-batch blocks on the device.

Mirrors the part of ``dgl.dataloading`` the reference's sampled training uses
(src/ogbn-proteins/gat.py:177-201, src/ogbn-products/gat.py:202-233):

    sampler = MultiLayerNeighborSampler([32] * n_layers)
    loader = NodeDataLoader(graph, train_idx, sampler, batch_sampler=..., num_workers=10)
    for input_nodes, output_nodes, subgraphs in loader:
        pred = model(subgraphs)          # reads srcdata["feat"], edata["feat"], dstdata["labels"], *data["deg"]

The reference samples in 10 CPU worker processes from ``graph.cpu()`` and copies every block to the GPU
(``b.to(device)``, gat.py:109); here the parent graph, its features and the sampler all live on the GPU: a layer's
frontier is drawn by ``botgat_sample_neighbors`` (one warp per seed, uniform without replacement), relabelled by
``botgat_block_compact`` and turned into a :class:`bot_b200.Graph` block whose ``srcdata`` / ``dstdata`` /
``edata`` gather the parent's rows on first access.  ``num_workers`` is accepted and ignored.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .graph import Graph, _stream

NID = "_ID"   # dgl.NID / dgl.EID
EID = "_ID"


class _LazyFrame(dict):
    """Feature dict of a block: a missing key is gathered from the parent's frame by the block's ids."""

    def __init__(self, parent_frame, ids):
        super().__init__()
        self._parent, self._ids = parent_frame, ids
        self[NID] = ids

    def __missing__(self, key):
        value = self._parent[key].index_select(0, self._ids)
        self[key] = value
        return value

    def __contains__(self, key):
        return dict.__contains__(self, key) or key in self._parent

    def keys(self):
        return list(dict.fromkeys(list(dict.keys(self)) + list(self._parent.keys())))

    def canonical(self, key):
        """A block's edges are emitted destination by destination: its own order IS its canonical order."""
        return self[key]


def sample_neighbors(g: Graph, seeds, fanout, seed=0):
    """Uniform sample without replacement of ``fanout`` in-edges per seed (all of them for in-degree <= fanout or
    fanout <= 0), as ``dgl.sampling.sample_neighbors(g, seeds, fanout)``.  Returns (src ids in ``g``, seed position
    of the destination, edge ids in ``g``, offsets (n_seeds + 1)); a seed's picks are contiguous."""
    lib, h = _lib.load(), g._ensure()
    dev = g.device
    seeds = torch.as_tensor(seeds, dtype=torch.int64, device=dev).contiguous()
    n = seeds.numel()
    offsets = torch.empty(n + 1, dtype=torch.int64, device=dev)
    ws = torch.empty(lib.botgat_sample_workspace_bytes(n), dtype=torch.uint8, device=dev)
    total = C.c_int64()
    with torch.cuda.device(dev):
        _lib.check(lib.botgat_sample_count(h, n, _lib.ptr(seeds), int(fanout), _lib.ptr(offsets), C.byref(total),
                                           _lib.ptr(ws), _stream()), "botgat_sample_count")
        src = torch.empty(total.value, dtype=torch.int64, device=dev)
        dst = torch.empty_like(src)
        eid = torch.empty_like(src)
        if total.value:
            _lib.check(lib.botgat_sample_neighbors(h, n, _lib.ptr(seeds), int(fanout), int(seed), _lib.ptr(offsets),
                                                   _lib.ptr(src), _lib.ptr(dst), _lib.ptr(eid), _stream()),
                       "botgat_sample_neighbors")
    if n == 0:
        offsets.zero_()
    perm = g.edge_perm()
    if perm is not None and eid.numel():
        eid = perm.index_select(0, eid)   # the library numbers edges canonically; callers see edge ids
    return src, dst, eid, offsets


def to_block(g: Graph, seeds, src, dst_pos, eid):
    """``dgl.to_block`` for a frontier given as (src ids in ``g``, destination = position in ``seeds``, edge ids):
    destination nodes are ``seeds`` in order, source nodes are ``seeds`` followed by the remaining sampled sources
    in ascending id.  The block's frames gather the parent's ``ndata`` / ``edata`` lazily."""
    lib = _lib.load()
    dev = g.device
    seeds = torch.as_tensor(seeds, dtype=torch.int64, device=dev).contiguous()
    n_parent, n_seeds, n_edges = g.number_of_nodes(), seeds.numel(), src.numel()
    src_local = torch.empty(n_edges, dtype=torch.int64, device=dev)
    src_nodes = torch.empty(n_seeds + n_edges, dtype=torch.int64, device=dev)
    ws = torch.empty(lib.botgat_block_workspace_bytes(n_parent), dtype=torch.uint8, device=dev)
    n_src = C.c_int64()
    with torch.cuda.device(dev):
        _lib.check(lib.botgat_block_compact(n_parent, n_seeds, _lib.ptr(seeds), n_edges, _lib.ptr(src), _lib.ptr(src_local),
                                            _lib.ptr(src_nodes), C.byref(n_src), _lib.ptr(ws),
                                            dev.index if dev.index is not None else torch.cuda.current_device(), _stream()),
                   "botgat_block_compact")
    src_nodes = src_nodes[: n_src.value]
    # a frontier lists its edges seed by seed, i.e. sorted by destination: already canonical
    block = Graph(src_local, dst_pos, n_src.value, n_seeds, is_block=True, presorted=True)
    block.srcdata = _LazyFrame(g.ndata, src_nodes)
    block.dstdata = _LazyFrame(g.ndata, seeds)
    block.ndata = block.srcdata
    block.edata = _LazyFrame(g.edata, eid)
    return block


class MultiLayerNeighborSampler:
    """``dgl.dataloading.MultiLayerNeighborSampler(fanouts)``: ``fanouts[i]`` in-neighbours per node for layer i
    (``None`` or <= 0: every neighbour, i.e. ``MultiLayerFullNeighborSampler``)."""

    def __init__(self, fanouts, replace=False, return_eids=False):
        if replace:
            raise NotImplementedError("sampling with replacement is not used by the reference and not implemented")
        self.fanouts = [(-1 if f is None else int(f)) for f in fanouts]

    def sample_blocks(self, g: Graph, seed_nodes, seed=None):
        """Blocks for one batch, input layer first (the order ``model(subgraphs)`` consumes them)."""
        if seed is None:
            seed = int(torch.randint(0, 2**62, (1,)).item())
        blocks = []
        seeds = torch.as_tensor(seed_nodes, dtype=torch.int64, device=g.device)
        for layer in reversed(range(len(self.fanouts))):
            src, dst_pos, eid, _ = sample_neighbors(g, seeds, self.fanouts[layer], seed + layer)
            block = to_block(g, seeds, src, dst_pos, eid)
            blocks.insert(0, block)
            seeds = block.srcdata[NID]
        return blocks


class MultiLayerFullNeighborSampler(MultiLayerNeighborSampler):
    def __init__(self, n_layers, return_eids=False):
        super().__init__([None] * n_layers)


class NodeDataLoader:
    """``dgl.dataloading.NodeDataLoader(g, nids, block_sampler, ...)``: iterates ``(input_nodes, output_nodes,
    blocks)``.  ``batch_sampler`` (an iterable of index batches into ``nids``, as the reference's ``BatchSampler``,
    utils.py) wins over ``batch_size`` / ``shuffle`` / ``drop_last``."""

    def __init__(self, g: Graph, nids, block_sampler, batch_size=1, shuffle=False, drop_last=False, batch_sampler=None,
                 num_workers=0, device=None, **unused):
        self.g, self.sampler = g, block_sampler
        self.nids = torch.as_tensor(nids, dtype=torch.int64, device=g.device)
        self.batch_size, self.shuffle, self.drop_last, self.batch_sampler = batch_size, shuffle, drop_last, batch_sampler

        self._bs_iter = None   # ONE persistent iterator over batch_sampler (the reference's BatchSampler never ends)

    def _plain_batches(self):
        n = self.nids.numel()
        order = torch.randperm(n, device=self.nids.device) if self.shuffle else torch.arange(n, device=self.nids.device)
        for lo in range(0, n, self.batch_size):
            if self.drop_last and lo + self.batch_size > n:
                break
            yield self.nids[order[lo: lo + self.batch_size]]

    def __len__(self):
        if self.batch_sampler is not None:
            bs = self.batch_sampler
            if hasattr(bs, "__len__"):
                return len(bs)
            if hasattr(bs, "n") and hasattr(bs, "batch_size"):   # the reference's BatchSampler (utils.py:22-32): batches per epoch
                return (bs.n + bs.batch_size - 1) // bs.batch_size
            raise TypeError("NodeDataLoader: batch_sampler has no length")
        n = self.nids.numel()
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        return _NodeDataLoaderIter(self)


class _NodeDataLoaderIter:
    """Iterator of :class:`NodeDataLoader`.  With a ``batch_sampler`` it draws from the loader's ONE persistent
    iterator: the reference's ``BatchSampler`` (utils.py:22-32) is infinite and yields ``None`` as the end-of-epoch
    sentinel, and ``DataLoaderWrapper`` (utils.py:8-19) keeps calling ``next`` on the same iterator epoch after
    epoch — so ``None`` ends the epoch (StopIteration) and the NEXT call starts the following one."""

    def __init__(self, loader):
        self.loader = loader
        if loader.batch_sampler is not None:
            if loader._bs_iter is None:
                loader._bs_iter = iter(loader.batch_sampler)
            self._plain = None
        else:
            self._plain = loader._plain_batches()

    def __iter__(self):
        return self

    def __next__(self):
        ld = self.loader
        if self._plain is not None:
            seeds = next(self._plain)
        else:
            try:
                idx = next(ld._bs_iter)
            except StopIteration:
                ld._bs_iter = None      # a finite sampler: start over at the next epoch
                raise
            if idx is None:             # end-of-epoch sentinel
                raise StopIteration
            seeds = ld.nids[torch.as_tensor(idx, dtype=torch.int64).to(ld.nids.device)]
        blocks = ld.sampler.sample_blocks(ld.g, seeds)
        return blocks[0].srcdata[NID], seeds, blocks
