"""Graph object for the GAT engine: the DGLGraph subset BoT's GATConv touches.

Mirrors the protocol listed in SURVEY.md section 8b: ``local_scope()``,
``in_degrees()``, ``out_degrees()``, ``is_block``, ``number_of_*()``, ``device``,
dict-like ``srcdata/dstdata/ndata/edata``, ``to()``, ``remove_self_loop()``,
``add_self_loop()``, ``create_formats_()`` (reference call sites:
src/no-sampling/run.py:133-148, src/no-sampling/models.py:476-555,
src/ogbn-proteins/gat.py:54-68, src/ogbn-proteins/models.py:88-156).

Structure (in-CSR / out-CSR / edge-id maps / degree tables) is built on the GPU by
``libbotgat.so`` (``botgat_graph_create``); this module never computes it on the
host.
"""
from __future__ import annotations

import contextlib
import ctypes as C

import torch

from . import _lib


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"{what}: bot_b200 has no CPU path — move the graph to a CUDA device first (graph.to('cuda'))")


class _DevArray:
    """int32 device array exposed through ``__cuda_array_interface__``."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 2}


class EdgeFrame(dict):
    """``graph.edata``.  What the user reads and writes is in EDGE-ID order, as in DGL (``edata`` is indexed by
    edge id, src/ogbn-proteins/models.py:130-133).  The kernels stream per-edge operands in the graph's CANONICAL
    order (= in-CSR order, see :class:`Graph`); ``canonical(key)`` is the same tensor in that order, permuted ONCE
    per assignment and cached — static edge features therefore never pay a per-step permutation."""

    def __init__(self, graph):
        super().__init__()
        self._graph = graph
        self._canon = {}
        self._lazy = {}     # key -> tensor given in canonical order whose edge-id view has not been asked for

    def put_canonical(self, key, t):
        """Assign ``key`` from a tensor that is ALREADY in canonical order (static features kept canonically on the host
        and fed every step, ``bot_b200.HostFeed``): no permutation pass; the edge-id-order view ``self[key]`` is
        materialised only if somebody reads it."""
        if dict.__contains__(self, key):
            dict.__delitem__(self, key)
        self._canon.pop(key, None)
        self._lazy[key] = t

    def __setitem__(self, key, t):
        self._lazy.pop(key, None)
        dict.__setitem__(self, key, t)

    def __contains__(self, key):
        return dict.__contains__(self, key) or key in self._lazy

    def __getitem__(self, key):
        if not dict.__contains__(self, key) and key in self._lazy:
            t = self._lazy[key]
            inv = self._graph.canonical_edge_ids() if t.is_cuda else None
            dict.__setitem__(self, key, t if inv is None else t.index_select(0, inv))
            self._canon[key] = (dict.__getitem__(self, key), t)
        return dict.__getitem__(self, key)

    def canonical(self, key):
        if key in self._lazy:
            return self._lazy[key]
        t = self[key]
        hit = self._canon.get(key)
        if hit is None or hit[0] is not t:
            perm = self._graph.edge_perm() if t.is_cuda else None
            hit = (t, t if perm is None else t.index_select(0, perm))
            self._canon[key] = hit
        return hit[1]


class Graph:
    """Homogeneous graph or bipartite block (``is_block=True``: dst nodes are the
    first ``num_dst_nodes`` src nodes, as in DGL message-flow blocks).

    Edge order.  Edge ids are the positions in the COO the graph was built from (DGL semantics) and everything
    user-facing (``edges()``, ``edata[...]``, ``structure("in_eid")``) speaks edge ids.  Internally (``canonical=True``,
    the default) the device structure is built from the COO sorted by (dst, src): the library's own edge numbering
    then EQUALS the in-CSR position, neighbour lists are sorted in both CSRs, and per-edge operands given in that
    canonical order (``EdgeFrame.canonical``, ``functional.gat_fused(..., edge_order="canonical")``) need no
    edge-id -> CSR permutation pass at all in the forward and a cache-blocked one in the backward
    (csrc/edge_ops.cu).  ``edge_perm()`` is the canonical -> edge-id map, ``canonical_edge_ids()`` its inverse.
    ``canonical=False`` builds the structure from the COO as given (DGL's order: inside a row, increasing edge id);
    ``presorted=True`` promises a COO already sorted by dst (the sampler's blocks) and skips the sort."""

    def __init__(self, src, dst, num_src_nodes=None, num_dst_nodes=None, is_block=False, canonical=True, presorted=False):
        src = torch.as_tensor(src, dtype=torch.int64)
        dst = torch.as_tensor(dst, dtype=torch.int64)
        if src.shape != dst.shape or src.dim() != 1:
            raise ValueError("src and dst must be 1-D tensors of equal length")
        if src.device != dst.device:
            raise ValueError("src and dst must live on one device")
        self._src, self._dst = src.contiguous(), dst.contiguous()
        if num_src_nodes is None:
            num_src_nodes = int(max(src.max().item(), dst.max().item())) + 1 if src.numel() else 0
        self._n_src = int(num_src_nodes)
        self._n_dst = int(self._n_src if num_dst_nodes is None else num_dst_nodes)
        self.is_block = bool(is_block)
        self._canonical, self._presorted = bool(canonical), bool(presorted)
        self._perm = self._inv = None    # canonical position -> edge id, and back (None = identity)
        self.edata = EdgeFrame(self)
        if self.is_block or self._n_src != self._n_dst:
            self.srcdata, self.dstdata = {}, {}
            self.ndata = self.srcdata
        else:
            self.ndata = {}
            self.srcdata = self.dstdata = self.ndata
        self._handle = None
        self._info = None
        self._cache = {}
        self._streams = {}   # cuda_stream handle -> torch stream: every stream the structure was used on

    # ---- lifetime of the device structure ---------------------------------
    def _ensure(self):
        """The device structure handle.  Every kernel launch that reads the structure goes through here, so this
        is also where the launching stream is remembered: ``__del__`` frees behind ALL of them."""
        if self._src.is_cuda:
            st = torch.cuda.current_stream(self._src.device)
            if st.cuda_stream not in self._streams:
                self._streams[st.cuda_stream] = st
        if self._handle is not None:
            return self._handle
        _require_cuda(self._src, "graph structure")
        lib = _lib.load()
        h = C.c_void_p()
        src, dst = self._src, self._dst
        with torch.cuda.device(self._src.device):
            if self._canonical and not self._presorted and src.numel() > 1:
                # canonical edge order = COO sorted by (dst, src), ties by edge id (a data movement, not arithmetic;
                # out-of-range ids are caught by botgat_graph_create below)
                key = dst * max(self._n_src, 1) + src
                if not bool((key[1:] >= key[:-1]).all()):
                    perm = torch.argsort(key, stable=True)
                    src, dst = src.index_select(0, perm), dst.index_select(0, perm)
                    self._perm = perm
                del key
            rc = lib.botgat_graph_create(self._n_src, self._n_dst, src.numel(), _lib.ptr(src),
                                         _lib.ptr(dst), self._src.device.index, _stream(), C.byref(h))
        if rc != 0:
            self._perm = None
        _lib.check(rc, "botgat_graph_create")
        del src, dst
        self._handle = h
        info = _lib.GraphInfo()
        _lib.check(lib.botgat_graph_get_info(h, C.byref(info)), "botgat_graph_get_info")
        self._info = info
        return h

    def __del__(self):
        h, self._handle = getattr(self, "_handle", None), None
        if h is not None and _lib._lib is not None:
            try:
                # back to the stream-ordered pool behind the work enqueued so far (no device synchronisation:
                # mini-batch blocks come and go every step).  The free is enqueued on the current stream; kernels
                # that read the structure on OTHER streams (a side stream, a user stream) are ordered before it first.
                cur = torch.cuda.current_stream(self._src.device)
                for handle, st in self._streams.items():
                    if handle != cur.cuda_stream:
                        cur.wait_stream(st)
                _lib._lib.botgat_graph_destroy_async(h, C.c_void_p(cur.cuda_stream))
            except Exception:
                try:
                    _lib._lib.botgat_graph_destroy(h)
                except Exception:
                    pass

    def edge_perm(self):
        """canonical position -> edge id (int64, device), or None when the two orders coincide."""
        self._ensure()
        return self._perm

    def canonical_edge_ids(self):
        """edge id -> canonical position (the library's own edge numbering; the in-kernel Philox streams are keyed
        on it), or None when the two orders coincide."""
        self._ensure()
        if self._perm is not None and self._inv is None:
            inv = torch.empty_like(self._perm)
            inv[self._perm] = torch.arange(self._perm.numel(), device=self._perm.device)
            self._inv = inv
        return self._inv

    def create_formats_(self):
        """``graph.create_formats_()`` (run.py:146): materialise CSR + CSC now."""
        if self._src.is_cuda:
            self._ensure()

    # ---- DGLGraph protocol -------------------------------------------------
    @property
    def device(self):
        return self._src.device

    def edges(self, order="eid"):
        return self._src, self._dst

    def number_of_edges(self):
        return self._src.numel()

    num_edges = number_of_edges

    def number_of_nodes(self):
        return self._n_src

    num_nodes = number_of_nodes

    def number_of_src_nodes(self):
        return self._n_src

    def number_of_dst_nodes(self):
        return self._n_dst

    def structure(self, name):
        """Copy of one device structure array as an int64 tensor (names: _lib.ARRAYS)."""
        h = self._ensure()
        p, n = C.c_void_p(), C.c_int64()
        _lib.check(_lib.load().botgat_graph_get(h, _lib.ARRAYS.index(name), C.byref(p), C.byref(n)), "botgat_graph_get")
        if n.value == 0:
            return torch.empty(0, dtype=torch.int64, device=self.device)
        view = torch.as_tensor(_DevArray(p.value, n.value), device=self.device)  # zero-copy view of library memory
        out = view.long()  # the copy is what escapes
        if name.endswith("_eid") and self._perm is not None:
            out = self._perm.index_select(0, out)   # the library numbers edges canonically; users see edge ids
        return out

    def in_degrees(self):
        if "in_deg" not in self._cache:
            self._cache["in_deg"] = self.structure("in_deg")
        return self._cache["in_deg"]

    def out_degrees(self):
        if "out_deg" not in self._cache:
            self._cache["out_deg"] = self.structure("out_deg")
        return self._cache["out_deg"]

    @property
    def has_zero_in_degree(self):
        """Cached replacement of ``(graph.in_degrees() == 0).any()`` (models.py:478): no host sync per forward."""
        self._ensure()
        return bool(self._info.has_zero_in_degree)

    def deg_scale(self, which, power):
        """``clamp(deg,1)^power`` exactly as the reference spells it (models.py:501-502, 551-552), cached."""
        key = (which, power)
        if key not in self._cache:
            deg = self.out_degrees() if which == "out" else self.in_degrees()
            self._cache[key] = torch.pow(deg.float().clamp(min=1), power).contiguous()
        return self._cache[key]

    @contextlib.contextmanager
    def local_scope(self):
        saved = [dict(d) for d in (self.srcdata, self.dstdata, self.edata)]
        try:
            yield
        finally:
            for d, s in zip((self.srcdata, self.dstdata, self.edata), saved):
                d.clear()
                d.update(s)

    def to(self, device):
        device = torch.device(device)
        if device.type == "cuda" and device.index is None:
            device = torch.device("cuda", torch.cuda.current_device())
        if device == self.device:
            return self
        g = Graph(self._src.to(device), self._dst.to(device), self._n_src, self._n_dst, self.is_block,
                  canonical=self._canonical, presorted=self._presorted)
        for name in ("srcdata", "dstdata", "edata"):
            getattr(g, name).update({k: v.to(device) for k, v in getattr(self, name).items()})
        return g

    def cpu(self):
        return self.to("cpu")

    # ---- preprocessing (run.py:133-148) ------------------------------------
    def _new_like(self, src, dst):
        if self.is_block or self._n_src != self._n_dst:
            raise RuntimeError("preprocessing ops are defined for homogeneous graphs only")
        g = Graph(src, dst, self._n_src)
        g.ndata.update(self.ndata)
        return g

    def remove_self_loop(self):
        _require_cuda(self._src, "remove_self_loop")
        lib, e = _lib.load(), self._src.numel()
        o_s, o_d, n = torch.empty_like(self._src), torch.empty_like(self._dst), C.c_int64()
        with torch.cuda.device(self.device):
            rc = lib.botgat_coo_remove_self_loop(e, _lib.ptr(self._src), _lib.ptr(self._dst), _lib.ptr(o_s), _lib.ptr(o_d),
                                                 C.byref(n), self.device.index, _stream())
        _lib.check(rc, "botgat_coo_remove_self_loop")
        return self._new_like(o_s[: n.value].clone(), o_d[: n.value].clone())

    def add_self_loop(self):
        _require_cuda(self._src, "add_self_loop")
        lib, e = _lib.load(), self._src.numel()
        o_s = torch.empty(e + self._n_src, dtype=torch.int64, device=self.device)
        o_d = torch.empty_like(o_s)
        with torch.cuda.device(self.device):
            rc = lib.botgat_coo_add_self_loop(self._n_src, e, _lib.ptr(self._src), _lib.ptr(self._dst), _lib.ptr(o_s),
                                              _lib.ptr(o_d), self.device.index, _stream())
        _lib.check(rc, "botgat_coo_add_self_loop")
        return self._new_like(o_s, o_d)

    def to_bidirected(self):
        """``dgl.to_bidirected(graph)`` (run.py:137): drops node/edge data like DGL does."""
        _require_cuda(self._src, "to_bidirected")
        if self.is_block or self._n_src != self._n_dst:
            raise RuntimeError("to_bidirected is defined for homogeneous graphs only")
        lib, e = _lib.load(), self._src.numel()
        o_s = torch.empty(2 * e, dtype=torch.int64, device=self.device)
        o_d, n = torch.empty_like(o_s), C.c_int64()
        with torch.cuda.device(self.device):
            rc = lib.botgat_coo_to_bidirected(self._n_src, e, _lib.ptr(self._src), _lib.ptr(self._dst), _lib.ptr(o_s),
                                              _lib.ptr(o_d), C.byref(n), self.device.index, _stream())
        _lib.check(rc, "botgat_coo_to_bidirected")
        return Graph(o_s[: n.value].clone(), o_d[: n.value].clone(), self._n_src)


def graph(data, num_nodes=None, device=None):
    """``dgl.graph((src, dst), num_nodes=...)`` look-alike."""
    src, dst = data
    src = torch.as_tensor(src, dtype=torch.int64)
    dst = torch.as_tensor(dst, dtype=torch.int64)
    if device is not None:
        src, dst = src.to(device), dst.to(device)
    return Graph(src, dst, num_nodes)


def to_bidirected(g):
    return g.to_bidirected()


def remove_self_loop(g):
    return g.remove_self_loop()


def add_self_loop(g):
    return g.add_self_loop()


def create_block(data, num_src_nodes, num_dst_nodes, device=None, **kw):
    """``dgl.create_block`` look-alike: bipartite block whose dst nodes are a prefix of src."""
    src, dst = data
    src = torch.as_tensor(src, dtype=torch.int64)
    dst = torch.as_tensor(dst, dtype=torch.int64)
    if device is not None:
        src, dst = src.to(device), dst.to(device)
    return Graph(src, dst, num_src_nodes, num_dst_nodes, is_block=True, **kw)
