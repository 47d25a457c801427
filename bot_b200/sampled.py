"""Drop-in for the block/edge-feature ``GATConv`` shared by
``src/ogbn-proteins/models.py`` and ``src/ogbn-products/models.py`` (the two
classes are byte-identical up to a comment) and for the two ``GAT`` wrappers.

Same constructor arguments, parameter names (``src_fc``, ``dst_fc``,
``attn_src_fc``, ``attn_dst_fc``, ``attn_edge_fc``) and ``forward(graph, feat_src,
feat_edge=None)`` signature.  The DGL calls are replaced by one call of
``bot_b200.functional.gat_fused``; all ``nn.Linear`` projections stay torch matmuls.

Reference: GATConv       src/ogbn-proteins/models.py:19-168 (= products :20-167)
           proteins GAT  src/ogbn-proteins/models.py:171-264
           products GAT  src/ogbn-products/models.py:168-265
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from .functional import (Deferred, EdgeEmbedding, GATConvSampledFn, LayerTail, edge_logits, gat_conv_inference, gat_fused,
                         to_canonical)
from .no_sampling import KeptEdges, draw_attn_mul, draw_edge_keep


# Fold the layer's four node-side Linears into two GEMMs (functional.GATConvSampledFn).  False: one nn.Linear call
# each, exactly the reference's op sequence.
fold_projections = os.environ.get("BOTGAT_FOLD", "1") != "0"


class GATConv(nn.Module):
    attn_dropout_mode = "fused"  # see bot_b200.no_sampling.GATConv

    def __init__(self, node_feats, edge_feats, out_feats, n_heads=1, attn_drop=0.0, edge_drop=0.0,
                 negative_slope=0.2, residual=True, activation=None, use_attn_dst=True,
                 allow_zero_in_degree=True, use_symmetric_norm=False):
        super().__init__()
        if not residual:
            # unusable in the reference as well: it builds nn.Parameter(int) (models.py:49)
            # and calls dst_fc unconditionally (models.py:107)
            raise TypeError("GATConv(residual=False) is not constructible in the reference either (models.py:49)")
        self._n_heads = n_heads
        if isinstance(node_feats, tuple):
            self._in_src_feats, self._in_dst_feats = node_feats
        else:
            self._in_src_feats = self._in_dst_feats = node_feats
        self._out_feats = out_feats
        self._allow_zero_in_degree = allow_zero_in_degree
        self._use_symmetric_norm = use_symmetric_norm
        self._negative_slope = negative_slope

        self.src_fc = nn.Linear(self._in_src_feats, out_feats * n_heads, bias=False)
        self.dst_fc = nn.Linear(self._in_src_feats, out_feats * n_heads)  # residual branch, WITH bias (models.py:45)
        self.bias = None
        self.attn_src_fc = nn.Linear(self._in_src_feats, n_heads, bias=False)
        self.attn_dst_fc = nn.Linear(self._in_src_feats, n_heads, bias=False) if use_attn_dst else None
        self.attn_edge_fc = nn.Linear(edge_feats, n_heads, bias=False) if edge_feats > 0 else None
        self.attn_drop = nn.Dropout(attn_drop)
        self.edge_drop = edge_drop
        self.leaky_relu = nn.LeakyReLU(negative_slope, inplace=True)
        self.activation = activation
        self.reset_parameters()

    def reset_parameters(self):
        gain = nn.init.calculate_gain("relu")
        nn.init.xavier_normal_(self.src_fc.weight, gain=gain)
        nn.init.xavier_normal_(self.dst_fc.weight, gain=gain)
        nn.init.xavier_normal_(self.attn_src_fc.weight, gain=gain)
        for lin in (self.attn_dst_fc, self.attn_edge_fc):
            if lin is not None:
                nn.init.xavier_normal_(lin.weight, gain=gain)

    def set_allow_zero_in_degree(self, set_value):
        self._allow_zero_in_degree = set_value

    def forward(self, graph, feat_src, feat_edge=None, tail=None):
        # ``tail`` (functional.LayerTail, an extension of the reference signature): the model's elementwise layer tail,
        # fused into the forward kernel in inference; the layer then returns (h, act(norm(h))) instead of rst
        H, D = self._n_heads, self._out_feats
        with graph.local_scope():
            if not self._allow_zero_in_degree and graph.has_zero_in_degree:   # models.py:89-91
                assert False
            n_dst = graph.number_of_dst_nodes()
            if isinstance(feat_src, Deferred):    # prefetched by bot_b200.HostFeed
                feat_src = feat_src.wait()
            feat_dst = feat_src[:n_dst] if graph.is_block else feat_src       # models.py:93-96

            dst_scale = None
            if self._use_symmetric_norm:                                      # models.py:98-104, 150-156
                # applied to the raw input, so it flows through src_fc and attn_src_fc alike
                feat_src = feat_src * torch.pow(graph.srcdata["deg"], -0.5).view(-1, *([1] * (feat_src.dim() - 1)))
                dst_scale = torch.pow(graph.dstdata["deg"], 0.5).float().contiguous()

            fold = fold_projections and feat_src.is_cuda and feat_src.dim() == 2
            if not fold:
                ft = self.src_fc(feat_src).view(-1, H, D)                     # models.py:106
                resid = self.dst_fc(feat_dst).view(-1, H, D)                  # models.py:107
                el = self.attn_src_fc(feat_src)                               # models.py:108  (N_s,H)
                er = self.attn_dst_fc(feat_dst) if self.attn_dst_fc is not None else None   # models.py:122-124
            E = graph.number_of_edges()
            keep = attn_mul = eids = None
            attn_p, seed = 0.0, 0
            if self.training and self.edge_drop > 0:                          # models.py:136-141
                keep, eids = draw_edge_keep(E, self.edge_drop, feat_src.device)
            if self.training and self.attn_drop.p > 0:
                if self.attn_dropout_mode == "exact":
                    attn_mul = draw_attn_mul(self.attn_drop, E, H, feat_src.device, eids)
                else:
                    attn_p = self.attn_drop.p
                    seed = int(torch.randint(0, 2**62, (1,)).item())

            ee = None
            if feat_edge is not None:                                         # models.py:130-131
                if isinstance(feat_edge, Deferred):   # e.g. a host-to-device copy still in flight on another stream
                    feat_edge = feat_edge.wait()
                # the same Linear as a streaming kernel, emitted as one aligned 32-byte record per edge
                # (functional.pad_heads); padding columns are ignored by the kernels and get zero gradient
                if isinstance(feat_edge, EdgeEmbedding):   # encoder + ReLU + this Linear in one kernel
                    ee = feat_edge.logits(self.attn_edge_fc.weight)
                    if not feat_edge.canonical:
                        ee = to_canonical(graph, ee)
                else:                                      # a plain tensor is in edge-id order (DGL edata semantics)
                    ee = to_canonical(graph, edge_logits(feat_edge, self.attn_edge_fc.weight))   # (E, pad_heads(H))
            # every per-edge operand is now brought to the graph's canonical order (bot_b200.Graph): the keep set of
            # the selection draw already is (it is a uniformly random subset of positions either way); the literal
            # randperm replay and an "exact" dropout mask without edge-drop are in edge-id order
            if not isinstance(eids, KeptEdges):
                keep, attn_mul = to_canonical(graph, keep), to_canonical(graph, attn_mul)

            if fold:
                # the four node-side Linears (models.py:106-108,122-124) as two GEMMs whose outputs the kernels
                # read in place, the residual add (models.py:159-160) included
                w_src = torch.cat([self.src_fc.weight, self.attn_src_fc.weight], 0)
                w_dst = self.dst_fc.weight if self.attn_dst_fc is None else \
                    torch.cat([self.dst_fc.weight, self.attn_dst_fc.weight], 0)
                if (tail is not None and tail.usable() and not torch.is_grad_enabled() and not self.training
                        and self.activation is None):
                    return gat_conv_inference(graph, feat_src, feat_dst, w_src, w_dst, self.dst_fc.bias, ee, None, None,
                                              dst_scale, H, D, self._negative_slope, tail)
                rst = GATConvSampledFn.apply(graph, feat_src, feat_dst, w_src, w_dst, self.dst_fc.bias, ee, keep, attn_mul,
                                             dst_scale, H, D, self._negative_slope, attn_p, seed)
            else:
                rst = gat_fused(graph, ft, el, er, ee, keep, attn_mul, None, dst_scale,
                                self._negative_slope, attn_p, seed, edge_order="canonical")   # models.py:125-156
                rst = rst + resid                                             # models.py:159-160
            if self.activation is not None:
                rst = self.activation(rst, inplace=True)
            return rst


class _SampledGAT(nn.Module):
    """Shared body of the proteins / products ``GAT`` wrappers."""

    def _build(self, in0, edge_feats, n_classes, n_layers, n_heads, n_hidden, edge_emb, attn_drop, edge_drop,
               use_attn_dst, allow_zero_in_degree):
        self.convs = nn.ModuleList()
        self.norms = nn.ModuleList()
        for i in range(n_layers):
            in_hidden = n_heads * n_hidden if i > 0 else in0
            if self.edge_encoder is not None:
                self.edge_encoder.append(nn.Linear(edge_feats, edge_emb))
            self.convs.append(GATConv(in_hidden, edge_emb, n_hidden, n_heads=n_heads, attn_drop=attn_drop,
                                      edge_drop=edge_drop, use_attn_dst=use_attn_dst,
                                      allow_zero_in_degree=allow_zero_in_degree, use_symmetric_norm=False))
            self.norms.append(nn.BatchNorm1d(n_heads * n_hidden))
        self.pred_linear = nn.Linear(n_heads * n_hidden, n_classes)

    def _layers(self, subgraphs, h, always_residual):
        h_last = None
        for i in range(self.n_layers):
            efeat_emb = None
            if self.edge_encoder is not None:
                # relu(edge_encoder[i](efeat)) (models.py:245-247), left to the layer to fuse with attn_edge_fc
                # static edge features are kept in the graph's canonical order (permuted once, EdgeFrame.canonical)
                efeat_emb = EdgeEmbedding(subgraphs[i].edata.canonical("feat"), self.edge_encoder[i], canonical=True)
            # inference: residual + eval-mode BatchNorm + ReLU (models.py:253-260) ride in the gather kernel's epilogue
            tail = None
            if not torch.is_grad_enabled() and not self.training and self.activation in (F.relu, torch.relu):
                tail = LayerTail(h_last if (always_residual or self.residual) else None, self.norms[i], relu=True)
            h = self.convs[i](subgraphs[i], h, efeat_emb, tail=tail)
            if isinstance(h, tuple):
                h_last, h = h            # dropout is the identity in eval mode
                continue
            h = h.flatten(1, -1)
            if h_last is not None and (always_residual or self.residual):
                h = h + h_last[: h.shape[0], :]
            h_last = h
            h = self.dropout(self.activation(self.norms[i](h)))
        return self.pred_linear(h)


class ProteinsGAT(_SampledGAT):
    """``GAT`` of src/ogbn-proteins/models.py:171-264 (node encoder, per-layer edge encoder,
    unconditional residual, BatchNorm)."""

    def __init__(self, node_feats, edge_feats, n_classes, n_layers, n_heads, n_hidden, edge_emb, activation, dropout,
                 input_drop, attn_drop, edge_drop, use_attn_dst=True, allow_zero_in_degree=False):
        super().__init__()
        self.n_layers, self.n_heads, self.n_hidden, self.n_classes = n_layers, n_heads, n_hidden, n_classes
        self.node_encoder = nn.Linear(node_feats, n_hidden)
        # the reference leaves edge_encoder undefined for edge_emb == 0 (models.py:199-200); None here
        self.edge_encoder = nn.ModuleList() if edge_emb > 0 else None
        self._build(n_hidden, edge_feats, n_classes, n_layers, n_heads, n_hidden, edge_emb, attn_drop, edge_drop,
                    use_attn_dst, allow_zero_in_degree)
        self.input_drop = nn.Dropout(input_drop)
        self.dropout = nn.Dropout(dropout)
        self.activation = activation

    def forward(self, g):
        subgraphs = g if isinstance(g, list) else [g] * self.n_layers
        h = self.input_drop(F.relu(self.node_encoder(subgraphs[0].srcdata["feat"])))
        return self._layers(subgraphs, h, always_residual=True)


class ProductsGAT(_SampledGAT):
    """``GAT`` of src/ogbn-products/models.py:168-265 (no node encoder in the data path — it is
    constructed and counted in #Params but never called, models.py:198; residual gated by a flag)."""

    def __init__(self, node_feats, edge_feats, n_classes, n_layers, n_heads, n_hidden, edge_emb, activation, dropout,
                 input_drop, attn_drop, edge_drop, use_attn_dst=True, allow_zero_in_degree=False, residual=False):
        super().__init__()
        self.n_layers, self.n_heads, self.n_hidden, self.n_classes = n_layers, n_heads, n_hidden, n_classes
        self.node_encoder = nn.Linear(node_feats, n_hidden)
        self.edge_encoder = nn.ModuleList() if edge_emb > 0 else None
        self._build(node_feats, edge_feats, n_classes, n_layers, n_heads, n_hidden, edge_emb, attn_drop, edge_drop,
                    use_attn_dst, allow_zero_in_degree)
        self.input_drop = nn.Dropout(input_drop)
        self.dropout = nn.Dropout(dropout)
        self.activation = activation
        self.residual = residual

    def forward(self, g, inference=False):
        subgraphs = g if isinstance(g, list) else [g] * self.n_layers
        h = self.input_drop(subgraphs[0].srcdata["feat"])
        return self._layers(subgraphs, h, always_residual=False)
