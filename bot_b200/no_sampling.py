"""Drop-in for ``src/no-sampling/models.py`` of AiRyunn/BoT: ``GATConv`` (full-graph
variant), ``ElementWiseLinear`` and the ``GAT`` wrapper, with the same constructor
arguments, parameter names (``fc``, ``attn_l``, ``attn_r``, ``res_fc``) and
``forward(graph, feat)`` signature, so ``run.py`` can do ``from bot_b200.no_sampling
import GAT`` unchanged.  The DGL calls in the layer body are replaced by one call of
``bot_b200.functional.gat_fused`` (hand-written sm_100a kernels); the dense
projections stay torch matmuls.

Reference: GATConv  src/no-sampling/models.py:416-566
           GAT      src/no-sampling/models.py:644-736
           ElementWiseLinear  src/no-sampling/models.py:18-50
"""
from __future__ import annotations

import torch
import torch.nn as nn

import torch.nn.functional as F

from .functional import LayerTail, gat_fused, to_canonical


edge_drop_mode = "select"   # "randperm": torch.randperm exactly as the reference calls it (replays its generator)


class KeptEdges:
    """The kept edge ids (ascending) of a keep mask whose population is known — materialised only when the
    exact attention-dropout replay asks for them (``torch.nonzero_static``: no host sync)."""

    def __init__(self, keep, count):
        self.keep, self.count, self._ids = keep, count, None

    def numel(self):
        return self.count

    def ids(self):
        if self._ids is None:
            self._ids = torch.nonzero_static(self.keep, size=self.count).flatten()
        return self._ids


def draw_edge_keep(n_edges, edge_drop, device):
    """Edge-drop keep set with the reference's semantics (models.py:529-532): ``perm = randperm(E)``,
    ``perm[:int(E*p)]`` dropped, ``eids = perm[int(E*p):]`` kept — a uniformly random subset of exactly
    ``int(E*p)`` edges is dropped.  On CUDA the subset is drawn by ``botgat_edge_drop_draw`` (a selection on Philox
    keys seeded from torch's generator) instead of a sort of E keys; ``edge_drop_mode = "randperm"`` keeps the
    literal torch call.  Returns ``(keep uint8 (E,), kept ids)``."""
    bound = int(n_edges * edge_drop)
    device = torch.device(device)
    if edge_drop_mode == "randperm" or device.type != "cuda":
        perm = torch.randperm(n_edges, device=device)
        keep = torch.ones(n_edges, dtype=torch.uint8, device=device)
        keep[perm[:bound]] = 0
        return keep, perm[bound:]
    from .functional import edge_drop_keep

    seed = int(torch.randint(0, 2**62, (1,)).item())
    keep = edge_drop_keep(n_edges, bound, seed, device)
    return keep, KeptEdges(keep, n_edges - bound)


def draw_attn_mul(attn_drop_module, n_edges, n_heads, device, eids=None):
    """Attention-dropout multiplier m/(1-p) drawn by the module's own ``nn.Dropout``
    on a tensor shaped like the reference's ``edge_softmax`` output ((E',H,1),
    models.py:537/544), scattered to edge-id order."""
    rows = n_edges if eids is None else eids.numel()
    mul = attn_drop_module(torch.ones((rows, n_heads, 1), dtype=torch.float32, device=device))
    if eids is None:
        return mul.view(n_edges, n_heads)
    full = torch.zeros((n_edges, n_heads), dtype=torch.float32, device=device)
    full[eids.ids() if isinstance(eids, KeptEdges) else eids] = mul.view(rows, n_heads)
    return full


class ElementWiseLinear(nn.Module):
    """Per-feature scale and/or shift (models.py:18-50)."""

    def __init__(self, size, weight=True, bias=True, inplace=False):
        super().__init__()
        self.weight = nn.Parameter(torch.ones(size)) if weight else None
        self.bias = nn.Parameter(torch.zeros(size)) if bias else None
        self.inplace = inplace

    def reset_parameters(self):
        if self.weight is not None:
            nn.init.ones_(self.weight)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def forward(self, x):
        if self.inplace:
            if self.weight is not None:
                x.mul_(self.weight)
            if self.bias is not None:
                x.add_(self.bias)
            return x
        if self.weight is not None:
            x = x * self.weight
        if self.bias is not None:
            x = x + self.bias
        return x


class GATConv(nn.Module):
    """GAT layer, full-graph variant (models.py:416-566).

    ``attn_dropout_mode``: "fused" (default) draws the attention-dropout mask inside the kernels from a
    Philox stream keyed on (seed, canonical edge number, head) — nothing E-sized is materialised; "exact"
    draws it with the module's own ``nn.Dropout`` on an (E', H, 1) tensor exactly as the reference does
    (models.py:537/544; four extra E x H passes per layer and step) and is what the golden-vector tests replay.

    ``fold_logits``: compute ``el`` / ``er`` as extra columns of the ``fc`` / ``res_fc`` GEMMs
    (``W_el[h] = W[h]^T attn_l[h]``) instead of a pass over the (N, H, D) projection
    (models.py:517-521); the kernels read ``ft`` in place inside the wide GEMM output.
    """

    attn_dropout_mode = "fused"
    fold_logits = True

    def __init__(self, in_feats, out_feats, num_heads=1, feat_drop=0.0, attn_drop=0.0, edge_drop=0.0,
                 negative_slope=0.2, linear=True, activation=None, allow_zero_in_degree=False,
                 use_symmetric_norm=False, non_interactive_attn=False):
        super().__init__()
        self._num_heads = num_heads
        if isinstance(in_feats, tuple):
            self._in_src_feats, self._in_dst_feats = in_feats
        else:
            self._in_src_feats = self._in_dst_feats = in_feats
        self._out_feats = out_feats
        self._allow_zero_in_degree = allow_zero_in_degree
        self._use_symmetric_norm = use_symmetric_norm
        self._negative_slope = negative_slope
        hd = out_feats * num_heads
        if isinstance(in_feats, tuple):
            self.fc_src = nn.Linear(self._in_src_feats, hd, bias=False)
            self.fc_dst = nn.Linear(self._in_dst_feats, hd, bias=False)
        else:
            self.fc = nn.Linear(self._in_src_feats, hd, bias=False)
        self.attn_l = nn.Parameter(torch.empty(1, num_heads, out_feats))
        # the flag name is inverted in the reference (models.py:444-447): True ADDS attn_r
        if non_interactive_attn:
            self.attn_r = nn.Parameter(torch.empty(1, num_heads, out_feats))
        else:
            self.register_buffer("attn_r", None)
        self.feat_drop = nn.Dropout(feat_drop)
        self.attn_drop = nn.Dropout(attn_drop)
        self.edge_drop = edge_drop
        self.leaky_relu = nn.LeakyReLU(negative_slope)
        if linear:
            self.res_fc = nn.Linear(self._in_dst_feats, hd, bias=False)
        else:
            self.register_buffer("res_fc", None)
        self.reset_parameters()
        self._activation = activation

    def reset_parameters(self):
        gain = nn.init.calculate_gain("relu")
        for lin in ("fc", "fc_src", "fc_dst"):
            if hasattr(self, lin):
                nn.init.xavier_normal_(getattr(self, lin).weight, gain=gain)
        nn.init.xavier_normal_(self.attn_l, gain=gain)
        if isinstance(self.attn_r, nn.Parameter):
            nn.init.xavier_normal_(self.attn_r, gain=gain)
        if isinstance(self.res_fc, nn.Linear):
            nn.init.xavier_normal_(self.res_fc.weight, gain=gain)

    def set_allow_zero_in_degree(self, set_value):
        self._allow_zero_in_degree = set_value

    def forward(self, graph, feat, tail=None):
        # ``tail`` (functional.LayerTail, an extension of the reference signature): the model's elementwise layer tail,
        # fused into the forward kernel in inference; the layer then returns (h, act(norm(h))), both (N, H*D)
        H, D = self._num_heads, self._out_feats
        with graph.local_scope():
            if not self._allow_zero_in_degree and graph.has_zero_in_degree:  # models.py:477-479
                assert False
            n_dst = graph.number_of_dst_nodes()
            fold = (self.fold_logits and not isinstance(feat, tuple) and hasattr(self, "fc") and self.res_fc is not None
                    and feat.is_cuda and feat.dim() == 2 and self._activation is None)
            if fold:
                return self._forward_folded(graph, feat, n_dst, tail)
            if isinstance(feat, tuple):                                      # models.py:481-488
                h_src, h_dst = self.feat_drop(feat[0]), self.feat_drop(feat[1])
                if not hasattr(self, "fc_src"):
                    self.fc_src, self.fc_dst = self.fc, self.fc
                ft = self.fc_src(h_src).view(-1, H, D)
                ft_dst = self.fc_dst(h_dst).view(-1, H, D)
            else:                                                            # models.py:490-498
                h_src = self.feat_drop(feat)
                ft = self.fc(h_src).view(-1, H, D)
                if graph.is_block:
                    h_dst, ft_dst = h_src[:n_dst], ft[:n_dst]
                else:
                    h_dst, ft_dst = h_src, ft

            # symmetric normalisation: the source scale applies to messages and el only;
            # er and the residual see the unscaled projection (models.py:497-505, 521)
            src_scale = dst_scale = None
            if self._use_symmetric_norm:
                src_scale = graph.deg_scale("out", -0.5)
                dst_scale = graph.deg_scale("in", 0.5)                        # +0.5, models.py:552
            el = torch.einsum("nhd,hd->nh", ft, self.attn_l[0])               # models.py:517
            if src_scale is not None:
                el = el * src_scale.unsqueeze(-1)
            er = None
            if self.attn_r is not None:
                er = torch.einsum("nhd,hd->nh", ft_dst, self.attn_r[0])       # models.py:521

            if (tail is not None and tail.usable() and not torch.is_grad_enabled() and not self.training
                    and self._activation is None and ft.is_cuda):
                from .functional import gat_fused_inference
                res = self.res_fc(h_dst) if self.res_fc is not None else None   # models.py:557-560
                return gat_fused_inference(graph, ft, el, er, None, None, None, src_scale, dst_scale, self._negative_slope,
                                           tail, res)
            keep, attn_mul, attn_p, seed = self._draw(graph, graph.number_of_edges(), H, ft.device)   # models.py:528-537
            rst = gat_fused(graph, ft, el, er, None, keep, attn_mul, src_scale, dst_scale,
                            self._negative_slope, attn_p, seed, edge_order="canonical")   # models.py:523-555

            if self.res_fc is not None:                                       # models.py:557-560
                rst = rst + self.res_fc(h_dst).view(h_dst.shape[0], -1, D)
            if self._activation is not None:
                rst = self._activation(rst)
            return rst


    def _draw(self, graph, E, H, device):
        """Edge-drop keep set and attention dropout of one forward, in the graph's canonical edge order."""
        keep = attn_mul = eids = None
        attn_p, seed = 0.0, 0
        if self.training and self.edge_drop > 0:                              # models.py:528-537
            keep, eids = draw_edge_keep(E, self.edge_drop, device)
        if self.training and self.attn_drop.p > 0:
            if self.attn_dropout_mode == "exact":
                attn_mul = draw_attn_mul(self.attn_drop, E, H, device, eids)
            else:
                attn_p = self.attn_drop.p
                seed = int(torch.randint(0, 2**62, (1,)).item())
        if not isinstance(eids, KeptEdges):   # the literal randperm replay / an "exact" mask alone are in edge-id order
            keep, attn_mul = to_canonical(graph, keep), to_canonical(graph, attn_mul)
        return keep, attn_mul, attn_p, seed

    def _forward_folded(self, graph, feat, n_dst, tail=None):
        """models.py:490-560 with the logits folded into the projections: ONE wide GEMM per side,
        ``[ft | el] = h_src @ [W ; W_el]^T`` and ``[res | er] = h_dst @ [W_res ; W_er]^T``, read in place by the kernels."""
        from .functional import GATConvSampledFn

        H, D = self._num_heads, self._out_feats
        h_src = self.feat_drop(feat)
        h_dst = h_src[:n_dst] if graph.is_block else h_src
        Wv = self.fc.weight.view(H, D, -1)
        w_src = torch.cat([self.fc.weight, torch.einsum("hdi,hd->hi", Wv, self.attn_l[0])], 0)       # el, models.py:517
        w_dst = self.res_fc.weight
        if self.attn_r is not None:                                           # er from the UNSCALED projection, :521
            w_dst = torch.cat([w_dst, torch.einsum("hdi,hd->hi", Wv, self.attn_r[0])], 0)
        src_scale = dst_scale = None
        if self._use_symmetric_norm:
            src_scale, dst_scale = graph.deg_scale("out", -0.5), graph.deg_scale("in", 0.5)
        keep, attn_mul, attn_p, seed = self._draw(graph, graph.number_of_edges(), H, feat.device)
        zero_bias = torch.zeros(H * D, dtype=feat.dtype, device=feat.device)   # res_fc has no bias (models.py:453)
        if tail is not None and tail.usable() and not torch.is_grad_enabled() and not self.training:
            from .functional import gat_conv_inference
            return gat_conv_inference(graph, h_src, h_dst, w_src, w_dst, zero_bias, None, None, None, dst_scale, H, D,
                                      self._negative_slope, tail, src_scale)
        return GATConvSampledFn.apply(graph, h_src, h_dst, w_src, w_dst, zero_bias, None, keep, attn_mul, dst_scale, H, D,
                                      self._negative_slope, attn_p, seed, src_scale)


class GAT(nn.Module):
    """Stack of GATConv layers with the reference's per-layer epilogue (models.py:644-736)."""

    def __init__(self, dim_node, dim_edge, dim_output, n_hidden, n_layers, n_heads, activation, norm="none",
                 dropout=0.0, input_drop=0.0, attn_drop=0.0, edge_drop=0.0, non_interactive_attn=False,
                 use_symmetric_norm=False, linear=False, residual=False):
        super().__init__()
        self.n_node_feats = dim_node
        self.n_hidden = n_hidden
        self.n_classes = dim_output
        self.n_layers = n_layers
        self.num_heads = n_heads

        self.convs = nn.ModuleList()
        # the reference compares against "none:" (sic, models.py:672) so norms is always a
        # ModuleList; an empty one is falsy and selects the bias path in forward
        self.norms = nn.ModuleList()
        self.biases = nn.ModuleList()
        for i in range(n_layers):
            last = i == n_layers - 1
            in_hidden = n_heads * n_hidden if i > 0 else dim_node
            out_hidden = dim_output if last else n_hidden
            heads = 1 if last else n_heads
            self.convs.append(GATConv(in_hidden, out_hidden, num_heads=heads, attn_drop=attn_drop,
                                      edge_drop=edge_drop, non_interactive_attn=non_interactive_attn,
                                      use_symmetric_norm=use_symmetric_norm, linear=linear))
            if last:
                self.biases.append(ElementWiseLinear(out_hidden, weight=False, bias=True))
            elif norm == "batch":
                self.norms.append(nn.BatchNorm1d(heads * out_hidden))
            elif norm == "none":
                self.biases.append(ElementWiseLinear(heads * out_hidden, weight=False, bias=True))
        self.input_drop = nn.Dropout(input_drop)
        self.dropout = nn.Dropout(dropout)
        self.activation = activation
        self.residual = residual

    def forward(self, graph, feat):
        h = self.input_drop(feat)
        h_last = None
        for i in range(self.n_layers):
            tail = None
            if (i < self.n_layers - 1 and not torch.is_grad_enabled() and not self.training
                    and self.activation in (F.relu, torch.relu)):
                # inference: `h += h_last`, the eval-mode norm / bias and ReLU (models.py:720-731) ride in the kernel epilogue
                tail = LayerTail(h_last if self.residual else None, self.norms[i] if self.norms else self.biases[i], relu=True)
            h = self.convs[i](graph, h, tail=tail) if tail is not None else self.convs[i](graph, h)
            if isinstance(h, tuple):
                hn, h = h
                h_last = hn.view(hn.shape[0], self.convs[i]._num_heads, -1)   # dropout is the identity in eval mode
                continue
            if i < self.n_layers - 1:
                if self.residual and h_last is not None:
                    h = h + h_last
                h_last = h
                h = h.flatten(1)
                h = self.norms[i](h) if self.norms else self.biases[i](h)
                h = self.dropout(self.activation(h))
        h = h.mean(1)
        return self.biases[-1](h)


class GraphConv(nn.Module):
    """GCN layer (src/no-sampling/models.py:114-413), SURVEY.md section 8f rank 3: the `copy_src`/`sum` aggregation
    (models.py:374,381) is the GAT gather with uniform attention, so it runs on the same fused kernel —
    ``softmax`` of equal logits is 1/in_degree, undone by a destination scale of in_degree."""

    def __init__(self, in_feats, out_feats, norm="both", weight=True, bias=True, activation=None, allow_zero_in_degree=False):
        super().__init__()
        if norm not in ("none", "both", "right"):
            raise ValueError(f'Invalid norm value. Must be either "none", "both" or "right". But got "{norm}".')
        self._in_feats, self._out_feats, self._norm = in_feats, out_feats, norm
        self._allow_zero_in_degree = allow_zero_in_degree
        if weight:
            self.weight = nn.Parameter(torch.empty(in_feats, out_feats))
        else:
            self.register_parameter("weight", None)
        if bias:
            self.bias = nn.Parameter(torch.empty(out_feats))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()
        self._activation = activation

    def reset_parameters(self):
        if self.weight is not None:
            nn.init.xavier_uniform_(self.weight)
        if self.bias is not None:
            nn.init.zeros_(self.bias)

    def set_allow_zero_in_degree(self, set_value):
        self._allow_zero_in_degree = set_value

    def _aggregate(self, graph, x, src_scale):
        """sum over in-edges of src_scale[u] * x[u]  (update_all(copy_src, sum), models.py:374/381)"""
        n_src = x.shape[0]
        flat = x.reshape(n_src, 1, -1)
        if "gcn_zero" not in graph._cache or graph._cache["gcn_zero"].shape[0] != n_src:
            graph._cache["gcn_zero"] = torch.zeros(n_src, 1, device=x.device)
            graph._cache["gcn_indeg"] = graph.in_degrees().float().contiguous()
        out = gat_fused(graph, flat, graph._cache["gcn_zero"], None, None, None, None, src_scale,
                        graph._cache["gcn_indeg"], 0.2, 0.0, 0)
        return out.reshape((out.shape[0],) + tuple(x.shape[1:]))

    def forward(self, graph, feat, weight=None):
        with graph.local_scope():
            if not self._allow_zero_in_degree and graph.has_zero_in_degree:   # models.py:334-347
                raise RuntimeError("There are 0-in-degree nodes in the graph, output for those nodes will be invalid. "
                                   "Add self-loops or construct the module with allow_zero_in_degree=True.")
            feat_src, feat_dst = feat if isinstance(feat, tuple) else (feat, feat)
            src_scale = graph.deg_scale("out", -0.5) if self._norm == "both" else None      # models.py:351-356
            if weight is not None:
                if self.weight is not None:
                    raise RuntimeError("External weight is provided while at the same time the module has defined its "
                                       "own weight parameter. Please create the module with flag weight=False.")
            else:
                weight = self.weight
            if self._in_feats > self._out_feats:                              # models.py:368-376: W first
                if weight is not None:
                    # the source scale commutes with the projection; applying it in the kernel saves a pass
                    feat_src = torch.matmul(feat_src, weight)
                rst = self._aggregate(graph, feat_src, src_scale)
            else:                                                             # models.py:377-385: aggregate first
                rst = self._aggregate(graph, feat_src, src_scale)
                if weight is not None:
                    rst = torch.matmul(rst, weight)
            if self._norm != "none":                                          # models.py:387-395
                degs = graph.in_degrees().float().clamp(min=1)
                norm = torch.pow(degs, -0.5) if self._norm == "both" else 1.0 / degs
                rst = rst * norm.reshape(norm.shape + (1,) * (feat_dst.dim() - 1))
            if self.bias is not None:
                rst = rst + self.bias
            if self._activation is not None:
                rst = self._activation(rst)
            return rst


class GCN(nn.Module):
    """Stack of GraphConv layers (src/no-sampling/models.py:569-641)."""

    def __init__(self, in_feats, n_classes, n_hidden, n_layers, activation, norm="none", norm_adj="symm", dropout=0.0,
                 input_drop=0, residual=False, use_linear=False):
        super().__init__()
        self.n_layers, self.n_hidden, self.n_classes = n_layers, n_hidden, n_classes
        self.use_linear, self.residual = use_linear, residual
        self.convs = nn.ModuleList()
        if use_linear:
            self.linear = nn.ModuleList()
        self.norms = nn.ModuleList()  # the reference's `norm != "none:"` (sic) is always true
        for i in range(n_layers):
            in_hidden = n_hidden if i > 0 else in_feats
            out_hidden = n_hidden if i < n_layers - 1 else n_classes
            bias = norm == "none" or i == n_layers - 1
            self.convs.append(GraphConv(in_hidden, out_hidden, "both" if norm_adj == "symm" else "right", bias=bias))
            if use_linear:
                self.linear.append(nn.Linear(in_hidden, out_hidden, bias=False))
            if i < n_layers - 1 and norm == "batch":
                self.norms.append(nn.BatchNorm1d(out_hidden))
        self.input_drop = nn.Dropout(input_drop)
        self.dropout = nn.Dropout(dropout)
        self.activation = activation

    def forward(self, graph, feat):
        h = self.input_drop(feat)
        h_last = None
        for i in range(self.n_layers):
            conv = self.convs[i](graph, h)
            h = conv + self.linear[i](h) if self.use_linear else conv
            if i < self.n_layers - 1:
                if self.residual and h_last is not None:
                    h = h + h_last
                h_last = h
                if self.norms:
                    h = self.norms[i](h)
                h = self.dropout(self.activation(h))
        return h
