"""Host-side mirror of the callers of the path: MessagePassing, the four drop-in target layers and the multi-hop / pooling callers.

`tensorlayerx` (the reference's nn substrate) is not available offline, so these are plain `torch.nn.Module`s that
restate, line for line in meaning, the `forward` of

    MessagePassing   /root/reference/gammagl/layers/conv/message_passing.py:35-167
    GCNConv          /root/reference/gammagl/layers/conv/gcn_conv.py:78-108   (+ utils/degree.py:10-40)
    SAGEConv         /root/reference/gammagl/layers/conv/sage_conv.py:56-108  (mean / gcn / pool)
    GATConv          /root/reference/gammagl/layers/conv/gat_conv.py:98-122   (+ utils/softmax.py:29-35)
    GINConv          /root/reference/gammagl/layers/conv/gin_conv.py:49-63
    SGConv, APPNPConv  /root/reference/gammagl/layers/conv/sgc_conv.py:35-51, appnp_conv.py:9-37 (multi-hop: K propagates on
                     one cached CSR), global_{sum,mean,max,min}_pool  /root/reference/gammagl/layers/pool/glob.py:7-135

with the same constructor arguments, argument names and weight layout (`Linear` weight stored [in, out] as in
tensorlayerx, tests/layers/conv/test_gcn_conv.py:14-17).  `fused=True` (default) routes propagate() through the
fused kernels (gspmm / bspmm / segment_softmax) -- the `message_aggregate` plug point the reference left commented
out (gcn_conv.py:110-115, gat_conv.py:124-129); `fused=False` runs the reference's own message + aggregate
decomposition on the B200 segment operators (materialised [E,K] messages).  Both are parity-tested.
The dense X.W products stay torch.matmul (cuBLAS): tensor cores live there, outside the aggregation metric.
"""
import math
import os

import torch

from . import mpops
from .csr import derived_of, segment_csr_of


_EPILOGUE = os.environ.get('GGLB200_EPILOGUE', '1') != '0'   # 0: bias / activation as separate torch ops (same results)


# 'tc' (default): every mirrored Linear whose shape the tcgen05 kernel supports (CUDA fp32, in / out features multiples of 4 up to
# 256) runs mpops.linear_tc (3xTF32 on the tensor cores, 1.3e-6 vs fp64 -- inside the 1e-5 fp32 tolerance; C3 100 -> 256: 0.93 ms vs
# 2.95 ms for the library GEMM) in forward and input gradient; other shapes and 'lib' use torch.matmul.  Per layer:
# Linear(..., tensor_core=True / False).
LAYER_LINEAR = os.environ.get('GGLB200_LAYER_LINEAR', 'tc')


class Linear(torch.nn.Module):
    """tlx.layers.Linear: y = x @ W (+ b), W stored [in_features, out_features]."""

    def __init__(self, in_features, out_features, bias=False, tensor_core=None):
        super().__init__()
        self.tensor_core = tensor_core
        w = torch.empty(in_features, out_features)
        bound = math.sqrt(6.0 / (in_features + out_features))  # xavier_uniform
        torch.nn.init.uniform_(w, -bound, bound)
        self.weights = torch.nn.Parameter(w)
        self.biases = torch.nn.Parameter(torch.zeros(out_features)) if bias else None

    def forward(self, x):
        tc = self.tensor_core if self.tensor_core is not None else LAYER_LINEAR == 'tc'
        if tc and x.is_cuda and x.dim() == 2 and x.shape[0] > 0 and mpops.linear_tc_supported(x, self.weights):
            return mpops.linear_tc(x, self.weights, self.biases)
        y = x @ self.weights
        return y if self.biases is None else y + self.biases


class MessagePassing(torch.nn.Module):
    """message_passing.py:35-167.  propagate() takes the fused branch when the subclass did not override
    `message` (the base message is gather x weight, exactly what gspmm fuses) and `fused` is on."""

    def __init__(self, fused=True, fuse_max=True):
        super().__init__()
        self.fused = fused
        # aggr='max': the reference layer runs message + unsorted_segment_max, whose backward credits the FIRST winning edge
        # only (src/segment_max.cpp:48-61), while its gspmm-max operator credits every parallel edge of the winning source
        # (cpu/spmm_max_cpu.cpp:84-96).  The fused max used here (mpops.gspmm_max_edge) records the winning EDGE, so it
        # reproduces the layer's gradients on multigraphs as well and never materialises the [E, K] message;
        # fuse_max=False keeps the decomposition (size-guarded).
        self.fuse_max = fuse_max

    def message(self, x, edge_index, edge_weight=None):
        msg = x[edge_index[0, :]]                                           # message_passing.py:56
        if edge_weight is not None:
            return msg * edge_weight.unsqueeze(-1)                          # message_passing.py:57-59
        return msg

    def aggregate(self, msg, edge_index, num_nodes=None, aggr='sum'):
        dst_index = edge_index[1, :]                                        # message_passing.py:84
        if aggr == 'sum':
            return mpops.unsorted_segment_sum(msg, dst_index, num_nodes)
        elif aggr == 'mean':
            return mpops.unsorted_segment_mean(msg, dst_index, num_nodes)
        elif aggr == 'max':
            return mpops.unsorted_segment_max(msg, dst_index, num_nodes)
        raise NotImplementedError('Not support for this opearator')

    def message_aggregate(self, x, edge_index, edge_weight=None, aggr='sum', num_nodes=None):
        if edge_weight is not None and edge_weight.dim() == 2 and x.dim() == 3:
            return mpops.bspmm(edge_index, edge_weight, x, aggr)            # per-head weights (GAT)
        if mpops.weight_needs_grad(edge_weight) and x.dim() == 2 and aggr == 'sum':
            # gspmm has no gradient for the weight (src/gspmm.cpp:30,79) but the reference layer's message * weight has:
            # the one-head bspmm is the same forward kernel plus the SDDMM weight gradient
            return mpops.bspmm(edge_index, edge_weight.reshape(-1, 1), x.unsqueeze(1), 'sum').squeeze(1)
        if aggr == 'max':
            return mpops.gspmm_max_edge(edge_index, edge_weight, x, num_dst=num_nodes)
        return mpops.gspmm(edge_index, edge_weight, x, aggr, num_dst=num_nodes)   # message_passing.py:99-104 (+ bipartite rows)

    def update(self, x):
        return x

    def propagate(self, x, edge_index, aggr='sum', **kwargs):
        if kwargs.get('num_nodes') is None:
            kwargs['num_nodes'] = x.shape[0]                                # message_passing.py:141-142
        edge_weight = kwargs.get('edge_weight')
        n_dst = kwargs['num_nodes']
        base_message = type(self).message is MessagePassing.message
        square = n_dst == x.shape[0]
        w_grad = mpops.weight_needs_grad(edge_weight)
        per_head = edge_weight is not None and edge_weight.dim() == 2 and x.dim() == 3
        fusable = (self.fused and base_message
                   and (square or not (per_head or (w_grad and aggr == 'sum')))          # bspmm is square-only
                   and (not w_grad or (aggr == 'sum' and x.dim() in (2, 3)))                # weight gradient: sum only (bspmm)
                   and (aggr != 'max' or (self.fuse_max and not per_head and not w_grad)))
        if fusable:
            out = self.message_aggregate(x, edge_index, edge_weight=edge_weight, aggr=aggr, num_nodes=n_dst)
        else:
            if base_message:
                mpops.check_message_bytes(edge_index.shape[1], x[0].numel() if x.dim() > 1 else 1, x.element_size(),
                                          '%s.propagate(aggr=%r)' % (type(self).__name__, aggr))
                msg = self.message(x, edge_index, edge_weight)
            else:
                msg = self.message(x, edge_index)
            out = self.aggregate(msg, edge_index, num_nodes=n_dst, aggr=aggr)
        return self.update(out)


def degree(index, num_nodes=None, dtype=torch.float32):
    """utils/degree.py:10-40: unsorted_segment_sum of ones == the row lengths of the segment CSR over `index` (exact
    integers; the CSR is the one the aggregation over the same index uses, built once and cached)."""
    N = int(index.max().item()) + 1 if num_nodes is None else num_nodes
    if not index.is_cuda:
        raise RuntimeError('degree: index is on %s: gammagl_b200 is a CUDA-only backend (no CPU fallback)' % index.device)
    return segment_csr_of(index, N).degrees().to(dtype)


def calc_gcn_norm(edge_index, num_nodes, edge_weight=None):
    """utils/norm.py:5-30."""
    src, dst = edge_index[0], edge_index[1]
    if edge_weight is None:
        edge_weight = torch.ones((edge_index.shape[1], 1), device=edge_index.device)
    deg = mpops.unsorted_segment_sum(edge_weight, src, num_nodes).reshape(-1)
    dis = deg.pow(-0.5)
    return dis[src] * edge_weight.reshape(-1) * dis[dst]


def add_self_loops(edge_index, num_nodes=None):
    """utils/loop.py:107-150 (unweighted form): append (i, i) for every node."""
    N = int(edge_index.max().item()) + 1 if num_nodes is None else num_nodes
    loops = torch.arange(N, device=edge_index.device, dtype=edge_index.dtype)
    return torch.cat([edge_index, torch.stack([loops, loops])], dim=1)


class GCNConv(MessagePassing):
    def __init__(self, in_channels, out_channels, norm='both', add_bias=True, fused=True, fold_norm=True, aggregate_first=False,
                 fused_activation=None):
        super().__init__(fused)
        self.fold_norm = fold_norm
        if fused_activation not in (None, 'relu'):
            raise ValueError("fused_activation must be None or 'relu'")
        # the ReLU the models apply right after the layer (models/gcn.py), fused with the bias into the
        # aggregation kernel's row write on the folded path; applied as a separate op on the other paths.  Default: none.
        self.fused_activation = fused_activation
        # A(XW) == (AX)W; aggregating first moves in_channels instead of out_channels floats per edge.
        # Off by default (the reference always transforms first, gcn_conv.py:79; the other order rounds differently).
        self.aggregate_first = aggregate_first and in_channels < out_channels
        if norm not in ['left', 'right', 'none', 'both']:
            raise ValueError('Invalid norm value. Must be either "none", "both", "right" or "left".'
                             ' But got "{}".'.format(norm))
        self.in_channels, self.out_channels, self.add_bias, self._norm = in_channels, out_channels, add_bias, norm
        self.linear = Linear(in_channels, out_channels, bias=False)
        self.bias = torch.nn.Parameter(torch.zeros(1, out_channels)) if add_bias else None

    def forward(self, x, edge_index, edge_weight=None, num_nodes=None):
        relu = self.fused_activation == 'relu'
        if _EPILOGUE and self.fused and self.fold_norm and edge_weight is None and not self.aggregate_first:
            # bias and activation ride in the aggregation kernel's epilogue (bit-identical to the separate passes)
            return mpops.gcn_propagate(edge_index, self.linear(x), self._norm, bias=self.bias if self.add_bias else None, relu=relu)
        if self.aggregate_first:
            agg = self._aggregate(x, edge_index, edge_weight)
            if self.fused and mpops.linear_tc_supported(agg, self.linear.weights):
                # (A X) W + b (+ ReLU) in ONE tcgen05 GEMM with the epilogue fused (fp32-accurate 3xTF32, csrc/linear_tc.cu)
                return mpops.linear_tc(agg, self.linear.weights, self.bias if self.add_bias else None, relu)
            out = self.linear(agg)
        else:
            out = self._aggregate(self.linear(x), edge_index, edge_weight)  # gcn_conv.py:79
        out = out + self.bias if self.add_bias else out                     # gcn_conv.py:104-106
        return torch.relu(out) if relu else out

    def _aggregate(self, x, edge_index, edge_weight=None):
        if self.fused and self.fold_norm and edge_weight is None:
            return mpops.gcn_propagate(edge_index, x, self._norm)          # same math, no per-edge weight stream
        src, dst = edge_index[0], edge_index[1]
        if edge_weight is None:
            edge_weight = torch.ones((edge_index.shape[1],), device=x.device)
        edge_weight = edge_weight.reshape(-1)
        weights = edge_weight
        num_nodes = x.shape[0]
        if self._norm in ['left', 'both']:                                  # gcn_conv.py:88-94
            deg = degree(src, num_nodes=num_nodes, dtype=torch.float32)
            norm = deg.pow(-0.5) if self._norm == 'both' else 1.0 / deg
            weights = norm[src] * edge_weight
        if self._norm in ['right', 'both']:                                 # gcn_conv.py:96-102
            deg = degree(dst, num_nodes=num_nodes, dtype=torch.float32)
            norm = deg.pow(-0.5) if self._norm == 'both' else 1.0 / deg
            weights = weights.reshape(-1) * norm[dst]
        return self.propagate(x, edge_index, edge_weight=weights, num_nodes=num_nodes)


class SAGEConv(MessagePassing):
    def __init__(self, in_channels, out_channels, activation=None, aggr='mean', add_bias=True, fused=True, fold_norm=True,
                 aggregate_first=False):
        super().__init__(fused)
        self.fold_norm = fold_norm
        # aggr='mean' only: mean(W x_j) = W mean(x_j), so the neighbour Linear can run on the n_dst aggregated rows instead of the
        # n_src source rows -- in a sampled block n_src is 10-25x n_dst, and `src_feat` may then be the FULL feature matrix with
        # global source ids in the edge list: the x[n_id] gather disappears into the aggregation kernel (examples/sage_minibatch.py).
        self.aggregate_first = aggregate_first
        if aggr not in ('mean', 'gcn', 'pool'):
            raise NotImplementedError("aggr '%s' (lstm) is outside the aggregation path" % aggr)
        self.aggr, self.act, self.in_feat = aggr, activation, in_channels
        self.fc_neigh = Linear(in_channels, out_channels)
        if aggr != 'gcn':
            self.fc_self = Linear(in_channels, out_channels)
        if aggr == 'pool':
            self.pool = Linear(in_channels, in_channels)
        self.add_bias = add_bias
        self.bias = torch.nn.Parameter(torch.zeros(1, out_channels)) if add_bias else None

    def forward(self, feat, edge):
        src_feat, dst_feat = feat if isinstance(feat, tuple) else (feat, feat)
        num_nodes = int(dst_feat.shape[0])
        if self.aggr == 'mean' and self.aggregate_first:
            out = self.fc_neigh(self.propagate(src_feat, edge, edge_weight=None, num_nodes=num_nodes, aggr='mean'))
        elif self.aggr == 'mean':                                           # sage_conv.py:79-81
            src_feat = self.fc_neigh(src_feat)
            out = self.propagate(src_feat, edge, edge_weight=None, num_nodes=num_nodes, aggr='mean')
        elif self.aggr == 'gcn':                                            # sage_conv.py:82-88
            src_feat = self.fc_neigh(src_feat)
            raw = edge

            def with_loops():                                               # two host syncs + a 2 (E + n) copy: once per edge list
                n_ = int(1 + raw[0].max().item())
                return n_, int(raw[1].max().item()), add_self_loops(raw, n_)
            n, max_dst, edge = derived_of(raw, 'sage_gcn_self_loops', with_loops) if raw.is_cuda else with_loops()
            if self.fused and self.fold_norm and n == num_nodes == int(src_feat.shape[0]) and max_dst < n:
                # the same sum with calc_gcn_norm's weights folded into two row scalings and the bias into the row write:
                # no [E] weight vector, no w[perm] gather, no separate bias pass (mpops.gcn_propagate, norm of utils/norm.py)
                out = mpops.gcn_propagate(edge, src_feat, 'both_src', bias=self.bias if self.add_bias else None)
                return self.act(out) if self.act is not None else out
            weight = calc_gcn_norm(edge, n)
            out = self.propagate(src_feat, edge, edge_weight=weight, num_nodes=n, aggr='sum')
            out = out[:num_nodes]
        else:                                                               # pool, sage_conv.py:89-92
            src_feat = torch.relu(self.pool(src_feat))
            out = self.propagate(src_feat, edge, edge_weight=None, num_nodes=num_nodes, aggr='max')
            out = self.fc_neigh(out)
        if self.aggr != 'gcn':
            out = out + self.fc_self(dst_feat)
        if self.add_bias:
            out = out + self.bias
        if self.act is not None:
            out = self.act(out)
        return out


class GATConv(MessagePassing):
    def __init__(self, in_channels, out_channels, heads=1, concat=True, negative_slope=0.2, dropout_rate=0., add_bias=True,
                 fused=True, fuse_attention=True):
        super().__init__(fused)
        self.fuse_attention = fuse_attention
        self.in_channels, self.out_channels, self.heads, self.concat = in_channels, out_channels, heads, concat
        self.negetive_slop, self.dropout_rate, self.add_bias = negative_slope, dropout_rate, add_bias
        self.w = torch.nn.Parameter(torch.nn.init.trunc_normal_(torch.empty(in_channels, out_channels * heads), std=0.05))
        self.att = torch.nn.Parameter(torch.nn.init.trunc_normal_(torch.empty(1, heads, out_channels * 2), std=0.05))
        self.dropout = torch.nn.Dropout(dropout_rate)
        if add_bias:
            self.bias = torch.nn.Parameter(torch.zeros(heads * out_channels if concat else out_channels))

    def forward(self, x, edge_index, num_nodes=None):
        x = (x @ self.w).reshape(-1, self.heads, self.out_channels)         # gat_conv.py:99-100
        num_nodes = x.shape[0] if num_nodes is None else num_nodes
        node_src, node_dst = edge_index[0, :], edge_index[1, :]
        # e = leaky_relu(sum_c concat(x[src], x[dst]) * att): the per-node halves are reduced before the gather
        # (same arithmetic association per half; avoids materialising [E, H, 2C])
        a_src, a_dst = self.att[..., :self.out_channels], self.att[..., self.out_channels:]
        if (self.fused and self.fuse_attention and x.dtype == torch.float32 and num_nodes == x.shape[0]
                and mpops.gat_fusable_grouped(self.heads, self.out_channels) and not (self.training and self.dropout_rate > 0)):
            # logits, edge softmax and aggregation in one streaming pass per direction: no [E, H] tensor at all
            out = mpops.gat_aggregate(edge_index, x, (x * a_src).sum(-1), (x * a_dst).sum(-1), self.negetive_slop)
            out = out.reshape(-1, self.heads * self.out_channels) if self.concat else out.mean(dim=1)
            return out + self.bias if self.add_bias else out
        if self.fused:
            e = (x * a_src).sum(-1)[node_src] + (x * a_dst).sum(-1)[node_dst]
        else:
            feat = torch.cat((x[node_src], x[node_dst]), dim=-1)            # gat_conv.py:103-108
            e = (feat * self.att).sum(-1)
        e = torch.nn.functional.leaky_relu(e, self.negetive_slop)
        if self.fused:
            alpha = self.dropout(mpops.segment_softmax(e, node_dst, num_nodes))
        else:                                                               # utils/softmax.py:29-35
            mx = mpops.unsorted_segment_max(e, node_dst, num_nodes)
            ex = torch.exp(e - mx[node_dst])
            den = mpops.unsorted_segment_sum(ex, node_dst, num_nodes)
            alpha = self.dropout(ex / (den[node_dst] + 1e-16))
        x = self.propagate(x, edge_index, num_nodes=num_nodes, edge_weight=alpha)   # gat_conv.py:112 (bspmm when fused)
        x = x.reshape(-1, self.heads * self.out_channels) if self.concat else x.mean(dim=1)
        if self.add_bias:
            x = x + self.bias
        return x


class GATV2Conv(MessagePassing):
    """gatv2_conv.py:58-125.  The reference's logit is mean_c(leaky_relu(x_j) * att_src + leaky_relu(x_i) * att_dst): a sum of
    a source-node and a destination-node scalar with NO non-linearity after the sum, so the attention is the fused GAT
    kernel with slope 1 (identity) on s_src[j] + s_dst[i] -- no [E, H, C] gathers, no [E, H] tensor."""

    def __init__(self, in_channels, out_channels, heads=1, concat=True, negative_slope=0.2, dropout_rate=0., add_bias=True, fused=True):
        super().__init__(fused)
        self.in_channels, self.out_channels, self.heads, self.concat = in_channels, out_channels, heads, concat
        self.negetive_slop, self.dropout_rate, self.add_bias = negative_slope, dropout_rate, add_bias
        self.linear = Linear(in_channels, out_channels * heads)
        self.att_src = torch.nn.Parameter(torch.nn.init.trunc_normal_(torch.empty(1, heads, out_channels), std=0.05))
        self.att_dst = torch.nn.Parameter(torch.nn.init.trunc_normal_(torch.empty(1, heads, out_channels), std=0.05))
        self.dropout = torch.nn.Dropout(dropout_rate)
        if add_bias:
            self.bias = torch.nn.Parameter(torch.zeros(heads * out_channels if concat else out_channels))

    def forward(self, x, edge_index, num_nodes=None):
        x = self.linear(x).reshape(-1, self.heads, self.out_channels)       # gatv2_conv.py:113
        num_nodes = x.shape[0] if num_nodes is None else num_nodes
        lx = torch.nn.functional.leaky_relu(x, self.negetive_slop)
        s_src = (lx * self.att_src).mean(-1)                                # gatv2_conv.py:99-101, per node instead of per edge
        s_dst = (lx * self.att_dst).mean(-1)
        if (self.fused and x.dtype == torch.float32 and num_nodes == x.shape[0] and mpops.gat_fusable_grouped(self.heads, self.out_channels)
                and not (self.training and self.dropout_rate > 0)):
            out = mpops.gat_aggregate(edge_index, x, s_src, s_dst, 1.0)
        else:
            src, dst = edge_index[0], edge_index[1]
            alpha = self.dropout(mpops.segment_softmax(s_src[src] + s_dst[dst], dst, num_nodes))   # gatv2_conv.py:103-104
            out = self.propagate(x, edge_index, num_nodes=num_nodes, edge_weight=alpha)            # per-head weighted sum (bspmm)
        out = out.reshape(-1, self.heads * self.out_channels) if self.concat else out.mean(dim=1)
        return out + self.bias if self.add_bias else out


GATv2Conv = GATV2Conv      # the reference exports the class as GATV2Conv (layers/conv/gatv2_conv.py:9)


class AGNNConv(MessagePassing):
    """agnn_conv.py:40-73: alpha = softmax_i(beta * cos(x_j, x_i)), out_i = sum_j alpha x_j.  The cosine logits are an SDDMM
    over the normalised rows (mpops.sddmm: no [E, K] gathers), the weighted sum the one-head bspmm (gradient for alpha)."""

    def __init__(self, in_channels, require_grad=True, fused=True):
        super().__init__(fused)
        self.in_channels = in_channels
        if require_grad:
            self.beta = torch.nn.Parameter(torch.rand(1))
        else:
            self.register_buffer('beta', torch.zeros(1))

    def forward(self, x, edge_index, num_nodes=None):
        num_nodes = x.shape[0] if num_nodes is None else num_nodes
        src, dst = edge_index[0], edge_index[1]
        xn = torch.nn.functional.normalize(x, p=2.0, dim=-1, eps=1e-12)      # tlx.l2_normalize
        if self.fused and x.dtype == torch.float32 and num_nodes == x.shape[0] and x.dim() == 2:
            cos = mpops.sddmm(edge_index, xn.unsqueeze(1), xn.unsqueeze(1)).reshape(-1)
            alpha = mpops.segment_softmax(cos * self.beta, dst, num_nodes)
            return mpops.bspmm(edge_index, alpha.reshape(-1, 1), x.unsqueeze(1), 'sum').squeeze(1)
        cos = (xn[src] * xn[dst]).sum(-1)                                    # agnn_conv.py:64-66
        mpops.check_message_bytes(edge_index.shape[1], x[0].numel(), x.element_size(), 'AGNNConv (unfused)')
        alpha = mpops.segment_softmax(cos * self.beta, dst, num_nodes)
        return mpops.unsorted_segment_sum(alpha.unsqueeze(-1) * x[src], dst, num_nodes)


class GINConv(MessagePassing):
    def __init__(self, nn, eps=0., train_eps=False, fused=True):
        super().__init__(fused)
        self.nn = nn
        if train_eps:
            self.eps = torch.nn.Parameter(torch.tensor([float(eps)]))
        else:
            self.register_buffer('eps', torch.tensor([float(eps)]))

    def forward(self, x, edge_index, size=None):
        if not isinstance(x, (list, tuple)):
            x = (x, x)
        out = self.propagate(x=x[0], edge_index=edge_index)                 # gin_conv.py:54 (sum, no weights)
        if x[1] is not None:
            out = out + (1 + self.eps) * x[1]
        return self.nn(out)


class SGConv(MessagePassing):
    """sgc_conv.py:35-51: x = linear(x); iter_K times propagate(sum) on the same edge_index (one cached CSR serves all hops)."""

    def __init__(self, in_channels, out_channels, iter_K=2, fused=True):
        super().__init__(fused)
        self.in_channels, self.out_channels, self.iter_K = in_channels, out_channels, iter_K
        self.linear = Linear(in_channels, out_channels, bias=False)

    def forward(self, x, edge_index, edge_weight=None, num_nodes=None):
        x = self.linear(x)                                                  # sgc_conv.py:46
        for _ in range(self.iter_K):                                        # sgc_conv.py:47-48
            x = self.propagate(x, edge_index, edge_weight=edge_weight, num_nodes=num_nodes)
        return x


class APPNPConv(MessagePassing):
    """appnp_conv.py:9-37: personalised PageRank propagation x <- (1 - alpha) * A x + alpha * h0, iter_K times."""

    def __init__(self, in_channels, out_channels, iter_K, alpha, drop_rate, fused=True):
        super().__init__(fused)
        self.in_channels, self.out_channels, self.iter_K, self.alpha, self.drop_rate = in_channels, out_channels, iter_K, alpha, drop_rate
        self.linear_w = Linear(in_channels, out_channels, bias=False)
        self.dropout = torch.nn.Dropout(drop_rate)
        self.dropout_e = torch.nn.Dropout(0.9)                              # appnp_conv.py:23 (edge-weight dropout, training only)

    def forward(self, x, edge_index, edge_weight=None, num_nodes=None):
        x = self.linear_w(x)
        h0 = x
        x = self.dropout(x)
        for _ in range(self.iter_K):                                        # appnp_conv.py:30-36
            if edge_weight is not None:
                edge_weight = self.dropout_e(edge_weight)
            x = self.propagate(x, edge_index, edge_weight=edge_weight, num_nodes=num_nodes)
            x = x * (1 - self.alpha) + self.alpha * h0
        return x


# ---- global pools (layers/pool/glob.py:7-135): one segment reduction over the batch vector ----------------------------
def _pool_size(batch, size):
    return int(batch.max().item()) + 1 if size is None else int(size)


def global_sum_pool(x, batch, size=None):
    if batch is None:
        return x.sum(dim=0, keepdim=True)
    return mpops.unsorted_segment_sum(x, batch, _pool_size(batch, size))


def global_mean_pool(x, batch, size=None):
    if batch is None:
        return x.mean(dim=0, keepdim=True)
    return mpops.unsorted_segment_mean(x, batch, _pool_size(batch, size))


def global_max_pool(x, batch, size=None):
    if batch is None:
        return x.max(dim=0, keepdim=True)[0]
    return mpops.unsorted_segment_max(x, batch, _pool_size(batch, size))


def global_min_pool(x, batch, size=None):
    if batch is None:
        return x.min(dim=0, keepdim=True)[0]
    return -mpops.unsorted_segment_max(-x, batch, _pool_size(batch, size))
