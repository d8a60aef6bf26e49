"""Dense fp64 restatements (plain torch, CPU) of the reference layers' forward algebra -- the layer oracles.

    gcn   /root/reference/gammagl/layers/conv/gcn_conv.py:78-108
    sage  /root/reference/gammagl/layers/conv/sage_conv.py:56-108   (+ utils/loop.py:107-150, utils/norm.py:5-30)
    gat   /root/reference/gammagl/layers/conv/gat_conv.py:98-122    (+ utils/softmax.py:29-35)
    gin   /root/reference/gammagl/layers/conv/gin_conv.py:49-63
    sgc   /root/reference/gammagl/layers/conv/sgc_conv.py:45-49
    appnp /root/reference/gammagl/layers/conv/appnp_conv.py:27-37   (eval mode)

TEST INFRASTRUCTURE ONLY.  tests/test_oracle_layers.py pins every function here to outputs of the reference layers
themselves (tests/golden/layers_ref.npz); the GPU layer tests compare gammagl_b200.layers against the same fixtures.
All functions take fp64 tensors, `ei` int64 [2, E], and return the layer output [N, ...] in fp64 (autograd-capable).
"""
import torch


def adj(ei, w, N):
    """A[dst, src] += w  (aggregation over destinations, message_passing.py:84)."""
    A = torch.zeros(N, N, dtype=torch.float64)
    return A.index_put((ei[1], ei[0]), w.double(), accumulate=True)


def counts(index, N):
    return torch.bincount(index, minlength=N).double()


def gcn(x, ei, W, bias, norm='both', edge_weight=None):
    N = x.shape[0]
    src, dst = ei[0], ei[1]
    w = torch.ones(ei.shape[1], dtype=torch.float64) if edge_weight is None else edge_weight.double()
    if norm in ('left', 'both'):
        d = counts(src, N)
        w = (d.pow(-0.5) if norm == 'both' else 1.0 / d)[src] * w
    if norm in ('right', 'both'):
        d = counts(dst, N)
        w = w * (d.pow(-0.5) if norm == 'both' else 1.0 / d)[dst]
    out = adj(ei, w, N) @ (x @ W)
    return out if bias is None else out + bias


def sage(x, ei, aggr, Wn, Ws, bias, Wp=None):
    N = x.shape[0]
    if aggr == 'mean':
        c = counts(ei[1], N).clamp_min(1.0)[:, None]
        out = (adj(ei, torch.ones(ei.shape[1]), N) @ (x @ Wn)) / c + x @ Ws
    elif aggr == 'gcn':
        n = int(ei[0].max()) + 1
        loops = torch.arange(n)
        e2 = torch.cat([ei, torch.stack([loops, loops])], dim=1)               # add_self_loops appends (i, i) for i < n
        d = counts(e2[0], n).pow(-0.5)                                          # calc_gcn_norm: degree over SOURCES for both ends
        out = (adj(e2, d[e2[0]] * d[e2[1]], n) @ (x @ Wn)[:n])[:N]
    else:
        h = torch.relu(x @ Wp)
        m = torch.full((N, h.shape[1]), -float('inf'), dtype=torch.float64)
        m = m.scatter_reduce(0, ei[1][:, None].expand(-1, h.shape[1]), h[ei[0]], 'amax')
        out = m @ Wn + x @ Ws
    return out if bias is None else out + bias


def segment_softmax(e, dst, N):
    mx = torch.full((N,) + tuple(e.shape[1:]), -float('inf'), dtype=torch.float64)
    mx = mx.scatter_reduce(0, dst.reshape((-1,) + (1,) * (e.dim() - 1)).expand_as(e), e, 'amax')
    ex = torch.exp(e - mx[dst].detach())
    den = torch.zeros_like(mx).index_add(0, dst, ex)
    return ex / (den[dst] + 1e-16)


def gat(x, ei, w, att, bias, heads, out_channels, concat=True, slope=0.2):
    N = x.shape[0]
    xh = (x @ w).reshape(N, heads, out_channels)
    feat = torch.cat([xh[ei[0]], xh[ei[1]]], dim=-1)
    e = torch.nn.functional.leaky_relu((feat * att).sum(-1), slope)
    alpha = segment_softmax(e, ei[1], N)
    out = torch.zeros(N, heads, out_channels, dtype=torch.float64).index_add(0, ei[1], xh[ei[0]] * alpha[..., None])
    out = out.reshape(N, heads * out_channels) if concat else out.mean(dim=1)
    return out if bias is None else out + bias


def gin(x, ei, W, eps):
    N = x.shape[0]
    return (adj(ei, torch.ones(ei.shape[1]), N) @ x + (1 + eps) * x) @ W


def sgc(x, ei, W, K, edge_weight=None):
    N = x.shape[0]
    A = adj(ei, torch.ones(ei.shape[1]) if edge_weight is None else edge_weight, N)
    h = x @ W
    for _ in range(K):
        h = A @ h
    return h


def appnp(x, ei, W, K, alpha, edge_weight=None):
    N = x.shape[0]
    A = adj(ei, torch.ones(ei.shape[1]) if edge_weight is None else edge_weight, N)
    h0 = x @ W
    h = h0
    for _ in range(K):
        h = (A @ h) * (1 - alpha) + alpha * h0
    return h


def cases(z):
    """Every layer case of tests/golden/layers_ref.npz as (tag, fn(x) -> dense output) with the fixture's parameters."""
    ei = torch.from_numpy(z['ei'])
    w = torch.from_numpy(z['w'])

    def P(tag, name):
        return torch.from_numpy(z['%s_p_%s' % (tag, name)]).double()
    out = []
    for norm in ('both', 'left', 'right', 'none'):
        t = 'gcn_' + norm
        out.append((t, lambda x, t=t, norm=norm: gcn(x, ei, P(t, 'linear__weights'), P(t, 'bias'), norm)))
    out.append(('gcn_weighted', lambda x: gcn(x, ei, P('gcn_weighted', 'linear__weights'), P('gcn_weighted', 'bias'), 'both', w)))
    for aggr in ('mean', 'gcn', 'pool'):
        t = 'sage_' + aggr
        out.append((t, lambda x, t=t, aggr=aggr: sage(x, ei, aggr, P(t, 'fc_neigh__weights'),
                                                       P(t, 'fc_self__weights') if aggr != 'gcn' else None, P(t, 'bias'),
                                                       P(t, 'pool__weights') if aggr == 'pool' else None)))
    for heads, concat in ((8, True), (4, False), (1, True)):
        t = 'gat_h%d_%s' % (heads, 'cat' if concat else 'avg')
        out.append((t, lambda x, t=t, heads=heads, concat=concat: gat(x, ei, P(t, 'w'), P(t, 'att'), P(t, 'bias'), heads, 8, concat)))
    out.append(('gin', lambda x: gin(x, ei, P('gin', 'nn__weights'), 0.25)))
    out.append(('sgc', lambda x: sgc(x, ei, P('sgc', 'linear__weights'), 3, w)))
    out.append(('appnp', lambda x: appnp(x, ei, P('appnp', 'linear_w__weights'), 4, 0.15, w)))
    return out
