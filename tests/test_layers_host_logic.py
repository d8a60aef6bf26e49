"""CPU test of the HOST logic of gammagl_b200.layers (the mirror of MessagePassing / GCN / SAGE / GAT / GIN / SGC / APPNP):
the operators of gammagl_b200.mpops are replaced by plain torch CPU stand-ins with the reference operators' semantics
(gspmm without a weight gradient, segment max = first-wins on distinct values, ...), and every layer case of
tests/golden/layers_ref.npz -- outputs and input gradients of the reference's OWN layers -- must be reproduced, through the
fused branch and through message + aggregate.  This is the CPU twin of tests/test_gpu_zz_reference_layers.py: it checks the
layer algebra, parameter layout, routing (fuse / no fuse for max, differentiable edge weights, bipartite fall-back); the
kernels behind the operators are covered by -m gpu."""
import os

import numpy as np
import pytest
import torch

from gammagl_b200 import layers as L
from gammagl_b200 import mpops

Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'layers_ref.npz'))
TAGS = ['gcn_both', 'gcn_left', 'gcn_right', 'gcn_none', 'gcn_weighted', 'sage_mean', 'sage_gcn', 'sage_pool',
        'gat_h8_cat', 'gat_h4_avg', 'gat_h1_cat', 'gin', 'sgc', 'appnp']


def _bshape(ids, x):
    return ids.reshape((-1,) + (1,) * (x.dim() - 1)).expand_as(x)


def seg_sum(x, ids, n=None):
    n = int(ids.max()) + 1 if n is None else int(n)
    return torch.zeros((n,) + tuple(x.shape[1:]), dtype=x.dtype).index_add(0, ids, x)


def seg_mean(x, ids, n=None):
    n = int(ids.max()) + 1 if n is None else int(n)
    cnt = torch.bincount(ids, minlength=n).clamp_min(1).to(x.dtype)
    return seg_sum(x, ids, n) / cnt.reshape((-1,) + (1,) * (x.dim() - 1))


def seg_max(x, ids, n=None):
    n = int(ids.max()) + 1 if n is None else int(n)
    out = torch.full((n,) + tuple(x.shape[1:]), torch.finfo(x.dtype).min, dtype=x.dtype)
    return out.scatter_reduce(0, _bshape(ids, x), x, 'amax', include_self=True)


def gspmm(index, weight=None, x=None, reduce='sum', num_dst=None):
    msg = x[index[0]]
    if weight is not None:
        msg = msg * weight.detach().reshape((-1,) + (1,) * (x.dim() - 1))     # no gradient for the weight (src/gspmm.cpp:30,79)
    return {'sum': seg_sum, 'mean': seg_mean, 'max': seg_max}[reduce](msg, index[1], x.shape[0] if num_dst is None else num_dst)


class _FirstEdgeMax(torch.autograd.Function):
    """message + unsorted_segment_max with the reference's gradient: the FIRST maximal edge of (dst, k) gets it."""

    @staticmethod
    def forward(ctx, msg, ids, n):
        E = msg.shape[0]
        out = torch.full((n,) + tuple(msg.shape[1:]), torch.finfo(msg.dtype).min, dtype=msg.dtype)
        arg = torch.full(out.shape, E, dtype=torch.int64)
        for e in range(E):                                                      # serial loop of cpu/segment_max_cpu.cpp:54-62
            r = int(ids[e])
            better = out[r] < msg[e]
            out[r] = torch.where(better, msg[e], out[r])
            arg[r] = torch.where(better, torch.full_like(arg[r], e), arg[r])
        ctx.save_for_backward(arg)
        ctx.E = E
        return out

    @staticmethod
    def backward(ctx, g):
        (arg,) = ctx.saved_tensors
        gm = torch.zeros((ctx.E + 1,) + tuple(g.shape[1:]), dtype=g.dtype)
        gm.scatter_(0, arg, g)
        return gm[:-1], None, None


def gspmm_max_edge(index, weight=None, x=None, num_dst=None, return_arg=False):
    msg = x[index[0]]
    if weight is not None:
        msg = msg * weight.detach().reshape((-1,) + (1,) * (x.dim() - 1))
    return _FirstEdgeMax.apply(msg, index[1], x.shape[0] if num_dst is None else num_dst)


def bspmm(index, weight=None, x=None, reduce='sum'):
    return seg_sum(x[index[0]] * weight.unsqueeze(-1), index[1], x.shape[0])


def segment_softmax(data, ids, n):
    mx = seg_max(data, ids, n)
    ex = torch.exp(data - mx[ids].detach())
    return ex / (seg_sum(ex, ids, n)[ids] + 1e-16)


def gcn_propagate(edge_index, x, norm='both', bias=None, relu=False):
    N = x.shape[0]
    src, dst = edge_index
    w = torch.ones(edge_index.shape[1], dtype=x.dtype)
    p = -0.5 if norm in ('both', 'both_src') else -1.0
    if norm == 'both_src':
        d = torch.bincount(src, minlength=N).to(x.dtype)
        dis = torch.where(d > 0, d.pow(p), torch.zeros_like(d))
        w = w * dis[src] * dis[dst]
    if norm in ('both', 'left'):
        d = torch.bincount(src, minlength=N).to(x.dtype)
        w = w * torch.where(d > 0, d.pow(p), torch.zeros_like(d))[src]
    if norm in ('both', 'right'):
        d = torch.bincount(dst, minlength=N).to(x.dtype)
        w = w * torch.where(d > 0, d.pow(p), torch.zeros_like(d))[dst]
    out = seg_sum(x[src] * w[:, None], dst, N)
    if bias is not None:
        out = out + bias.reshape(1, -1)
    return torch.relu(out) if relu else out


def gat_aggregate(edge_index, x, s_src, s_dst, negative_slope=0.2):
    e = torch.nn.functional.leaky_relu(s_src[edge_index[0]] + s_dst[edge_index[1]], negative_slope)
    return bspmm(edge_index, segment_softmax(e, edge_index[1], x.shape[0]), x)


@pytest.fixture()
def host(monkeypatch):
    for name, fn in (('unsorted_segment_sum', seg_sum), ('unsorted_segment_mean', seg_mean), ('unsorted_segment_max', seg_max),
                     ('gspmm', gspmm), ('gspmm_max_edge', gspmm_max_edge), ('bspmm', bspmm), ('segment_softmax', segment_softmax), ('gcn_propagate', gcn_propagate),
                     ('gat_aggregate', gat_aggregate)):
        monkeypatch.setattr(mpops, name, fn)
    monkeypatch.setattr(L, 'degree', lambda index, num_nodes=None, dtype=torch.float32:
                        torch.bincount(index, minlength=int(index.max()) + 1 if num_nodes is None else num_nodes).to(dtype))
    return L


def build(tag):
    F, H = Z['x'].shape[1], 20
    w = torch.from_numpy(Z['w'])
    if tag.startswith('gcn_'):
        norm = tag[4:]
        return (L.GCNConv(F, H, norm='both'), (w,)) if norm == 'weighted' else (L.GCNConv(F, H, norm=norm), ())
    if tag.startswith('sage_'):
        return L.SAGEConv(F, H, aggr=tag[5:]), ()
    if tag.startswith('gat_'):
        return L.GATConv(F, 8, heads=int(tag.split('_')[1][1:]), concat=tag.endswith('cat')), ()
    if tag == 'gin':
        return L.GINConv(L.Linear(F, H), eps=0.25), ()
    if tag == 'sgc':
        return L.SGConv(F, H, iter_K=3), (w,)
    return L.APPNPConv(F, H, iter_K=4, alpha=0.15, drop_rate=0.5), (w,)


def load(layer, tag):
    for key in Z.files:
        if key.startswith(tag + '_p_'):
            obj, path = layer, key[len(tag) + 3:].split('__')
            for a in path[:-1]:
                obj = getattr(obj, a, None)
            if obj is not None and hasattr(obj, path[-1]):
                getattr(obj, path[-1]).data = torch.from_numpy(Z[key]).clone()
    return layer.eval()


def rel(a, b):
    a, b = a.detach().double().numpy(), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize('fused', [True, False], ids=['fused', 'message+aggregate'])
@pytest.mark.parametrize('tag', TAGS)
def test_layer_mirror_reproduces_reference_layer(host, tag, fused):
    layer, extra = build(tag)
    layer = load(layer, tag)
    layer.fused = fused
    ei = torch.from_numpy(Z['ei'])
    x = torch.from_numpy(Z['x']).requires_grad_(True)
    out = layer(x, ei, *extra)
    assert rel(out, Z[tag + '_out']) < 2e-5
    out.backward(torch.from_numpy(Z[tag + '_gout']))
    assert rel(x.grad, Z[tag + '_gx']) < 2e-5


def test_routing_of_propagate(host):
    """sum / mean / max fuse (max through the edge-id form, bipartite calls through num_dst); a weight that needs grad is never
    dropped: sum goes to bspmm, everything else to the size-guarded decomposition."""
    calls = []
    names = ('gspmm', 'gspmm_max_edge', 'bspmm', 'unsorted_segment_max', 'unsorted_segment_sum', 'unsorted_segment_mean')
    real = {k: getattr(mpops, k) for k in names}
    for k, f in real.items():
        setattr(mpops, k, (lambda k, f: lambda *a, **kw: (calls.append(k), f(*a, **kw))[1])(k, f))
    try:
        mp = L.MessagePassing()
        ei = torch.tensor([[0, 1, 2, 2, 3], [1, 2, 1, 0, 3]])
        x = torch.arange(8, dtype=torch.float32).reshape(4, 2)
        w = torch.rand(5)
        for aggr, want in (('sum', 'gspmm'), ('mean', 'gspmm'), ('max', 'gspmm_max_edge')):
            calls.clear()
            mp.propagate(x, ei, aggr=aggr, edge_weight=w)
            assert calls == [want], (aggr, calls)
        mp.fuse_max = False
        calls.clear(); mp.propagate(x, ei, aggr='max'); assert calls == ['unsorted_segment_max']
        mp.fuse_max = True
        wg = w.clone().requires_grad_(True)
        calls.clear(); out = mp.propagate(x, ei, edge_weight=wg); assert calls == ['bspmm']
        out.sum().backward(); assert wg.grad is not None and float(wg.grad.abs().sum()) > 0
        calls.clear(); mp.propagate(x, ei, edge_weight=wg, aggr='mean'); assert calls == ['unsorted_segment_mean']
        calls.clear(); mp.propagate(x, ei, edge_weight=wg, aggr='max'); assert calls == ['unsorted_segment_max']
        with torch.no_grad():
            calls.clear(); mp.propagate(x, ei, edge_weight=wg); assert calls == ['gspmm']
        calls.clear(); o6 = mp.propagate(x, ei, num_nodes=6); assert calls == ['gspmm'] and o6.shape[0] == 6   # bipartite: fused, 6 rows
        calls.clear(); o6 = mp.propagate(x, ei, num_nodes=6, aggr='max'); assert calls == ['gspmm_max_edge'] and o6.shape[0] == 6
        calls.clear(); mp.propagate(x, ei, num_nodes=6, edge_weight=wg); assert calls == ['unsorted_segment_sum']   # bspmm is square-only
        # the decomposition is size-guarded: a message above the limit raises instead of being allocated
        old = mpops.MAX_MESSAGE_BYTES
        mpops.MAX_MESSAGE_BYTES = 16
        try:
            with pytest.raises(Exception, match='message tensor'):
                mp.propagate(x, ei, edge_weight=wg, aggr='mean')
        finally:
            mpops.MAX_MESSAGE_BYTES = old
        # reference KAT (tests/layers/conv/test_message_passing.py:12-24)
        assert mp.propagate(x, ei, aggr='sum').tolist() == [[4.0, 5.0], [4.0, 6.0], [2.0, 3.0], [6.0, 7.0]]
        assert mp.propagate(x, ei, aggr='mean').tolist() == [[4.0, 5.0], [2.0, 3.0], [2.0, 3.0], [6.0, 7.0]]
        assert mp.propagate(x, ei, aggr='max').tolist() == [[4.0, 5.0], [4.0, 5.0], [2.0, 3.0], [6.0, 7.0]]
    finally:
        for k, f in real.items():
            setattr(mpops, k, f)


def test_sage_mean_aggregate_first_matches_reference_order(host):
    """SAGEConv(aggr='mean', aggregate_first=True): the neighbour Linear after the mean (n_dst rows) instead of before it (n_src rows),
    homogeneous call and bipartite block with the FULL feature matrix + global source ids -- same output and parameter gradients."""
    g = torch.Generator().manual_seed(3)
    N, n_src, n_dst, E, F, H = 60, 25, 7, 90, 6, 5
    x = torch.randn(N, F, generator=g)
    n_id = torch.randperm(N, generator=g)[:n_src]
    src, dst = torch.randint(0, n_src, (E,), generator=g), torch.randint(0, n_dst, (E,), generator=g)
    torch.manual_seed(0)
    a = L.SAGEConv(F, H, aggr='mean')
    b = L.SAGEConv(F, H, aggr='mean', aggregate_first=True)
    b.load_state_dict(a.state_dict())
    xs = x[n_id]
    ref = a((xs, xs[:n_dst]), torch.stack([src, dst]))
    got = b((x, x[n_id[:n_dst]]), torch.stack([n_id[src], dst]))
    assert ref.shape == got.shape == (n_dst, H) and rel(got, ref.detach().numpy()) < 1e-5
    go = torch.randn(n_dst, H, generator=g)
    ref.backward(go); got.backward(go)
    for (na, pa), (_, pb) in zip(a.named_parameters(), b.named_parameters()):
        assert rel(pb.grad, pa.grad.numpy()) < 1e-5, na
    ei = torch.stack([torch.randint(0, N, (200,), generator=g), torch.randint(0, N, (200,), generator=g)])
    assert rel(b(x, ei), a(x, ei).detach().numpy()) < 1e-5


def test_sage_gcn_self_loop_list_is_derived_once(host):
    """The self-looped edge list of SAGEConv(aggr='gcn') is memoised per edge_index (csr.derived_of) only for CUDA tensors; on the CPU
    every call derives it afresh -- and both the folded and the literal form give the reference's numbers (fixture test above)."""
    from gammagl_b200 import csr
    csr.clear_caches()
    layer = L.SAGEConv(4, 3, aggr='gcn')
    ei = torch.tensor([[0, 1, 2, 3], [1, 2, 3, 0]])
    x = torch.randn(4, 4)
    a, b = layer(x, ei), layer(x, ei)
    assert torch.equal(a, b) and len(csr._derived.d) == 0
