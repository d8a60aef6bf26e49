"""GPU test (-m gpu): gammagl_b200.layers loaded with the parameters of the reference's own layers must reproduce the
reference layers' outputs and input gradients (tests/golden/layers_ref.npz: GCNConv / SAGEConv / GATConv / GINConv /
SGConv / APPNPConv executed unmodified on the reference's compiled CPU operators, make_golden_layers.py).
fp32 on both sides: 2e-5 relative to the largest output (1e-4 for GAT, whose fused kernel uses ex2-based exponentials)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
Z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'layers_ref.npz'))


def rel(a, b):
    a, b = a.detach().double().cpu().numpy(), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def load(layer, tag):
    """Copy every reference parameter `<tag>_p_<dotted name>` onto the attribute of the same name."""
    n = 0
    for key in Z.files:
        if not key.startswith(tag + '_p_'):
            continue
        obj = layer
        path = key[len(tag) + 3:].split('__')
        for a in path[:-1]:
            obj = getattr(obj, a, None)
        if obj is None or not hasattr(obj, path[-1]):
            continue                                   # GATConv.linear: allocated by the reference, never used (gat_conv.py:78-80,99)
        p = getattr(obj, path[-1])
        v = torch.from_numpy(Z[key])
        assert tuple(p.shape) == tuple(v.shape), (key, tuple(p.shape), tuple(v.shape))
        p.data = v.clone()
        n += 1
    assert n >= 1, tag
    return layer.to(DEV).eval()


def build(tag):
    from gammagl_b200 import layers as L
    F, H = Z['x'].shape[1], 20
    w = torch.from_numpy(Z['w']).to(DEV)
    if tag.startswith('gcn_'):
        norm = tag[4:]
        return (L.GCNConv(F, H, norm='both'), (w,)) if norm == 'weighted' else (L.GCNConv(F, H, norm=norm), ())
    if tag.startswith('sage_'):
        return L.SAGEConv(F, H, aggr=tag[5:]), ()
    if tag.startswith('gat_'):
        heads = int(tag.split('_')[1][1:])
        return L.GATConv(F, 8, heads=heads, concat=tag.endswith('cat')), ()
    if tag == 'gin':
        return L.GINConv(L.Linear(F, H), eps=0.25), ()
    if tag == 'sgc':
        return L.SGConv(F, H, iter_K=3), (w,)
    assert tag == 'appnp'
    return L.APPNPConv(F, H, iter_K=4, alpha=0.15, drop_rate=0.5), (w,)


TAGS = ['gcn_both', 'gcn_left', 'gcn_right', 'gcn_none', 'gcn_weighted', 'sage_mean', 'sage_gcn', 'sage_pool',
        'gat_h8_cat', 'gat_h4_avg', 'gat_h1_cat', 'gin', 'sgc', 'appnp']


@pytest.mark.parametrize('fused', [True, False], ids=['fused', 'message+aggregate'])
@pytest.mark.parametrize('tag', TAGS)
def test_layer_reproduces_reference_layer(tag, fused):
    layer, extra = build(tag)
    layer = load(layer, tag)
    layer.fused = fused
    ei = torch.from_numpy(Z['ei']).to(DEV)
    x = torch.from_numpy(Z['x']).to(DEV).requires_grad_(True)
    out = layer(x, ei, *extra)
    tol = 1e-4 if tag.startswith('gat') else 2e-5
    assert rel(out, Z[tag + '_out']) < tol
    out.backward(torch.from_numpy(Z[tag + '_gout']).to(DEV))
    assert rel(x.grad, Z[tag + '_gx']) < tol


@pytest.mark.parametrize('K', [16, 12])
def test_fused_propagate_keeps_the_edge_weight_gradient(K):
    """The reference layer's message = x[src] * w (message_passing.py:55-61) is differentiable in w; gspmm is not
    (src/gspmm.cpp:30,79).  The fused branch must still deliver d out / d w (one-head bspmm: same forward + SDDMM)."""
    from gammagl_b200.layers import MessagePassing
    N, E = 400, 6000
    g = torch.Generator().manual_seed(21)
    ei = torch.randint(0, N, (2, E), generator=g)
    x = torch.randn(N, K, generator=g)
    w = torch.rand(E, generator=g)
    go = torch.randn(N, K, generator=g)
    outs = []
    for fused in (True, False):
        mp = MessagePassing(fused=fused)
        xd = x.to(DEV).requires_grad_(True)
        wd = w.to(DEV).requires_grad_(True)
        out = mp.propagate(xd, ei.to(DEV), edge_weight=wd)
        out.backward(go.to(DEV))
        outs.append((out.detach(), xd.grad, wd.grad))
    x64 = x.double().requires_grad_(True)
    w64 = w.double().requires_grad_(True)
    ref = torch.zeros(N, K, dtype=torch.float64).index_add(0, ei[1], x64[ei[0]] * w64[:, None])
    ref.backward(go.double())
    for out, gx, gw in outs:
        assert gw is not None
        assert rel(out, ref.detach().numpy()) < 1e-5 and rel(gx, x64.grad.numpy()) < 1e-5 and rel(gw, w64.grad.numpy()) < 1e-5
    # mean with a differentiable weight takes the decomposition; without grad tracking the fused gspmm path is used
    mp = MessagePassing()
    wd = w.to(DEV).requires_grad_(True)
    mp.propagate(x.to(DEV), ei.to(DEV), edge_weight=wd, aggr='mean').sum().backward()
    assert wd.grad is not None and float(wd.grad.abs().sum()) > 0
    with torch.no_grad():
        a = mp.propagate(x.to(DEV), ei.to(DEV), edge_weight=wd)
    assert rel(a, ref.detach().numpy()) < 1e-5
