"""GPU twin of tests/test_plugin.py: plugin.enable() on a stand-in `gammagl` module tree that follows
the reference's propagate switch (message_passing.py:144) and Inspector contract (utils/inspector.py:90-107), CUDA tensors,
outputs and gradients against the CPU oracle for the call shapes of GCNConv / SAGEConv / GATConv / GINConv:
1-D weights, [E,H] weights, bipartite (src, dst), max, differentiable weights, and a layer built BEFORE enable()."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
import standin_gammagl

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.fixture()
def tree():
    from gammagl_b200 import plugin
    mods = standin_gammagl.make()
    early = mods['gammagl.layers.conv'].SAGEConv()          # constructed BEFORE enable(): its Inspector holds the base signature
    n = plugin.enable(modules=mods)
    assert n > 0
    yield mods, early
    plugin.disable()


def graph(N_src, N_dst, E, seed, split_row=True):
    g = torch.Generator().manual_seed(seed)
    ei = torch.stack([torch.randint(0, N_src, (E,), generator=g), torch.randint(0, N_dst, (E,), generator=g)])
    if split_row:
        ei[1, :1500] = 3                                      # one row above GGLB200_LONG_ROW
    return ei, g


def short_rows(ei, N):
    return np.bincount(ei[1].numpy(), minlength=N) <= 1024


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a, np.float64) - b)) / max(1e-30, np.max(np.abs(b))))


def test_gcn_call_shape_1d_weights(tree):
    mods, _ = tree
    N, E, K = 900, 20000, 100
    ei, g = graph(N, N, E, 1)
    x, w, gout = torch.randn(N, K, generator=g), torch.rand(E, generator=g), torch.randn(N, K, generator=g)
    layer = mods['gammagl.layers.conv'].GCNConv()
    xx = x.to(DEV).requires_grad_(True)
    out = layer(xx, ei.to(DEV), w.to(DEV), N)
    out.backward(gout.to(DEV))
    ref, _ = O.spmm_fwd('sum', ei.numpy(), w.numpy(), x.numpy())
    gref = O.spmm_bwd('sum', ei.numpy(), w.numpy(), gout.numpy())
    s = short_rows(ei, N)
    assert np.array_equal(out.detach().cpu().numpy()[s], ref[s])
    assert rel(out.detach().cpu().numpy(), ref) <= 1e-5
    assert np.array_equal(xx.grad.cpu().numpy(), gref)        # out-degrees here stay below the split threshold


def test_gat_call_shape_per_head_weights_through_segment_softmax(tree):
    mods, _ = tree
    N, E, H, C = 500, 12000, 4, 8
    ei, g = graph(N, N, E, 2)
    x, e, gout = torch.randn(N, H, C, generator=g), torch.randn(E, H, generator=g), torch.randn(N, H, C, generator=g)
    layer = mods['gammagl.layers.conv'].GATConv()
    xx, ee = x.to(DEV).requires_grad_(True), e.to(DEV).requires_grad_(True)
    out = layer(xx, ee, ei.to(DEV), N)
    out.backward(gout.to(DEV))
    alpha = O.segment_softmax(e.numpy(), ei[1].numpy(), N)
    ref = O.bspmm_fwd(ei.numpy(), alpha, x.numpy())
    gx_ref, galpha = O.bspmm_bwd(ei.numpy(), alpha, x.numpy(), gout.numpy())
    assert rel(out.detach().cpu().numpy(), ref) <= 2e-5
    assert rel(xx.grad.cpu().numpy(), gx_ref) <= 2e-5
    # softmax backward in fp64: ge = alpha * (galpha - sum_seg(alpha * galpha)[dst])
    a64, ga64 = alpha.astype(np.float64), galpha.astype(np.float64)
    ssum = np.zeros((N, H))
    np.add.at(ssum, ei[1].numpy(), a64 * ga64)
    ge = a64 * (ga64 - ssum[ei[1].numpy()])
    assert rel(ee.grad.cpu().numpy(), ge) <= 1e-4


@pytest.mark.parametrize('aggr', ['mean', 'max', 'sum'])
def test_sage_call_shape_bipartite(tree, aggr):
    mods, early = tree
    Ns, Nd, E, K = 700, 260, 9000, 32
    ei, g = graph(Ns, Nd, E, 3)
    xs, gout = torch.randn(Ns, K, generator=g), torch.randn(Nd, K, generator=g)
    for layer in (mods['gammagl.layers.conv'].SAGEConv(), early):      # `early` was built before enable(): re-inspected
        xx = xs.to(DEV).requires_grad_(True)
        out = layer((xx, xx[:Nd]), ei.to(DEV), aggr)
        assert tuple(out.shape) == (Nd, K)                              # rows = #destinations, not x rows
        out.backward(gout.to(DEV))
        msg = xs.numpy()[ei[0].numpy()]
        if aggr == 'max':
            ref, arg = O.segment('max', msg, ei[1].numpy(), Nd)
            gm = O.segment_bwd('max', gout.numpy(), ei[1].numpy(), E, arg)
            assert np.array_equal(out.detach().cpu().numpy(), ref)
        else:
            ref = O.segment(aggr, msg, ei[1].numpy(), Nd)
            gm = O.segment_bwd(aggr, gout.numpy(), ei[1].numpy(), E)
            assert rel(out.detach().cpu().numpy(), ref) <= 1e-5
        gx = np.zeros((Ns, K), dtype=np.float64)
        np.add.at(gx, ei[0].numpy(), gm.astype(np.float64))
        assert rel(xx.grad.cpu().numpy(), gx) <= 1e-5


def test_gin_call_shape_no_weights(tree):
    mods, _ = tree
    N, E, K = 800, 15000, 47
    ei, g = graph(N, N, E, 4)
    x, gout = torch.randn(N, K, generator=g), torch.randn(N, K, generator=g)
    xx = x.to(DEV).requires_grad_(True)
    out = mods['gammagl.layers.conv'].GINConv()(xx, ei.to(DEV))
    out.backward(gout.to(DEV))
    ones = np.ones(E, dtype=np.float32)
    ref, _ = O.spmm_fwd('sum', ei.numpy(), ones, x.numpy())
    s = short_rows(ei, N)
    assert np.array_equal(out.detach().cpu().numpy()[s], ref[s])
    assert np.array_equal(xx.grad.cpu().numpy(), O.spmm_bwd('sum', ei.numpy(), ones, gout.numpy()))


def test_differentiable_edge_weights_keep_their_gradient(tree):
    mods, _ = tree
    N, E, K = 600, 8000, 24
    ei, g = graph(N, N, E, 5, split_row=False)
    x, w, gout = torch.randn(N, K, generator=g), torch.rand(E, generator=g), torch.randn(N, K, generator=g)
    xx, ww = x.to(DEV).requires_grad_(True), w.to(DEV).requires_grad_(True)
    out = mods['gammagl.layers.conv'].GCNConv()(xx, ei.to(DEV), ww, N)
    out.backward(gout.to(DEV))
    ref, _ = O.spmm_fwd('sum', ei.numpy(), w.numpy(), x.numpy())
    assert np.array_equal(out.detach().cpu().numpy(), ref)
    gw = np.einsum('ek,ek->e', x.numpy()[ei[0].numpy()].astype(np.float64), gout.numpy()[ei[1].numpy()].astype(np.float64))
    assert rel(ww.grad.cpu().numpy(), gw) <= 1e-5
    assert np.array_equal(xx.grad.cpu().numpy(), O.spmm_bwd('sum', ei.numpy(), w.numpy(), gout.numpy()))


def test_unfused_layer_still_runs_on_the_b200_segment_operators(tree):
    mods, _ = tree
    N, E, K = 300, 4000, 16
    ei, g = graph(N, N, E, 6, split_row=False)
    x = torch.randn(N, K, generator=g)
    layer = mods['gammagl.layers.conv'].HGTConv()
    assert 'message_aggregate' not in type(layer).__dict__
    out = layer.propagate(x.to(DEV), ei.to(DEV))
    ref = O.segment('sum', 2.0 * x.numpy()[ei[0].numpy()], ei[1].numpy(), N)
    assert np.array_equal(out.cpu().numpy(), ref)


def _attention_reference(x, ei, logits, N):
    """fp64: softmax of `logits` [E, H] over incoming edges, then out[i] = sum alpha x[src]."""
    src, dst = ei
    H = logits.shape[1]
    mx = torch.full((N, H), -1e300, dtype=torch.float64).scatter_reduce(0, dst[:, None].expand(-1, H), logits, 'amax')
    ex = torch.exp(logits - mx[dst].detach())
    den = torch.zeros(N, H, dtype=torch.float64).index_add_(0, dst, ex)
    alpha = ex / (den[dst] + 1e-16)
    return torch.zeros((N,) + tuple(x.shape[1:]), dtype=torch.float64).index_add_(0, dst, alpha.reshape(alpha.shape + (1,) * (x.dim() - 2)) * x[src])


def test_gatv2_and_agnn_take_their_fused_operators(tree):
    mods, _ = tree
    from gammagl_b200 import mpops
    conv = mods['gammagl.layers.conv']
    N, E = 700, 14000
    ei, g = graph(N, N, E, 8)
    # ---- GATv2 ----
    H, C = 4, 16
    layer = conv.GATV2Conv(H, C).to(DEV)
    assert type(layer).message_aggregate.__name__ == 'message_aggregate'
    x = torch.randn(N, H, C, generator=g)
    gout = torch.randn(N, H, C, generator=g)
    called = []
    real = mpops.gat_aggregate
    mpops.gat_aggregate = lambda *a, **k: (called.append('gat'), real(*a, **k))[1]
    try:
        xx = x.to(DEV).requires_grad_(True)
        out = layer(xx, ei.to(DEV), N)
        out.backward(gout.to(DEV))
    finally:
        mpops.gat_aggregate = real
    assert called == ['gat']
    xd = x.double().requires_grad_(True)
    lr = torch.nn.functional.leaky_relu
    a_s, a_d = layer.att_src.detach().double().cpu(), layer.att_dst.detach().double().cpu()
    logits = (lr(xd[ei[0]], 0.2) * a_s + lr(xd[ei[1]], 0.2) * a_d).mean(-1)
    ref = _attention_reference(xd, ei, logits, N)
    ref.backward(gout.double())
    assert rel(out.detach().cpu().numpy(), ref.detach().numpy()) < 2e-5
    assert rel(xx.grad.cpu().numpy(), xd.grad.numpy()) < 1e-4
    # ---- AGNN ----
    K = 48
    layer = conv.AGNNConv().to(DEV)
    x = torch.randn(N, K, generator=g)
    gout = torch.randn(N, K, generator=g)
    xx = x.to(DEV).requires_grad_(True)
    out = layer(xx, ei.to(DEV), N)
    out.backward(gout.to(DEV))
    xd = x.double().requires_grad_(True)
    beta = layer.beta.detach().double().cpu().requires_grad_(True)
    nrm = lambda t: torch.nn.functional.normalize(t, p=2.0, dim=-1, eps=1e-12)
    logits = ((nrm(xd[ei[0]]) * nrm(xd[ei[1]])).sum(-1) * beta).unsqueeze(-1)
    ref = _attention_reference(xd, ei, logits, N)
    ref.backward(gout.double())
    assert rel(out.detach().cpu().numpy(), ref.detach().numpy()) < 2e-5
    assert rel(xx.grad.cpu().numpy(), xd.grad.numpy()) < 1e-4
    assert rel(layer.beta.grad.cpu().numpy(), beta.grad.numpy()) < 1e-4
