"""CPU test (-m "not gpu"; needs /root/reference, so it runs in the authoring container and is skipped on the GPU box):
the fused `message_aggregate` that gammagl_b200.plugin.enable() installs on the reference's layer classes is executed here
ON THE REFERENCE'S OWN LAYERS with the reference's own CPU operators (gspmm / bspmm of gammagl/mpops/torch.py on the
compiled torch_ext) standing in for the B200 ones.  For every class in plugin._FUSED_LAYERS that the tensorlayerx shim can
construct, the layer output and input gradient with the fused branch must equal the stock message + aggregate run:
  * the reference's own switch (message_passing.py:144) and Inspector.distribute accept the installed signature,
  * the call shapes of these layers (2-D x with 1-D weights, [N,H,C] with [E,H] weights, no weights, mean / max) are the
    ones the fused operator serves, and max / differentiable weights take the documented fall-backs."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

REF = '/root/reference/gammagl'
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason='needs the reference checkout (authoring container only)')
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope='module')
def ref():
    spec = importlib.util.spec_from_file_location('make_golden_layers', os.path.join(HERE, 'golden', 'make_golden_layers.py'))
    G = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(G)
    saved = {k: v for k, v in sys.modules.items() if k == 'gammagl' or k.startswith('gammagl.') or k.startswith('tensorlayerx')}
    L = G.load_reference_layers()
    for cls, f in (('GCNIIConv', 'gcnii_conv'), ('GPRConv', 'gpr_conv'), ('MixHopConv', 'mixhop_conv'), ('MAGCLConv', 'magcl_conv')):
        L[cls] = G.load_extra_layer(cls, f)
    from gammagl_b200 import mpops as b200_mpops, plugin
    ref_mpops = sys.modules['gammagl.mpops']

    def ref_gspmm(index, weight=None, x=None, reduce='sum', num_dst=None):
        assert num_dst is None or num_dst == x.shape[0]          # the reference operator is square-only
        return ref_mpops.gspmm(index, weight, x, reduce)

    def ref_gspmm_max_edge(index, weight=None, x=None, num_dst=None, return_arg=False):
        # the semantics the fused kernel reproduces: message + unsorted_segment_max of the reference itself
        msg = x[index[0]] if weight is None else x[index[0]] * weight.reshape((-1,) + (1,) * (x.dim() - 1))
        return ref_mpops.unsorted_segment_max(msg, index[1], x.shape[0] if num_dst is None else num_dst)

    class RefOps:                                   # the reference's CPU operators in the role of gammagl_b200.mpops
        gspmm = staticmethod(ref_gspmm)
        gspmm_max_edge = staticmethod(ref_gspmm_max_edge)
        bspmm = staticmethod(ref_mpops.bspmm)
        weight_needs_grad = staticmethod(b200_mpops.weight_needs_grad)
        check_message_bytes = staticmethod(b200_mpops.check_message_bytes)
    old = plugin.mpops
    plugin.mpops = RefOps
    yield G, L, plugin, sys.modules['tensorlayerx']
    plugin.mpops = old
    for k in [k for k in sys.modules if k == 'gammagl' or k.startswith('gammagl.') or k.startswith('tensorlayerx')]:
        del sys.modules[k]
    sys.modules.update(saved)


def cases(G, L, tlx):
    N, E, F, H = 120, 1500, 12, 12
    ei = G.graph(N, E, 1)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(N, F, generator=g)
    w = torch.rand(ei.shape[1], generator=g)
    out = [('GCNConv', lambda: L['GCNConv'](F, H), lambda m, x: m(x, ei)),
           ('GCNConv-weighted', lambda: L['GCNConv'](F, H, norm='left'), lambda m, x: m(x, ei, w)),
           ('SAGEConv-mean', lambda: L['SAGEConv'](F, H, aggr='mean'), lambda m, x: m(x, ei)),
           ('SAGEConv-gcn', lambda: L['SAGEConv'](F, H, aggr='gcn'), lambda m, x: m(x, ei)),
           ('SAGEConv-pool', lambda: L['SAGEConv'](F, H, aggr='pool'), lambda m, x: m(x, ei)),
           ('GATConv', lambda: L['GATConv'](F, 4, heads=3), lambda m, x: m(x, ei)),
           ('GINConv', lambda: L['GINConv'](tlx.nn.Linear(out_features=H, in_features=F)), lambda m, x: m(x, ei)),
           ('SGConv', lambda: L['SGConv'](F, H, iter_K=3), lambda m, x: m(x, ei, w)),
           ('APPNPConv', lambda: L['APPNPConv'](F, H, iter_K=3, alpha=0.1, drop_rate=0.3), lambda m, x: m(x, ei, w)),
           ('GCNIIConv', lambda: L['GCNIIConv'](F, F, alpha=0.1, beta=0.4), lambda m, x: m(x * 0.5, x, ei, w, N)),
           ('GCNIIConv-variant', lambda: L['GCNIIConv'](F, F, alpha=0.1, beta=0.4, variant=True), lambda m, x: m(x * 0.5, x, ei, w, N)),
           ('GPRConv', lambda: L['GPRConv'](K=4, alpha=0.2), lambda m, x: m(x, ei, w, N)),
           ('MixHopConv', lambda: L['MixHopConv'](F, H, p=[0, 1, 2]), lambda m, x: m(x, ei)),
           ('MAGCLConv', lambda: L['MAGCLConv'](F, H), lambda m, x: m(x, ei, 3))]
    return x, out


def run(build, call, x):
    torch.manual_seed(0)
    m = build()
    m.eval()
    xx = x.clone().requires_grad_(True)
    out = call(m, xx)
    gout = torch.randn(out.shape, generator=torch.Generator().manual_seed(5))
    out.backward(gout)
    return out.detach().numpy(), xx.grad.numpy()


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def test_fused_branch_on_reference_layers_equals_message_plus_aggregate(ref):
    G, L, plugin, tlx = ref
    x, cs = cases(G, L, tlx)
    names = set()
    for name, build, call in cs:
        cls = type(build())
        names.add(cls.__name__)
        assert cls.__name__ in plugin._FUSED_LAYERS
        assert 'message_aggregate' not in cls.__dict__
        o0, g0 = run(build, call, x)                                           # stock: message + aggregate
        cls.message_aggregate = plugin._fused_message_aggregate               # what plugin.enable() installs
        try:
            taken = []
            orig = plugin._fused_message_aggregate

            def message_aggregate(self, x, edge_index, edge_weight=None, aggr='sum', num_nodes=None):
                taken.append(1)
                return orig(self, x, edge_index, edge_weight=edge_weight, aggr=aggr, num_nodes=num_nodes)
            assert orig.__name__ == 'message_aggregate'                        # Inspector files signatures by __name__
            cls.message_aggregate = message_aggregate
            o1, g1 = run(build, call, x)
        finally:
            del cls.message_aggregate
        assert taken, name + ': the reference switch (message_passing.py:144) did not take the fused branch'
        assert rel(o1, o0) < 1e-5, name
        assert rel(g1, g0) < 1e-5, name
    # the four target layers and the multi-hop ones the shim can build (ChebConv and Hid need get_laplacian / cal_g_gradient;
    # their propagate() calls were checked by reading cheb_conv.py:117-121 and hid_conv.py:95-96)
    assert names >= {'GCNConv', 'SAGEConv', 'GATConv', 'GINConv', 'SGConv', 'APPNPConv', 'GCNIIConv', 'GPRConv', 'MixHopConv', 'MAGCLConv'}


def test_enable_rebinds_the_real_module_tree(ref):
    """plugin.enable() on the reference's real modules (loaded by path): every copy of the operators and of segment_softmax is
    rebound -- also the ones `from gammagl.utils import segment_softmax` left in gat_conv.py -- and disable() restores them."""
    G, L, plugin, tlx = ref
    from gammagl_b200 import mpops as b200
    saved_ops, plugin.mpops = plugin.mpops, b200
    mods = {k: v for k, v in sys.modules.items() if k == 'gammagl' or k.startswith('gammagl.')}
    before = {k: getattr(v, 'segment_softmax', None) for k, v in mods.items()}
    try:
        n = plugin.enable(modules=mods)
        assert n > 40
        assert sys.modules['gammagl.layers.conv.message_passing'].unsorted_segment_sum is b200.unsorted_segment_sum
        assert sys.modules['gammagl.layers.conv.message_passing'].gspmm is b200.gspmm
        assert sys.modules['gammagl.mpops.torch'].c_spmm_sum is b200.c_spmm_sum and sys.modules['gammagl.mpops'].use_ext is True
        for name in ('gammagl.utils.softmax', 'gammagl.utils', 'gammagl.layers.conv.gat_conv'):
            assert sys.modules[name].segment_softmax is b200.segment_softmax, name
        assert sys.modules['gammagl.utils.norm'].unsorted_segment_sum is b200.unsorted_segment_sum      # calc_gcn_norm
        assert sys.modules['gammagl.utils.degree'].unsorted_segment_sum is b200.unsorted_segment_sum    # degree
        for cls in ('GCNConv', 'SAGEConv', 'GATConv', 'GINConv', 'SGConv', 'APPNPConv', 'GCNIIConv', 'GPRConv', 'MixHopConv', 'MAGCLConv'):
            assert 'message_aggregate' in L[cls].__dict__, cls
        L['GCNConv'](4, 4)                                                     # constructing a patched layer must not raise
    finally:
        plugin.disable()
        plugin.mpops = saved_ops
    for k, v in mods.items():
        assert getattr(v, 'segment_softmax', None) is before[k], k
    assert 'message_aggregate' not in L['GCNConv'].__dict__
    assert sys.modules['gammagl.layers.conv.message_passing'].gspmm is not b200.gspmm


def test_fused_branch_keeps_the_edge_weight_gradient_of_the_reference_layer(ref):
    """SGConv forwards a user edge weight straight into propagate (sgc_conv.py:45-49); the reference layer's message * weight
    is differentiable in it.  With the fused branch installed (one-head bspmm for sum) d out / d w must equal the stock one."""
    G, L, plugin, tlx = ref
    N, E, F, H = 90, 900, 8, 8
    ei = G.graph(N, E, 3)
    g = torch.Generator().manual_seed(4)
    x = torch.randn(N, F, generator=g)
    w0 = torch.rand(ei.shape[1], generator=g)
    gout = torch.randn(N, H, generator=g)

    def grads():
        m = L['SGConv'](F, H, iter_K=2)
        w = w0.clone().requires_grad_(True)
        xx = x.clone().requires_grad_(True)
        out = m(xx, ei, w)
        out.backward(gout)
        return out.detach().numpy(), xx.grad.numpy(), w.grad.numpy()
    o0, gx0, gw0 = grads()
    L['SGConv'].message_aggregate = plugin._fused_message_aggregate
    try:
        o1, gx1, gw1 = grads()
    finally:
        del L['SGConv'].message_aggregate
    assert float(np.abs(gw0).sum()) > 0
    assert rel(o1, o0) < 1e-5 and rel(gx1, gx0) < 1e-5 and rel(gw1, gw0) < 1e-5


def test_special_layers_accept_the_installed_signature(ref):
    """GATv2Conv / AGNNConv of the reference (their own message()): the class-specific message_aggregate that enable() installs is
    routed by the reference's switch and Inspector, and on CPU tensors (no fused operator) reproduces the stock result."""
    G, L, plugin, tlx = ref
    N, E = 90, 700
    ei = G.graph(N, E, 3)
    g = torch.Generator().manual_seed(4)
    for cls_name, fname, build, call, xshape in (
            ('GATV2Conv', 'gatv2_conv', lambda c: c(12, 5, heads=3), lambda m, x: m(x, ei, N), (N, 12)),
            ('AGNNConv', 'agnn_conv', lambda c: c(12, require_grad=False), lambda m, x: m(x, ei, N), (N, 12))):
        cls = G.load_extra_layer(cls_name, fname)
        assert 'message_aggregate' not in cls.__dict__
        x = torch.randn(*xshape, generator=g)
        torch.manual_seed(0)
        m = build(cls)
        m.eval() if hasattr(m, 'eval') else None
        o0 = call(m, x).detach().numpy()
        cls.message_aggregate = plugin._SPECIAL_LAYERS[cls_name]
        try:
            torch.manual_seed(0)
            m1 = build(cls)
            m1.eval() if hasattr(m1, 'eval') else None
            o1 = call(m1, x).detach().numpy()
        finally:
            del cls.message_aggregate
        assert rel(o1, o0) < 1e-6, cls_name
