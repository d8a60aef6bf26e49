"""Generate tests/golden/layers_*.npz by RUNNING THE REFERENCE's own layers -- GCNConv, SAGEConv, GATConv, GINConv,
SGConv, APPNPConv (/root/reference/gammagl/layers/conv/*.py, executed unmodified) on the reference's own compiled CPU
operators (oracle/_ref/_torch_ext.so) -- in the authoring container only:

    python tests/golden/make_golden_layers.py

The layers are written against `tensorlayerx`, which is not installed here; what they use of it on this path is a thin
layer over torch, supplied below as a test-only shim module (Module/_get_weights, Linear with [in, out]
weights, Dropout, LeakyReLU, ReLU and ~20 tensor helpers).  The reference files are loaded by path into a stand-in
`gammagl` package tree so that the heavy `gammagl/__init__.py` (datasets, loaders, ...) is not executed.

Each fixture holds the inputs, every parameter (named as in the reference) and the layer output AND input gradient
(`d sum(out * gout) / dx`) in fp32.  tests/test_oracle_layers.py pins the dense fp64 restatements of the layer algebra
that tests/test_gpu_layers.py uses as layer oracles to these fixtures; the same GPU test file loads the parameters into
gammagl_b200.layers and compares with the reference outputs directly.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, '..', '..'))
REF = '/root/reference/gammagl'
sys.path.insert(0, ROOT)


# ----------------------------------------------------------------------------------------------------------------------
# tensorlayerx shim (torch backend semantics of the calls the six layers make)
# ----------------------------------------------------------------------------------------------------------------------
def make_tlx():
    tlx = types.ModuleType('tensorlayerx')
    tlx.BACKEND = 'torch'
    tlx.float32, tlx.int64 = torch.float32, torch.int64

    class Module(torch.nn.Module):
        def _get_weights(self, var_name, shape, init=None, trainable=True, order=False):
            return torch.nn.Parameter(init(shape) if callable(init) else torch.zeros(shape))   # registered by `self.<name> = ...`

        @property
        def trainable_weights(self):
            return list(self.parameters())

    class Linear(Module):
        def __init__(self, out_features, in_features=None, W_init=None, b_init=None, act=None, name=None):
            super().__init__()
            g = torch.Generator().manual_seed(out_features * 7919 + in_features)
            self.weights = torch.nn.Parameter(torch.randn(in_features, out_features, generator=g) / in_features ** 0.5)
            self.biases = None if b_init is None else torch.nn.Parameter(torch.zeros(out_features))

        def forward(self, x):
            y = torch.matmul(x, self.weights)
            return y if self.biases is None else y + self.biases

    class Dropout(Module):
        def __init__(self, p=0.5):
            super().__init__()
            self.p = p

        def forward(self, x):
            return torch.nn.functional.dropout(x, self.p, self.training)

    class LeakyReLU(Module):
        def __init__(self, negative_slope=0.01):
            super().__init__()
            self.s = negative_slope

        def forward(self, x):
            return torch.nn.functional.leaky_relu(x, self.s)

    class ReLU(Module):
        def forward(self, x):
            return torch.relu(x)

    def _init(fn):
        return lambda *a, **k: (lambda shape, dtype=torch.float32: fn(tuple(shape)))
    inits = types.ModuleType('tensorlayerx.initializers')
    inits.Zeros = inits.zeros = _init(lambda s: torch.zeros(s))
    inits.TruncatedNormal = inits.he_normal = inits.XavierUniform = _init(
        lambda s: torch.randn(s, generator=torch.Generator().manual_seed(int(np.prod(s)))) * 0.1)
    nn = types.ModuleType('tensorlayerx.nn')
    nn.Module, nn.Linear, nn.Dropout, nn.LeakyReLU, nn.ReLU, nn.Parameter = Module, Linear, Dropout, LeakyReLU, ReLU, torch.nn.Parameter
    nn.layers = nn
    tlx.nn = tlx.layers = nn
    tlx.initializers = inits

    def gather(x, indices, axis=0):
        return torch.index_select(x, axis, torch.as_tensor(indices, dtype=torch.int64).reshape(-1))
    ops = types.ModuleType('tensorlayerx.ops')
    ops.gather = gather
    ops.is_tensor = lambda x: isinstance(x, torch.Tensor)
    tlx.ops = ops
    tlx.gather = gather
    tlx.is_tensor = ops.is_tensor
    tlx.ones = lambda shape, dtype=torch.float32, device=None: torch.ones(tuple(shape), dtype=dtype)
    tlx.zeros = lambda shape, dtype=torch.float32, device=None: torch.zeros(tuple(shape), dtype=dtype)
    tlx.get_tensor_shape = lambda x: list(x.shape)
    tlx.reshape = lambda x, shape: torch.reshape(x, tuple(shape))
    tlx.pow = torch.pow
    tlx.exp = torch.exp
    tlx.matmul = torch.matmul
    tlx.concat = lambda xs, axis=0: torch.cat(list(xs), dim=axis)
    tlx.stack = lambda xs, axis=0: torch.stack(list(xs), dim=axis)
    tlx.reduce_sum = lambda x, axis=None: x.sum() if axis is None else x.sum(dim=axis)
    tlx.l2_normalize = lambda x, axis=None, eps=1e-12: torch.nn.functional.normalize(x, p=2.0, dim=-1 if axis is None else axis, eps=eps)
    tlx.reduce_mean = lambda x, axis=None: x.mean() if axis is None else x.mean(dim=axis)
    tlx.reduce_max = lambda x, axis=None: x.max() if axis is None else x.max(dim=axis)[0]
    tlx.arange = lambda start, limit=None, delta=1, dtype=torch.int64: torch.arange(start, limit, delta, dtype=dtype)
    tlx.expand_dims = lambda x, axis: torch.unsqueeze(x, axis)
    tlx.convert_to_tensor = lambda x, dtype=None: torch.as_tensor(np.asarray(x) if not isinstance(x, torch.Tensor) else x, dtype=dtype)
    tlx.convert_to_numpy = lambda x: x.detach().cpu().numpy()
    tlx.cast = lambda x, dtype=None: x if dtype is None or x.dtype == dtype else x.to(dtype)
    tlx.zeros_like = lambda x, dtype=None: torch.zeros_like(x, dtype=dtype)
    tlx.constant = lambda value, dtype=torch.float32, shape=None, device=None: torch.full(tuple(shape), value, dtype=dtype)
    inits.Constant = lambda value=0.0: (lambda shape, dtype=torch.float32: torch.as_tensor(np.asarray(value), dtype=dtype).reshape(tuple(shape)).clone())
    nn.ModuleList, nn.ModuleDict = torch.nn.ModuleList, torch.nn.ModuleDict
    for name, mod in (('tensorlayerx', tlx), ('tensorlayerx.nn', nn), ('tensorlayerx.ops', ops), ('tensorlayerx.initializers', inits)):
        sys.modules[name] = mod
    return tlx


# ----------------------------------------------------------------------------------------------------------------------
# stand-in gammagl package tree holding the reference files that matter, loaded by path
# ----------------------------------------------------------------------------------------------------------------------
def _pkg(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    m.__package__ = name
    sys.modules[name] = m
    return m


def _load(name, relpath):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def load_reference_layers():
    from oracle.build_ref import load_ref
    ext = load_ref()
    assert ext is not None, 'reference extension could not be built (needs /root/reference)'
    make_tlx()
    os.environ['TL_BACKEND'] = 'torch'
    _pkg('gammagl', REF)
    mp = _pkg('gammagl.mpops', os.path.join(REF, 'mpops'))
    _pkg('gammagl.mpops.torch_ext', os.path.join(REF, 'mpops', 'torch_ext'))
    sys.modules['gammagl.mpops.torch_ext._torch_ext'] = ext                   # the reference's compiled CPU operators
    mpt = _load('gammagl.mpops.torch', 'mpops/torch.py')
    assert mpt.use_ext, 'the reference mpops did not bind its extension'
    for k, v in vars(mpt).items():                                             # `from .torch import *` (mpops/__init__.py:22)
        if not k.startswith('_'):
            setattr(mp, k, v)
    sys.modules.setdefault('pyparsing', types.ModuleType('pyparsing'))       # only used by Inspector.types(), not on this path
    ut = _pkg('gammagl.utils', os.path.join(REF, 'utils'))
    ut.Inspector = _load('gammagl.utils.inspector', 'utils/inspector.py').Inspector
    ut.degree = _load('gammagl.utils.degree', 'utils/degree.py').degree
    ut.segment_softmax = _load('gammagl.utils.softmax', 'utils/softmax.py').segment_softmax
    ut.calc_gcn_norm = _load('gammagl.utils.norm', 'utils/norm.py').calc_gcn_norm
    ut.add_self_loops = _load('gammagl.utils.loop', 'utils/loop.py').add_self_loops
    _pkg('gammagl.layers', os.path.join(REF, 'layers'))
    conv = _pkg('gammagl.layers.conv', os.path.join(REF, 'layers', 'conv'))
    conv.MessagePassing = _load('gammagl.layers.conv.message_passing', 'layers/conv/message_passing.py').MessagePassing
    out = {}
    for cls, f in (('GCNConv', 'gcn_conv'), ('SAGEConv', 'sage_conv'), ('GATConv', 'gat_conv'), ('GINConv', 'gin_conv'),
                   ('SGConv', 'sgc_conv'), ('APPNPConv', 'appnp_conv')):
        out[cls] = getattr(_load('gammagl.layers.conv.' + f, 'layers/conv/%s.py' % f), cls)
    return out


def load_extra_layer(cls, f):
    """Another reference conv layer (after load_reference_layers()), e.g. load_extra_layer('GCNIIConv', 'gcnii_conv')."""
    return getattr(_load('gammagl.layers.conv.' + f, 'layers/conv/%s.py' % f), cls)


def graph(N, E, seed):
    g = torch.Generator().manual_seed(seed)
    src = torch.randint(0, N, (E,), generator=g)
    dst = torch.randint(0, N - N // 10, (E,), generator=g)
    loops = torch.arange(N)
    return torch.stack([torch.cat([src, loops]), torch.cat([dst, loops])])   # self loops: no zero degrees


def run(layer, x, *args):
    layer.eval()
    x = x.clone().requires_grad_(True)
    out = layer(x, *args)
    gout = torch.randn(out.shape, generator=torch.Generator().manual_seed(99))
    out.backward(gout)
    return out.detach().numpy(), x.grad.numpy(), gout.numpy()


def params(layer):
    return {'p_' + k.replace('.', '__'): v.detach().numpy() for k, v in layer.named_parameters()}


def main():
    L = load_reference_layers()
    N, E, F, H = 300, 4000, 33, 20
    ei = graph(N, E, 1)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(N, F, generator=g)
    w = torch.rand(ei.shape[1], generator=g)
    rec = {'ei': ei.numpy(), 'x': x.numpy(), 'w': w.numpy()}

    def put(tag, layer, *args):
        for p in layer.parameters():                                           # non-trivial biases etc.
            if float(p.detach().abs().sum()) == 0.0:
                p.data = torch.randn(p.shape, generator=torch.Generator().manual_seed(p.numel())) * 0.1
        out, gx, gout = run(layer, x, *args)
        rec.update({tag + '_out': out, tag + '_gx': gx, tag + '_gout': gout})
        rec.update({tag + '_' + k: v for k, v in params(layer).items()})
        print(tag, out.shape, float(np.abs(out).max()))

    for norm in ('both', 'left', 'right', 'none'):
        put('gcn_' + norm, L['GCNConv'](F, H, norm=norm), ei)
    put('gcn_weighted', L['GCNConv'](F, H), ei, w)
    for aggr in ('mean', 'gcn', 'pool'):
        put('sage_' + aggr, L['SAGEConv'](F, H, aggr=aggr), ei)
    for heads, concat in ((8, True), (4, False), (1, True)):
        put('gat_h%d_%s' % (heads, 'cat' if concat else 'avg'), L['GATConv'](F, 8, heads=heads, concat=concat), ei)
    mlp = sys.modules['tensorlayerx'].nn.Linear(out_features=H, in_features=F)
    put('gin', L['GINConv'](mlp, eps=0.25), ei)
    put('sgc', L['SGConv'](F, H, iter_K=3), ei, w)
    put('appnp', L['APPNPConv'](F, H, iter_K=4, alpha=0.15, drop_rate=0.5), ei, w)
    np.savez_compressed(os.path.join(HERE, 'layers_ref.npz'), **rec)
    print('wrote layers_ref.npz with', len(rec), 'arrays')


if __name__ == '__main__':
    main()
