"""enable(): plug the B200 backend into an installed GammaGL (PyTorch backend) without editing it.

Two plug points, both left by the reference itself:
  1. `gammagl/mpops/torch.py:4` binds seven symbols from `_torch_ext` and `gammagl/mpops/__init__.py:10-30`
     star-exports nine functions; every caller (`layers/conv/message_passing.py:2`, `utils/softmax.py`,
     `utils/degree.py`, `utils/norm.py`, 39 files) imported them by name.  enable() rebinds those names, in
     `gammagl.mpops`, `gammagl.mpops.torch` and in every already-imported `gammagl.*` module that holds a copy.
  2. `MessagePassing.propagate` takes the fused branch iff the concrete class defines `message_aggregate`
     (`message_passing.py:144`).  enable(fuse_layers=True) installs one on GCNConv / SAGEConv / GATConv / GINConv and the multi-hop layers --
     the override the authors left commented out (`gcn_conv.py:110-115`, `gat_conv.py:124-129`).
  3. (utils=True) `gammagl.utils.sort_edge_index` / `coalesce` copies are wrapped: CUDA tensors go to the GPU two-key sort
     (`gammagl_b200.utils`), numpy arrays and CPU tensors to the original.
disable() restores the originals.

Layers constructed BEFORE enable(): MessagePassing.__init__ files the signature of `message_aggregate` in the instance's
Inspector (message_passing.py:29-33), and the base signature has no `num_nodes`, so such an instance would hand
num_nodes=None to the fused method (wrong row count for bipartite (src, dst) calls).  enable() therefore re-inspects every
live instance of the patched classes (and disable() restores them); models may be built before or after enable().
"""
import functools
import sys
import types

import torch

from . import mpops
from . import utils as _utils

_NAMES = ['unsorted_segment_sum', 'unsorted_segment_mean', 'unsorted_segment_max', 'segment_sum', 'segment_mean',
          'segment_max', 'gspmm', 'bspmm', 'c_segment_sum', 'c_segment_mean', 'c_segment_max', 'c_spmm_sum',
          'c_spmm_mean', 'c_spmm_max', 'c_bspmm_sum']
# Layers whose `message` is the base gather-times-weight (message_passing.py:55-61) and that call propagate() with 1-D (or
# per-head) edge weights: the fused operator computes exactly their message + aggregate.  The second line are the multi-hop
# layers (K propagates on one edge_index per forward: sgc_conv.py:47-48, appnp_conv.py:30-36, gcnii_conv.py:75-86,
# gpr_conv.py:72-73, mixhop_conv.py:95-100, cheb_conv.py:117-121, hid_conv.py:95-96, magcl_conv.py:85-87), where the CSR
# cached on the first hop serves all the others.
_FUSED_LAYERS = ['GCNConv', 'SAGEConv', 'GATConv', 'GINConv',
                 'SGConv', 'APPNPConv', 'GCNIIConv', 'GPRConv', 'MixHopConv', 'ChebConv', 'Hid_conv', 'MAGCLConv']
_UTILS = ['sort_edge_index', 'coalesce']   # gammagl/utils/sort_edge_index.py, coalesce.py: host argsort -> GPU two-key sort
_saved = []


def _cuda_or_original(new, old):
    """The reference utilities also serve numpy arrays and CPU tensors: only CUDA tensors are routed to the GPU sort."""
    @functools.wraps(old)
    def dispatch(edge_index, *args, **kwargs):
        if isinstance(edge_index, torch.Tensor) and edge_index.is_cuda:
            return new(edge_index, *args, **kwargs)
        return old(edge_index, *args, **kwargs)
    dispatch._gglb200_original = old
    return dispatch


def _fused_message_aggregate(self, x, edge_index, edge_weight=None, aggr='sum', num_nodes=None):
    """Fused gather x weight -> reduce for the call shapes of the reference layers:
      [N,K] x, 1-D / no weights, sum | mean    -> gspmm (bipartite calls: num_dst = num_nodes output rows)
      [N,K] x, 1-D / no weights, max           -> gspmm_max_edge (winning EDGE recorded: the gradient of message +
                                                   unsorted_segment_max, src/segment_max.cpp:48-61, on multigraphs too)
      [N,H,C] x, [E,H] weights, sum            -> bspmm (gradient for the weights through the SDDMM kernel)
      differentiable 1-D weights, sum          -> one-head bspmm (gspmm has no weight gradient, src/gspmm.cpp:30,79)
    Anything else runs the reference's own decomposition on the B200 segment operators, size-guarded: the [E,K] message
    is refused above GGLB200_MAX_MESSAGE_BYTES instead of being allocated silently."""
    w_grad = mpops.weight_needs_grad(edge_weight)
    square = num_nodes is None or num_nodes == x.shape[0]
    per_head = edge_weight is not None and edge_weight.dim() == 2 and x.dim() == 3
    if per_head and square and aggr == 'sum':
        return mpops.bspmm(edge_index, edge_weight, x, aggr)
    if w_grad and aggr == 'sum' and x.dim() == 2 and edge_weight.dim() == 1 and square:
        return mpops.bspmm(edge_index, edge_weight.reshape(-1, 1), x.unsqueeze(1), 'sum').squeeze(1)
    if not w_grad and not per_head and (edge_weight is None or edge_weight.dim() == 1):
        if aggr == 'max':
            return mpops.gspmm_max_edge(edge_index, edge_weight, x, num_dst=num_nodes)
        if aggr in ('sum', 'mean'):
            return mpops.gspmm(edge_index, edge_weight, x, aggr, num_dst=num_nodes)
    mpops.check_message_bytes(edge_index.shape[1], x[0].numel() if x.dim() > 1 else 1, x.element_size(),
                              '%s.message_aggregate(aggr=%r)' % (type(self).__name__, aggr))
    msg = self.message(x, edge_index, edge_weight) if edge_weight is not None else self.message(x, edge_index)
    return self.aggregate(msg, edge_index, num_nodes=num_nodes, aggr=aggr)


def _gatv2_message_aggregate(self, x, edge_index, edge_weight=None, num_nodes=None, aggr='sum'):
    """GATv2Conv (layers/conv/gatv2_conv.py:95-108): its logit mean_c(leaky(x_j) att_src + leaky(x_i) att_dst) is a source-node
    scalar plus a destination-node scalar, so attention + aggregation is the fused GAT kernel with slope 1 on those scalars."""
    H, C = int(x.shape[1]), int(x.shape[2])
    training = bool(getattr(self, 'is_train', getattr(self, 'training', False)))
    if (edge_weight is None and aggr == 'sum' and x.is_cuda and x.dtype == torch.float32 and mpops.gat_fusable_grouped(H, C)
            and (num_nodes is None or num_nodes == x.shape[0]) and not (training and float(getattr(self, 'dropout_rate', 0.)) > 0)):
        lx = torch.nn.functional.leaky_relu(x, float(self.negetive_slop))
        return mpops.gat_aggregate(edge_index, x, (lx * self.att_src).mean(-1), (lx * self.att_dst).mean(-1), 1.0)
    mpops.check_message_bytes(edge_index.shape[1], x[0].numel(), x.element_size(), 'GATv2Conv.message')
    return self.aggregate(self.message(x, edge_index, edge_weight, num_nodes), edge_index, num_nodes=num_nodes, aggr=aggr)


def _agnn_message_aggregate(self, x, edge_index, num_nodes=None, edge_weight=None, aggr='sum'):
    """AGNNConv (layers/conv/agnn_conv.py:57-70): cosine logits by SDDMM over the normalised rows, edge softmax, one-head bspmm."""
    if edge_weight is None and aggr == 'sum' and x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and (num_nodes is None or num_nodes == x.shape[0]):
        xn = torch.nn.functional.normalize(x, p=2.0, dim=-1, eps=1e-12)
        cos = mpops.sddmm(edge_index, xn.unsqueeze(1), xn.unsqueeze(1)).reshape(-1)
        alpha = mpops.segment_softmax(cos * self.beta, edge_index[1], x.shape[0])
        return mpops.bspmm(edge_index, alpha.reshape(-1, 1), x.unsqueeze(1), 'sum').squeeze(1)
    mpops.check_message_bytes(edge_index.shape[1], x[0].numel(), x.element_size(), 'AGNNConv.message')
    return self.aggregate(self.message(x, edge_index, num_nodes, edge_weight), edge_index, num_nodes=num_nodes, aggr=aggr)


for _f in (_gatv2_message_aggregate, _agnn_message_aggregate):
    _f.__name__ = _f.__qualname__ = 'message_aggregate'      # Inspector files signatures by __name__ (see below)
# layers with their OWN message(): the attention they compute maps onto a fused operator as a whole
_SPECIAL_LAYERS = {'GATV2Conv': _gatv2_message_aggregate, 'AGNNConv': _agnn_message_aggregate}


# MessagePassing.__init__ registers the method with Inspector.inspect(), which files the signature under func.__name__
# (utils/inspector.py:27-41) and then looks it up as 'message_aggregate' (message_passing.py:29-33): the installed function
# must carry that name or constructing any patched layer raises KeyError.
_fused_message_aggregate.__name__ = 'message_aggregate'
_fused_message_aggregate.__qualname__ = 'message_aggregate'


def enable(fuse_layers=True, modules=None, utils=True):
    """Rebind the aggregation operators of an imported `gammagl` to this backend.  Returns the number of rebinds."""
    mods = modules if modules is not None else {k: v for k, v in sys.modules.items()
                                                if v is not None and (k == 'gammagl' or k.startswith('gammagl.'))}
    if not mods:
        raise ImportError('gammagl is not imported; import it (TL_BACKEND=torch) before gammagl_b200.plugin.enable()')
    count = 0
    for name, mod in list(mods.items()):
        for fn in _NAMES + ['use_ext']:
            if hasattr(mod, fn):
                new = True if fn == 'use_ext' else getattr(mpops, fn)
                old = getattr(mod, fn)
                if old is not new:
                    _saved.append((mod, fn, old, True))
                    setattr(mod, fn, new)
                    count += 1
        # every copy: utils/softmax.py defines it, utils/__init__ and gat_conv.py (`from gammagl.utils import segment_softmax`) hold one
        if callable(getattr(mod, 'segment_softmax', None)) and mod.segment_softmax is not mpops.segment_softmax:
            _saved.append((mod, 'segment_softmax', getattr(mod, 'segment_softmax'), True))
            setattr(mod, 'segment_softmax', mpops.segment_softmax)
            count += 1
        if utils:
            for fn in _UTILS:
                old = getattr(mod, fn, None)
                if callable(old) and not isinstance(old, types.ModuleType) and not hasattr(old, '_gglb200_original'):
                    _saved.append((mod, fn, old, True))
                    setattr(mod, fn, _cuda_or_original(getattr(_utils, fn), old))
                    count += 1
        if fuse_layers:
            for cls_name in _FUSED_LAYERS + list(_SPECIAL_LAYERS):
                cls = getattr(mod, cls_name, None)
                if isinstance(cls, type) and 'message_aggregate' not in cls.__dict__ and hasattr(cls, 'propagate'):
                    _saved.append((cls, 'message_aggregate', None, False))
                    cls.message_aggregate = _SPECIAL_LAYERS.get(cls_name, _fused_message_aggregate)
                    _patched_classes.append(cls)
                    count += 1
    if fuse_layers:
        _reinspect_live_instances()
    return count


_patched_classes = []


def _reinspect_live_instances():
    """Refresh the Inspector entry of `message_aggregate` on every live instance of the patched classes (see module doc)."""
    import gc
    import warnings
    classes = tuple(_patched_classes)
    if not classes:
        return
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')      # isinstance() on lazy module proxies may emit deprecation warnings
        for obj in gc.get_objects():
            try:
                if isinstance(obj, classes) and hasattr(obj, 'inspector') and hasattr(obj.inspector, 'inspect'):
                    obj.inspector.inspect(obj.message_aggregate)
            except Exception:   # objects that misbehave under isinstance / getattr are not ours
                continue


def disable():
    while _saved:
        obj, name, old, had = _saved.pop()
        if had:
            setattr(obj, name, old)
        elif name in obj.__dict__:
            delattr(obj, name)
    _reinspect_live_instances()          # back to the base signature on the instances
    del _patched_classes[:]
