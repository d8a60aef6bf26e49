"""enable(): plug the B200 backend into an installed GammaGL (PyTorch backend) without editing it.

Two plug points, both left by the reference itself:
  1. `gammagl/mpops/torch.py:4` binds seven symbols from `_torch_ext` and `gammagl/mpops/__init__.py:10-30`
     star-exports nine functions; every caller (`layers/conv/message_passing.py:2`, `utils/softmax.py`,
     `utils/degree.py`, `utils/norm.py`, 39 files) imported them by name.  enable() rebinds those names, in
     `gammagl.mpops`, `gammagl.mpops.torch` and in every already-imported `gammagl.*` module that holds a copy.
  2. `MessagePassing.propagate` takes the fused branch iff the concrete class defines `message_aggregate`
     (`message_passing.py:144`).  enable(fuse_layers=True) installs one on GCNConv / SAGEConv / GATConv / GINConv --
     the override the authors left commented out (`gcn_conv.py:110-115`, `gat_conv.py:124-129`).
disable() restores the originals.
"""
import sys

from . import mpops

_NAMES = ['unsorted_segment_sum', 'unsorted_segment_mean', 'unsorted_segment_max', 'segment_sum', 'segment_mean',
          'segment_max', 'gspmm', 'bspmm', 'c_segment_sum', 'c_segment_mean', 'c_segment_max', 'c_spmm_sum',
          'c_spmm_mean', 'c_spmm_max', 'c_bspmm_sum']
_FUSED_LAYERS = ['GCNConv', 'SAGEConv', 'GATConv', 'GINConv']
_saved = []


def _fused_message_aggregate(self, x, edge_index, edge_weight=None, aggr='sum', num_nodes=None):
    """Fused gather x weight -> reduce.  Falls back to the reference's own decomposition (still on the B200 segment
    operators) for bipartite calls, where gspmm's output row count (= x rows) would be wrong."""
    if num_nodes is not None and num_nodes != x.shape[0]:
        msg = self.message(x, edge_index, edge_weight) if edge_weight is not None else self.message(x, edge_index)
        return self.aggregate(msg, edge_index, num_nodes=num_nodes, aggr=aggr)
    if edge_weight is not None and edge_weight.dim() == 2 and x.dim() == 3:
        return mpops.bspmm(edge_index, edge_weight, x, aggr)
    return mpops.gspmm(edge_index, edge_weight, x, aggr)


def enable(fuse_layers=True, modules=None):
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
        if hasattr(mod, 'segment_softmax') and name.endswith('softmax'):
            _saved.append((mod, 'segment_softmax', getattr(mod, 'segment_softmax'), True))
            setattr(mod, 'segment_softmax', mpops.segment_softmax)
            count += 1
        if fuse_layers:
            for cls_name in _FUSED_LAYERS:
                cls = getattr(mod, cls_name, None)
                if isinstance(cls, type) and 'message_aggregate' not in cls.__dict__ and hasattr(cls, 'propagate'):
                    _saved.append((cls, 'message_aggregate', None, False))
                    cls.message_aggregate = _fused_message_aggregate
                    count += 1
    return count


def disable():
    while _saved:
        obj, name, old, had = _saved.pop()
        if had:
            setattr(obj, name, old)
        elif name in obj.__dict__:
            delattr(obj, name)
