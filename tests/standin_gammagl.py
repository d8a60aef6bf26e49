"""A stand-in `gammagl` module tree for the plug-in tests (the reference checkout does not travel to the GPU box).

Written from the CONTRACT of the reference, not from its text:
  * gammagl.mpops / gammagl.mpops.torch export the nine operator names + seven c_* symbols (mpops/__init__.py:10-30,
    mpops/torch.py:4); `from gammagl.mpops import *` leaves COPIES of them in gammagl.layers.conv.message_passing and in
    gammagl.utils.softmax (message_passing.py:2, utils/softmax.py) -- the copies are what the layers call;
  * Inspector.inspect files inspect.signature(func).parameters under func.__name__, Inspector.distribute picks arguments by
    name from a dict and raises TypeError for a required one that is missing (utils/inspector.py:27-41,90-107);
  * MessagePassing.__init__ inspects `message` and `message_aggregate`; propagate() fills num_nodes with x.shape[0] when
    absent and takes the fused branch iff the CONCRETE class defines message_aggregate (message_passing.py:122-156);
    message = x[src] (* w[..., None]); aggregate = unsorted_segment_{sum,mean,max}(msg, dst, num_nodes).
The stand-in operators raise: after plugin.enable() none of them may run."""
import inspect
import sys
import types
from collections import OrderedDict

import torch

NAMES = ['unsorted_segment_sum', 'unsorted_segment_mean', 'unsorted_segment_max', 'segment_sum', 'segment_mean', 'segment_max',
         'gspmm', 'bspmm', 'c_segment_sum', 'c_segment_mean', 'c_segment_max', 'c_spmm_sum', 'c_spmm_mean', 'c_spmm_max',
         'c_bspmm_sum']


def _must_not_run(*a, **k):
    raise AssertionError('an original (stand-in) operator ran after plugin.enable()')


class Inspector:
    def __init__(self, base):
        self.base_class = base
        self.params = {}

    def inspect(self, func, pop_first=False):
        params = OrderedDict(inspect.signature(func).parameters)
        if pop_first:
            params.popitem(last=False)
        self.params[func.__name__] = params

    def keys(self, func_names=None):
        ks = []
        for f in func_names or list(self.params):
            ks += list(self.params[f])
        return set(ks)

    def distribute(self, func_name, kwargs):
        out = {}
        for key, param in self.params[func_name].items():
            data = kwargs.get(key, inspect.Parameter.empty)
            if data is inspect.Parameter.empty:
                if param.default is inspect.Parameter.empty:
                    raise TypeError('Required parameter %s is empty.' % key)
                data = param.default
            out[key] = data
        return out


def make():
    """Returns {module name: module} (not inserted into sys.modules) laid out like the reference package."""
    pkg = types.ModuleType('gammagl')
    mp = types.ModuleType('gammagl.mpops')
    mpt = types.ModuleType('gammagl.mpops.torch')
    for m in (mp, mpt):
        for n in NAMES:
            setattr(m, n, _must_not_run)
        m.use_ext = False
    msgmod = types.ModuleType('gammagl.layers.conv.message_passing')
    for n in NAMES[:8]:
        setattr(msgmod, n, _must_not_run)          # star-import copies
    G = msgmod.__dict__

    class MessagePassing(torch.nn.Module):
        special_args = {'edge_index', 'x', 'edge_weight'}

        def __init__(self):
            super().__init__()
            self.inspector = Inspector(self)
            self.inspector.inspect(self.message)
            self.inspector.inspect(self.message_aggregate)

        def message(self, x, edge_index, edge_weight=None):
            msg = x[edge_index[0]]
            return msg * edge_weight.unsqueeze(-1) if edge_weight is not None else msg

        def aggregate(self, msg, edge_index, num_nodes=None, aggr='sum'):
            dst = edge_index[1]
            return G['unsorted_segment_' + aggr](msg, dst, num_nodes)      # module-level copy, looked up at call time

        def message_aggregate(self, x, edge_index, edge_weight=None, aggr='sum'):
            return G['gspmm'](edge_index, edge_weight, x, aggr)

        def update(self, x):
            return x

        def propagate(self, x, edge_index, aggr='sum', **kwargs):
            if kwargs.get('num_nodes') is None:
                kwargs['num_nodes'] = x.shape[0]
            coll = dict(kwargs, x=x, edge_index=edge_index, aggr=aggr)
            if 'message_aggregate' in self.__class__.__dict__:
                x = self.message_aggregate(**self.inspector.distribute('message_aggregate', coll))
            else:
                msg = self.message(**self.inspector.distribute('message', coll))
                x = self.aggregate(msg, edge_index, num_nodes=kwargs['num_nodes'], aggr=aggr)
            return self.update(x)
    msgmod.MessagePassing = MessagePassing

    sm = types.ModuleType('gammagl.utils.softmax')
    sm.segment_softmax = _must_not_run
    sm.unsorted_segment_max = _must_not_run
    sm.unsorted_segment_sum = _must_not_run
    S = sm.__dict__

    conv = types.ModuleType('gammagl.layers.conv')

    class GCNConv(MessagePassing):                  # call shape of gcn_conv.py:104: 2-D x, 1-D weights, num_nodes
        def forward(self, x, edge_index, edge_weight, num_nodes):
            return self.propagate(x, edge_index, edge_weight=edge_weight, num_nodes=num_nodes)

    class SAGEConv(MessagePassing):                 # sage_conv.py:75-92: (src_feat, dst_feat), num_nodes = #dst, mean / max
        def forward(self, feat, edge_index, aggr):
            src_feat, dst_feat = feat if isinstance(feat, tuple) else (feat, feat)
            return self.propagate(src_feat, edge_index, edge_weight=None, num_nodes=int(dst_feat.shape[0]), aggr=aggr)

    class GATConv(MessagePassing):                  # gat_conv.py:103-112: [N,H,C] x, [E,H] weights from segment_softmax
        def forward(self, x, e, edge_index, num_nodes):
            alpha = S['segment_softmax'](e, edge_index[1], num_nodes)
            return self.propagate(x, edge_index, num_nodes=num_nodes, edge_weight=alpha)

    class GINConv(MessagePassing):                  # gin_conv.py:55: no weights, sum
        def forward(self, x, edge_index):
            return self.propagate(x, edge_index, edge_weight=None, aggr='sum')

    class GATV2Conv(MessagePassing):                # gatv2_conv.py:95-123: own message() = attention + weighting
        def __init__(self, heads, channels, negative_slope=0.2):
            super().__init__()
            self.negetive_slop, self.dropout_rate = negative_slope, 0.
            g = torch.Generator().manual_seed(3)
            self.att_src = torch.nn.Parameter(torch.randn(1, heads, channels, generator=g))
            self.att_dst = torch.nn.Parameter(torch.randn(1, heads, channels, generator=g))

        def message(self, x, edge_index, edge_weight=None, num_nodes=None):
            src, dst = edge_index[0], edge_index[1]
            lr = torch.nn.functional.leaky_relu
            weight = (lr(x[src], self.negetive_slop) * self.att_src + lr(x[dst], self.negetive_slop) * self.att_dst).mean(-1)
            alpha = S['segment_softmax'](weight, dst, num_nodes)
            return x[src] * alpha.unsqueeze(-1)

        def forward(self, x, edge_index, num_nodes=None):
            return self.propagate(x, edge_index, num_nodes=num_nodes)

    class AGNNConv(MessagePassing):                 # agnn_conv.py:57-73
        def __init__(self):
            super().__init__()
            self.beta = torch.nn.Parameter(torch.tensor([0.7]))

        def message(self, x, edge_index, num_nodes, edge_weight=None):
            src, dst = edge_index[0], edge_index[1]
            nrm = lambda t: torch.nn.functional.normalize(t, p=2.0, dim=-1, eps=1e-12)
            cos = (nrm(x[src]) * nrm(x[dst])).sum(-1)
            return S['segment_softmax'](cos * self.beta, dst, num_nodes).unsqueeze(-1) * x[src]

        def forward(self, x, edge_index, num_nodes):
            return self.propagate(x, edge_index, num_nodes=num_nodes)

    class HGTConv(MessagePassing):                  # a layer with its own message(): must stay unfused
        def message(self, x, edge_index, edge_weight=None):
            return x[edge_index[0]] * 2.0
    for c in (GCNConv, SAGEConv, GATConv, GINConv, GATV2Conv, AGNNConv, HGTConv):
        setattr(conv, c.__name__, c)
    return {'gammagl': pkg, 'gammagl.mpops': mp, 'gammagl.mpops.torch': mpt, 'gammagl.layers.conv.message_passing': msgmod,
            'gammagl.layers.conv': conv, 'gammagl.utils.softmax': sm}
