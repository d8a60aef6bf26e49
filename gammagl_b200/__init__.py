"""gammagl_b200 -- Blackwell (sm_100a) message-passing aggregation backend for GammaGL's PyTorch path.

Scope: the one hot path `gammagl.mpops` (unsorted_segment_sum/mean/max, gspmm, bspmm) behind
MessagePassing.propagate(); see DESIGN.md.  Layout:

    csrc/       hand-written CUDA kernels + the C ABI (libgglb200.so, declared in include/gglb200.h)
    _lib.py     ctypes binding (fails loudly; no CPU fallback)
    csr.py      COO -> CSR ingest + caches
    mpops.py    the 9 `gammagl.mpops` functions and the 7 `c_*` extension symbols, as autograd Functions
    layers.py   host-side mirror of MessagePassing / GCNConv / SAGEConv / GATConv / GINConv forward paths
    plugin.py   enable(): rebinding of gammagl.mpops.* inside an installed GammaGL
"""
from . import _lib  # noqa: F401
from .mpops import (bspmm, c_bspmm_sum, c_segment_max, c_segment_mean, c_segment_sum, c_spmm_max, c_spmm_mean,  # noqa: F401
                    c_spmm_sum, gspmm, segment_max, segment_mean, segment_softmax, segment_sum,
                    unsorted_segment_max, unsorted_segment_mean, unsorted_segment_sum)

__version__ = '0.1.0'
