"""numpy front of the CPU oracle (oracle/mpops_oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; nothing under gammagl_b200/ does.  Parity status: PINNED (see the header of mpops_oracle.c and
tests/test_oracle.py: every golden vector of the reference's own tests + the compiled reference extension).

The shared object is compiled on first use with `gcc -O2 -ffp-contract=off` into oracle/_build/.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'mpops_oracle.c')
OUT_DIR = os.path.join(HERE, '_build')
SO = os.path.join(OUT_DIR, 'libmpops_oracle.so')

_lib = None


def build(force=False):
    if not force and os.path.exists(SO) and os.path.getmtime(SO) >= os.path.getmtime(SRC):
        return SO
    os.makedirs(OUT_DIR, exist_ok=True)
    subprocess.run(['gcc', '-O2', '-ffp-contract=off', '-shared', '-fPIC', '-o', SO, SRC], check=True)
    return SO


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _i64(v):
    return ctypes.c_int64(int(v))


_SUF = {np.dtype(np.float32): 'f32', np.dtype(np.float64): 'f64', np.dtype(np.float16): 'f16',
        np.dtype(np.int8): 'i8', np.dtype(np.int16): 'i16', np.dtype(np.int32): 'i32', np.dtype(np.int64): 'i64'}


def _prep(x, idx):
    x = np.ascontiguousarray(x)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    E = x.shape[0]
    K = int(np.prod(x.shape[1:])) if x.ndim > 1 else 1
    return x, idx, E, K


def segment(op, x, idx, N):
    """op in {'sum','mean','max'}; returns out (and arg for max), shapes [N, *x.shape[1:]]."""
    x, idx, E, K = _prep(x, idx)
    suf = _SUF[x.dtype]
    out = np.zeros((N,) + x.shape[1:], dtype=x.dtype)
    if op == 'max':
        arg = np.full(out.shape, N, dtype=np.int64)
        if x.size == 0:
            return out, arg                                   # cpu/segment_max_cpu.cpp:28-30: zeros, arg = N
        getattr(lib(), 'seg_max_' + suf)(_p(x), _p(idx), _i64(E), _i64(K), _i64(N), _p(out), _p(arg))
        return out, arg
    if x.size == 0:
        return out
    getattr(lib(), 'seg_%s_%s' % (op, suf))(_p(x), _p(idx), _i64(E), _i64(K), _i64(N), _p(out))
    return out


def segment_bwd(op, g, idx, E, arg=None):
    g = np.ascontiguousarray(g, dtype=np.float32)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    N = g.shape[0]
    K = int(np.prod(g.shape[1:])) if g.ndim > 1 else 1
    gx = np.zeros((E,) + g.shape[1:], dtype=np.float32)
    if op == 'sum':
        lib().seg_sum_bwd_f32(_p(g), _p(idx), _i64(E), _i64(K), _p(gx))
    elif op == 'mean':
        lib().seg_mean_bwd_f32(_p(g), _p(idx), _i64(E), _i64(K), _i64(N), _p(gx))
    else:
        arg = np.ascontiguousarray(arg, dtype=np.int64)
        lib().seg_max_bwd_f32(_p(g), _p(arg), _p(idx), _i64(E), _i64(K), _i64(N), _p(gx))
    return gx


def _prep_spmm(index, w, x):
    index = np.ascontiguousarray(index, dtype=np.int64)
    w = np.ascontiguousarray(w, dtype=np.float32)
    x = np.ascontiguousarray(x, dtype=np.float32)
    return index, w, x, index.shape[1], x.shape[0], int(np.prod(x.shape[1:]))


def spmm_fwd(op, index, w, x):
    """Returns (out, aux): aux = count (mean), argsrc (max), None (sum)."""
    index, w, x, E, N, K = _prep_spmm(index, w, x)
    out = np.empty_like(x)
    if op == 'sum':
        lib().spmm_sum_fwd(_p(index), _p(w), _p(x), _i64(E), _i64(N), _i64(K), _p(out))
        return out, None
    if op == 'mean':
        cnt = np.zeros(N, dtype=np.int64)
        lib().spmm_mean_fwd(_p(index), _p(w), _p(x), _i64(E), _i64(N), _i64(K), _p(out), _p(cnt))
        return out, cnt
    arg = np.zeros(x.shape, dtype=np.int64)
    lib().spmm_max_fwd(_p(index), _p(w), _p(x), _i64(E), _i64(N), _i64(K), _p(out), _p(arg))
    return out, arg


def spmm_bwd(op, index, w, g, aux=None):
    index, w, g, E, N, K = _prep_spmm(index, w, g)
    gx = np.empty_like(g)
    if op == 'sum':
        lib().spmm_sum_bwd(_p(index), _p(w), _p(g), _i64(E), _i64(N), _i64(K), _p(gx))
    elif op == 'mean':
        aux = np.ascontiguousarray(aux, dtype=np.int64)
        lib().spmm_mean_bwd(_p(index), _p(w), _p(g), _p(aux), _i64(E), _i64(N), _i64(K), _p(gx))
    else:
        aux = np.ascontiguousarray(aux, dtype=np.int64)
        lib().spmm_max_bwd(_p(index), _p(w), _p(g), _p(aux), _i64(E), _i64(N), _i64(K), _p(gx))
    return gx


def bspmm_fwd(index, w, x):
    index = np.ascontiguousarray(index, dtype=np.int64)
    w = np.ascontiguousarray(w, dtype=np.float32)
    x = np.ascontiguousarray(x, dtype=np.float32)
    N, H, C = x.shape
    out = np.empty_like(x)
    lib().bspmm_sum_fwd(_p(index), _p(w), _p(x), _i64(index.shape[1]), _i64(N), _i64(H), _i64(C), _p(out))
    return out


def bspmm_bwd(index, w, x, g):
    index = np.ascontiguousarray(index, dtype=np.int64)
    w = np.ascontiguousarray(w, dtype=np.float32)
    x = np.ascontiguousarray(x, dtype=np.float32)
    g = np.ascontiguousarray(g, dtype=np.float32)
    N, H, C = x.shape
    gx = np.empty_like(x)
    gw = np.empty_like(w)
    lib().bspmm_sum_bwd(_p(index), _p(w), _p(x), _p(g), _i64(index.shape[1]), _i64(N), _i64(H), _i64(C), _p(gx), _p(gw))
    return gx, gw


def segment_softmax(data, idx, N):
    """gammagl/utils/softmax.py:29-35 composed from the oracle's segment_max / segment_sum (float32)."""
    data = np.ascontiguousarray(data, dtype=np.float32)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    mx, _ = segment('max', data, idx, N)
    ex = np.exp(data - mx[idx]).astype(np.float32)
    den = segment('sum', ex, idx, N)
    return (ex / (den[idx] + np.float32(1e-16))).astype(np.float32)


def ind2ptr(ind, M):
    ind = np.ascontiguousarray(ind, dtype=np.int64)
    ptr = np.zeros(M + 1, dtype=np.int64)
    lib().ind2ptr_i64(_p(ind), _i64(ind.shape[0]), _i64(M), _p(ptr))
    return ptr


def ptr2ind(ptr, E):
    ptr = np.ascontiguousarray(ptr, dtype=np.int64)
    ind = np.zeros(E, dtype=np.int64)
    lib().ptr2ind_i64(_p(ptr), _i64(ptr.shape[0] - 1), _i64(E), _p(ind))
    return ind


def sort_edge_index(edge_index, num_nodes, sort_by_row=True):
    """gammagl/utils/sort_edge_index.py:33-38: order by edge_index[1-sbr] * N + edge_index[sbr].
    Stable argsort (the reference's torch.argsort leaves the order of equal keys open).  Returns (sorted, perm)."""
    ei = np.ascontiguousarray(edge_index, dtype=np.int64)
    s = int(bool(sort_by_row))
    idx = ei[1 - s] * np.int64(num_nodes) + ei[s]
    perm = np.argsort(idx, kind='stable')
    return ei[:, perm], perm


def coalesce(edge_index, edge_attr, num_nodes, sort_by_row=True):
    """gammagl/utils/coalesce.py:41-78: sort, keep first occurrences (mask = idx[1:] > idx[:-1]), merge duplicate
    attributes with segment_sum over idx = arange - cumsum(1 - mask) (:73-76).  Returns (edge_index, attr | None)."""
    ei, perm = sort_edge_index(edge_index, num_nodes, sort_by_row)
    s = int(bool(sort_by_row))
    key = np.concatenate([[-1], ei[1 - s] * np.int64(num_nodes) + ei[s]])
    mask = key[1:] > key[:-1]
    attr = None if edge_attr is None else np.ascontiguousarray(edge_attr)[perm]
    if mask.all():
        return ei, attr
    out = ei[:, mask]
    if attr is None:
        return out, None
    idx = np.arange(ei.shape[1]) - (1 - mask).cumsum()
    return out, segment('sum', attr, idx, out.shape[1])


def to_undirected(edge_index, edge_attr, num_nodes):
    """gammagl/utils/undirected.py:99-109: both directions concatenated, then coalesce."""
    ei = np.ascontiguousarray(edge_index, dtype=np.int64)
    both = np.stack([np.concatenate([ei[0], ei[1]]), np.concatenate([ei[1], ei[0]])])
    attr = None if edge_attr is None else np.concatenate([edge_attr, edge_attr], axis=0)
    return coalesce(both, attr, num_nodes, True)


def is_undirected(edge_index, edge_attr, num_nodes):
    """gammagl/utils/undirected.py:35-59."""
    ei1, p1 = sort_edge_index(edge_index, num_nodes, True)
    ei2, p2 = sort_edge_index(ei1, num_nodes, False)
    ok = bool(np.all(ei1[0] == ei2[1]) and np.all(ei1[1] == ei2[0]))
    if edge_attr is not None:
        a1 = np.asarray(edge_attr)[p1]
        ok = ok and bool(np.all(a1 == a1[p2]))
    return ok
