"""gammagl_b200.mpops -- host-side mirror of `gammagl.mpops` (PyTorch backend) on the B200 kernels.

Same names, argument meaning and return values as /root/reference/gammagl/mpops/torch.py:
    unsorted_segment_sum / mean / max     (torch.py:43-202)
    segment_sum / mean / max              (torch.py:209-299; aliases, sortedness is not exploited there either)
    gspmm(index, weight, x, reduce)       (torch.py:302-351)
    bspmm(index, weight, x, reduce)       (torch.py:354-365)
and the seven extension symbols the reference binds from `_torch_ext` (torch.py:4,
torch_ext/src/operators.cpp:51-59):
    c_segment_sum / c_segment_mean / c_segment_max / c_spmm_sum / c_spmm_mean / c_spmm_max / c_bspmm_sum

Differences, all deliberate (DESIGN.md "Boundary"):
  * CUDA tensors only; any failure raises (the reference catches everything and falls back, torch.py:83-86).
  * empty segments of *_max hold lowest() like the compiled reference extension
    (cpu/segment_max_cpu.cpp:40), not 0 like the pure-torch fallback (torch.py:40).
  * the gradient of an empty max-segment is dropped instead of leaking into edge row N
    (src/segment_max.cpp:55-59 scatters the sentinel N).
  * gspmm / bspmm additionally accept bf16 / f16 x (fp32 accumulate); the reference is fp32-only.
"""
import ctypes
import os

import torch

from . import _lib
from .csr import _ptr, _stream, graph_of, note_sighting, require_cuda, segment_csr_if_cached, segment_csr_of, validate_ids

use_ext = True  # mirrors the module flag of the reference (torch.py:2-7); always the native path here
PROFILE = None  # bench.py sets this to a list: spmm_csr then appends (n_rows, start_event, end_event) per launch

_DTYPES = {
    torch.float32: _lib.GGL_F32, torch.bfloat16: _lib.GGL_BF16, torch.float16: _lib.GGL_F16,
    torch.float64: _lib.GGL_F64, torch.int8: _lib.GGL_I8, torch.int16: _lib.GGL_I16,
    torch.int32: _lib.GGL_I32, torch.int64: _lib.GGL_I64, torch.uint8: _lib.GGL_U8,
}
_REDUCE = {'sum': _lib.GGL_SUM, 'mean': _lib.GGL_MEAN, 'max': _lib.GGL_MAX}


def _dtype_code(t):
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise _lib.GglError('unsupported dtype %s' % t.dtype)


_PAD_ODD_K = os.environ.get('GGLB200_PAD_ODD_K', '1') != '0'


def spmm_csr(csr, col, x2d, reduce, w=None, wperm=None, heads=1, want_arg=False, arg_empty=0, n_rows=None, out_scale=None,
             bias=None, relu=False):
    """Launch the aggregation kernel family on a Csr.  x2d: [n_src, K] contiguous.  Returns (out, arg)."""
    lib = _lib.load()
    n_rows = csr.n_rows if n_rows is None else n_rows
    K = x2d.shape[1]
    dev = x2d.device
    if _PAD_ODD_K and heads == 1 and K > 4 and K % 4 != 0 and x2d.dtype == torch.float32 and 4 * x2d.shape[0] <= csr.E:
        # fp32 rows that are not a multiple of 16 bytes (47 classes: 188 B) run one element per lane: 14.4 ms per pass at C3 against
        # 4 ms for K = 48.  Gathered feature matrices are small next to the edge stream, so the rows are padded to the next
        # multiple of 4 columns (one copy of x), aggregated with 16-byte lanes and sliced back: columns are independent, results
        # are bit-identical.  (Segment operators reduce [E, K] messages: padding those would cost more than it saves.)
        pad = (-K) % 4
        xp = torch.nn.functional.pad(x2d, (0, pad))
        bp = None if bias is None else torch.nn.functional.pad(bias.reshape(-1), (0, pad))
        outp, argp = spmm_csr(csr, col, xp, reduce, w=w, wperm=wperm, heads=1, want_arg=want_arg, arg_empty=arg_empty, n_rows=n_rows,
                              out_scale=out_scale, bias=bp, relu=relu)
        return outp[:, :K].contiguous(), (None if argp is None else argp[:, :K].contiguous())
    with torch.cuda.device(dev):
        _lib.tune_device(dev.index)
        csr.use_on_current_stream()        # a CSR built on another stream: order this stream after its build
        out = torch.empty((n_rows, K), dtype=x2d.dtype, device=dev)
        arg = torch.empty((n_rows, K), dtype=torch.int64, device=dev) if want_arg else None
        ws, ws_bytes = csr.workspace(K, reduce)
        if PROFILE is not None:
            ev0 = torch.cuda.Event(enable_timing=True)
            ev0.record()
        if bias is not None or relu:       # layer epilogue fused into the row write (GGL_SUM only)
            _lib.check(lib.gglb200_spmm_csr_epilogue(_ptr(csr.rowptr), _ptr(col), _ptr(csr.plan), _ptr(w), _ptr(wperm), heads,
                                                     _ptr(x2d), _ptr(out), _ptr(out_scale), _ptr(bias), int(bool(relu)), n_rows,
                                                     x2d.shape[0], csr.E, K, _dtype_code(x2d), _ptr(ws), ws_bytes, _stream(dev)),
                       'spmm_csr_epilogue')
        else:
            _lib.check(lib.gglb200_spmm_csr(reduce, _ptr(csr.rowptr), _ptr(col), _ptr(csr.plan),
                                            _ptr(w), _ptr(wperm), heads, _ptr(x2d), _ptr(out), _ptr(arg), arg_empty, _ptr(out_scale),
                                            n_rows, x2d.shape[0], csr.E, K, _dtype_code(x2d), _ptr(ws), ws_bytes, _stream(dev)),
                       'spmm_csr')
        if PROFILE is not None:
            ev1 = torch.cuda.Event(enable_timing=True)
            ev1.record()
            PROFILE.append((n_rows, ev0, ev1))
    return out, arg


# ------------------------------------------------------------------------------------------------
# segment ops
# ------------------------------------------------------------------------------------------------
def _check_segment_args(x, index, N):
    require_cuda(x, 'x')
    require_cuda(index, 'segment_ids')
    if index.dim() != 1:
        raise IndexError('index dimension should be 1, but got %d' % index.dim())      # segment_sum_cpu.cpp:15-16
    if x.shape[0] != index.shape[0]:
        raise IndexError('fisrt dimension of x and index should be same')              # segment_sum_cpu.cpp:17-19
    if int(N) < 0:
        raise ValueError('num_segments must be >= 0')


# SUM / MEAN backward: row gather gx[e] = g[ids[e]] while the gradient rows stay L2-resident, scatter form through the
# forward CSR (every gradient row read once) beyond that.  Measured on B200 (K=32 fp32): N*K*4 = 30 MB: gather 3.3 ms vs
# scatter 3.7 ms; 313 MB: gather 4.8 ms vs scatter 4.0 ms.  GGLB200_SEG_BWD=gather|expand forces one form.
_SEG_BWD = os.environ.get('GGLB200_SEG_BWD', 'auto')
_SEG_BWD_L2_BYTES = 64 << 20


# Which path reduces an unsorted segment op (fp32; other dtypes always take the CSR path):
#   csr      sort the ids once (cached per ids tensor), then gather x[perm[i]] per pass: fp32 sums bit-identical to the
#            reference's serial CPU loop
#   oneshot  sort-free streaming pass with L2 reductions (csrc/oneshot.cu): no build, x read coalesced; max (+ arg) is
#            bit-exact there too, fp32 sums / means depend on arrival order (rounding only, <= 1e-6 relative; exact for
#            integer-valued data such as degree())
# GGLB200_SEGMENT = auto (default) | fast | csr | oneshot.  Measured on B200 (profiles/microbench_r2_*.jsonl, probe_atomics_r2.jsonl):
# vector reductions run at 0.9 G x 10^12 fp32 elements/s into an L2-resident output (red.global.add.v4.f32), 64-bit max atomics at
# 0.13 x 10^12/s -- so
#   auto  rows of at most GGLB200_ONESHOT_ROW_BYTES (16: K <= 4 fp32 -- degree() and calc_gcn_norm() reduce K = 1 rows, whose
#         permuted 4-byte gather wastes 7/8 of every DRAM sector: 1.75 / 0.85 ms one-shot against 4.0 / 4.2 ms through the CSR
#         at 1.2e8 edges) go one-shot for every reducer; wider rows stay on the bit-exact CSR path.
#   fast  additionally, sums / means of wider rows go one-shot the FIRST time an ids tensor is seen (4.6 ms at K = 32 against a
#         6 ms sort + 3.9 ms pass) and build the CSR from the second sighting on.  Wide max never goes one-shot (one 64-bit
#         atomic per element: 29 ms at K = 32).
# With torch.use_deterministic_algorithms(True) sums and means always take the CSR path.
_SEGMENT_MODE = os.environ.get('GGLB200_SEGMENT', 'auto')
_ONESHOT_ROW_BYTES = int(os.environ.get('GGLB200_ONESHOT_ROW_BYTES', '16'))


def _segment_path(x, index, N, reduce):
    if x.dtype != torch.float32 or _SEGMENT_MODE == 'csr' or x.shape[0] >= (1 << 32) - 1:
        return 'csr'
    if _SEGMENT_MODE == 'oneshot':
        return 'oneshot'
    exact = reduce == _lib.GGL_MAX
    if not exact and torch.are_deterministic_algorithms_enabled():
        return 'csr'
    row_bytes = (x.numel() // max(x.shape[0], 1)) * x.element_size()
    if row_bytes <= _ONESHOT_ROW_BYTES:
        return 'oneshot'
    if exact or _SEGMENT_MODE != 'fast':
        return 'csr'
    if segment_csr_if_cached(index, N) is not None:
        return 'csr'
    return 'oneshot' if note_sighting(index, N) == 1 else 'csr'


def segment_reduce_oneshot(x2d, ids, N, reduce):
    """Launch gglb200_segment_reduce.  x2d [E, K] fp32 contiguous, ids [E] int64.  Returns (out, arg | None, cnt | None)."""
    lib = _lib.load()
    E, K = x2d.shape
    dev = x2d.device
    with torch.cuda.device(dev):
        out = torch.empty((N, K), dtype=torch.float32, device=dev)
        arg = torch.empty((N, K), dtype=torch.int64, device=dev) if reduce == _lib.GGL_MAX else None
        cnt = torch.empty(N, dtype=torch.int32, device=dev) if reduce != _lib.GGL_SUM else None
        n_bad = torch.empty(1, dtype=torch.int64, device=dev)
        nb = lib.gglb200_segment_reduce_workspace_bytes(reduce, N, K)
        ws = torch.empty(nb, dtype=torch.uint8, device=dev) if nb else None
        _lib.check(lib.gglb200_segment_reduce(reduce, _ptr(x2d), _ptr(ids), _ptr(out), _ptr(arg), N, _ptr(cnt), _ptr(n_bad), E, K, N,
                                              _lib.GGL_F32, _ptr(ws), nb, _stream(dev)), 'segment_reduce')
        bad = int(n_bad.item())             # range errors are raised, never folded silently (one host sync, like the CSR build's)
        if bad:
            raise IndexError('index out of range: %d segment ids outside [0, %d)' % (bad, N))
    return out, arg, cnt


class _SegmentReduce(torch.autograd.Function):
    """SegmentSum / SegmentMean / SegmentMax (torch_ext/src/segment_{sum,mean,max}.cpp)."""

    @staticmethod
    def forward(ctx, x, index, N, reduce):
        _check_segment_args(x, index, N)
        N = int(N)
        x = x.contiguous()
        out_shape = (N,) + tuple(x.shape[1:])
        ctx.x_shape = tuple(x.shape)
        ctx.reduce = reduce
        ctx.N = N
        if x.numel() == 0 or N == 0:                                                  # segment_sum_cpu.cpp:28-30
            ctx.empty = True
            return torch.zeros(out_shape, dtype=x.dtype, device=x.device)
        ctx.empty = False
        E = x.shape[0]
        x2d = x.view(E, -1)
        ctx.index = index if index.dtype == torch.int64 and index.is_contiguous() else index.to(torch.int64).contiguous()
        if _segment_path(x, index, N, reduce) == 'oneshot':
            out, arg, cnt = segment_reduce_oneshot(x2d, ctx.index, N, reduce)
            ctx.csr, ctx.cnt = None, cnt
        else:
            csr = segment_csr_of(index, N)
            out, arg = spmm_csr(csr, csr.perm, x2d, reduce, want_arg=(reduce == _lib.GGL_MAX), arg_empty=N)
            ctx.csr, ctx.cnt = csr, None
        ctx.arg = arg
        return out.view(out_shape)

    @staticmethod
    def _rowptr(ctx, dev):
        """rowptr of the segments: the CSR's, or the prefix sum of the one-shot pass's counts."""
        if ctx.csr is not None:
            return ctx.csr.rowptr
        rp = torch.zeros(ctx.N + 1, dtype=torch.int32, device=dev)
        torch.cumsum(ctx.cnt, 0, dtype=torch.int32, out=rp[1:])
        return rp

    @staticmethod
    def backward(ctx, grad_out):
        if ctx.empty:
            return torch.zeros(ctx.x_shape, dtype=grad_out.dtype, device=grad_out.device), None, None, None
        lib = _lib.load()
        g = grad_out.contiguous()
        E = ctx.x_shape[0]
        K = g.numel() // max(ctx.N, 1)
        dev = g.device
        with torch.cuda.device(dev):
            gx = torch.empty(ctx.x_shape, dtype=g.dtype, device=dev)
            if ctx.reduce == _lib.GGL_MAX:
                rowptr = _SegmentReduce._rowptr(ctx, dev)
                _lib.check(lib.gglb200_segment_max_bwd(_ptr(g), _ptr(ctx.arg), _ptr(rowptr), _ptr(gx), E, ctx.N, K,
                                                       _dtype_code(g), _stream(dev)), 'segment_max_bwd')
            else:
                reduce, c = ctx.reduce, ctx.csr
                rowptr = _SegmentReduce._rowptr(ctx, dev) if reduce == _lib.GGL_MEAN else None
                if reduce == _lib.GGL_MEAN and g.dtype == torch.float32:
                    # g[r] / count[r] once per segment row instead of once per message: same operands, same quotient
                    gs = torch.empty_like(g)
                    _lib.check(lib.gglb200_scale_rows_by_count(_ptr(g), _ptr(rowptr), _ptr(gs), ctx.N, K, _stream(dev)),
                               'scale_rows_by_count')
                    g, reduce = gs, _lib.GGL_SUM
                if c is not None:
                    c.use_on_current_stream()
                if c is not None and (_SEG_BWD == 'expand' or (_SEG_BWD == 'auto' and g.numel() * g.element_size() > _SEG_BWD_L2_BYTES)):
                    _lib.check(lib.gglb200_segment_expand_bwd(reduce, _ptr(g), _ptr(c.rowptr), _ptr(c.perm), _ptr(c.plan),
                                                              _ptr(gx), E, ctx.N, K, _dtype_code(g), _stream(dev)),
                               'segment_expand_bwd')
                else:
                    _lib.check(lib.gglb200_segment_gather_bwd(reduce, _ptr(g), _ptr(ctx.index), _ptr(rowptr),
                                                              _ptr(gx), E, K, _dtype_code(g), _stream(dev)),
                               'segment_gather_bwd')
        return gx, None, None, None


def c_segment_sum(x, index, N):
    return _SegmentReduce.apply(x, index, N, _lib.GGL_SUM)


def c_segment_mean(x, index, N):
    return _SegmentReduce.apply(x, index, N, _lib.GGL_MEAN)


def c_segment_max(x, index, N):
    return _SegmentReduce.apply(x, index, N, _lib.GGL_MAX)


def _num_segments(segment_ids, num_segments):
    if num_segments is None:
        num_segments = int(segment_ids.max().item()) + 1                              # torch.py:74-75 (host sync)
    return num_segments


def unsorted_segment_sum(x, segment_ids, num_segments=None):
    assert x.shape[0] == segment_ids.shape[0], "the length of segment_ids should be equal to data.shape[0]."
    return c_segment_sum(x, segment_ids, _num_segments(segment_ids, num_segments))


def unsorted_segment_mean(x, segment_ids, num_segments=None):
    assert x.shape[0] == segment_ids.shape[0], "the length of segment_ids should be equal to data.shape[0]."
    return c_segment_mean(x, segment_ids, _num_segments(segment_ids, num_segments))


def unsorted_segment_max(x, segment_ids, num_segments=None):
    assert x.shape[0] == segment_ids.shape[0], "the length of segment_ids should be equal to data.shape[0]."
    return c_segment_max(x, segment_ids, _num_segments(segment_ids, num_segments))


segment_sum = unsorted_segment_sum        # torch.py:209-237
segment_mean = unsorted_segment_mean      # torch.py:240-268
segment_max = unsorted_segment_max        # torch.py:271-299


# ------------------------------------------------------------------------------------------------
# gspmm / bspmm
# ------------------------------------------------------------------------------------------------
def _check_spmm_args(index, weight, x):
    require_cuda(x, 'x')
    require_cuda(index, 'index')
    if weight is not None:
        require_cuda(weight, 'weight')
    if not (x.device == index.device) or (weight is not None and weight.device != x.device):
        raise _lib.GglError('Tensor device inconsistent error.')                       # src/gspmm.cpp:46
    if index.dim() != 2 or index.shape[0] != 2:
        raise ValueError('index must have shape [2, num_edges]')
    if weight is not None and weight.dtype != torch.float32:
        raise _lib.GglError('edge weight must be float32 (got %s)' % weight.dtype)     # spmm_sum_cpu.cpp:21-24
    if x.dtype not in (torch.float32, torch.bfloat16, torch.float16):
        raise _lib.GglError('x must be float32 / bfloat16 / float16 (got %s)' % x.dtype)


class _SpMM(torch.autograd.Function):
    """SpMMSum / SpMMMean / SpMMMax (torch_ext/src/gspmm.cpp:26-202).  No gradient for `weight`
    (mark_non_differentiable, src/gspmm.cpp:30,79).  n_dst (extension): number of output rows when the graph is bipartite
    (x holds the n_src source rows); None = x.shape[0] like the reference operator."""

    @staticmethod
    def forward(ctx, index, weight, x, reduce, n_dst=None):
        _check_spmm_args(index, weight, x)
        E = index.shape[1]
        if weight is not None and weight.numel() != E:
            raise ValueError('weight must have num_edges elements')
        x = x.contiguous()
        # weight=None is the reference's torch.ones(E) (torch.py:332-333): 1.0 * x is exact, so the unweighted kernel
        # gives the identical result without streaming E ones
        weight = weight.contiguous().view(-1) if weight is not None else None
        n_src = x.shape[0]
        n_dst = n_src if n_dst is None else int(n_dst)
        g = graph_of(index, n_dst, n_src)
        x2d = x.view(n_src, -1)
        out, arg = spmm_csr(g.fwd, g.fwd.col, x2d, reduce, w=weight, wperm=g.fwd.perm if weight is not None else None,
                            want_arg=(reduce == _lib.GGL_MAX), arg_empty=0)
        ctx.graph, ctx.reduce, ctx.x_shape = g, reduce, tuple(x.shape)
        ctx.has_w = weight is not None
        if weight is not None:
            ctx.save_for_backward(weight)
        ctx.arg = arg
        return out.view((n_dst,) + tuple(x.shape[1:]))

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        weight = ctx.saved_tensors[0] if ctx.has_w else None
        g = ctx.graph
        grad = grad_out.contiguous().view(g.n_dst, -1)
        K = grad.shape[1]
        dev = grad.device
        with torch.cuda.device(dev):
            if ctx.reduce == _lib.GGL_MAX:
                if grad.dtype != torch.float32:
                    raise _lib.GglError('gspmm(max) backward is float32 only')
                gx = torch.empty((g.n_src, K), dtype=grad.dtype, device=dev)
                _lib.check(lib.gglb200_spmm_max_bwd(_ptr(g.bwd.rowptr), _ptr(g.bwd.col), _ptr(weight), _ptr(g.bwd.perm) if ctx.has_w else None,
                                                    _ptr(grad), _ptr(ctx.arg), _ptr(gx), g.n_src, K, _stream(dev)),
                           'spmm_max_bwd')
            else:
                if ctx.reduce == _lib.GGL_MEAN:
                    if grad.dtype != torch.float32:
                        raise _lib.GglError('gspmm(mean) backward is float32 only')
                    scaled = torch.empty_like(grad)
                    _lib.check(lib.gglb200_scale_rows_by_count(_ptr(grad), _ptr(g.fwd.rowptr), _ptr(scaled), g.n_dst, K,
                                                               _stream(dev)), 'scale_rows_by_count')
                    grad = scaled
                gx, _ = spmm_csr(g.bwd, g.bwd.col, grad, _lib.GGL_SUM, w=weight, wperm=g.bwd.perm if ctx.has_w else None)
        return None, None, gx.view(ctx.x_shape), None, None


class _SpMMMaxEdge(torch.autograd.Function):
    """max over in-edges of w[e] * x[src[e]] with the winning EDGE recorded -- `message` + `unsorted_segment_max`
    (message_passing.py:56-59,88-90; src/segment_max.cpp:37-61) as one kernel, no [E,K] message tensor.  Differs from
    c_spmm_max only in the backward on multigraphs: one edge per (dst, k) receives the gradient (the first maximal edge,
    like segment_max's arg), not every parallel edge of the winning source (cpu/spmm_max_cpu.cpp:84-96)."""

    @staticmethod
    def forward(ctx, index, weight, x, n_dst=None):
        lib = _lib.load()
        _check_spmm_args(index, weight, x)
        E = index.shape[1]
        if weight is not None and weight.numel() != E:
            raise ValueError('weight must have num_edges elements')
        x = x.contiguous()
        weight = weight.contiguous().view(-1) if weight is not None else None
        n_src = x.shape[0]
        n_dst = n_src if n_dst is None else int(n_dst)
        g = graph_of(index, n_dst, n_src)
        c = g.fwd
        x2d = x.view(n_src, -1)
        K = x2d.shape[1]
        dev = x.device
        with torch.cuda.device(dev):
            _lib.tune_device(dev.index)
            c.use_on_current_stream()
            out = torch.empty((n_dst, K), dtype=x.dtype, device=dev)
            arg = torch.empty((n_dst, K), dtype=torch.int64, device=dev)
            ws, nb = c.workspace(K, _lib.GGL_MAX)
            _lib.check(lib.gglb200_spmm_csr_max_edge(_ptr(c.rowptr), _ptr(c.col), _ptr(c.perm), _ptr(c.plan), _ptr(weight), _ptr(x2d),
                                                     _ptr(out), _ptr(arg), E, n_dst, n_src, E, K, _dtype_code(x2d), _ptr(ws), nb,
                                                     _stream(dev)), 'spmm_csr_max_edge')
        ctx.graph, ctx.arg, ctx.x_shape = g, arg, tuple(x.shape)
        ctx.has_w = weight is not None
        if weight is not None:
            ctx.save_for_backward(weight)
        ctx.mark_non_differentiable(arg)
        return out.view((n_dst,) + tuple(x.shape[1:])), arg.view((n_dst,) + tuple(x.shape[1:]))

    @staticmethod
    def backward(ctx, grad_out, _grad_arg):
        lib = _lib.load()
        weight = ctx.saved_tensors[0] if ctx.has_w else None
        g = ctx.graph
        grad = grad_out.contiguous().view(g.n_dst, -1)
        if grad.dtype != torch.float32:
            raise _lib.GglError('gspmm_max_edge backward is float32 only')
        K = grad.shape[1]
        dev = grad.device
        with torch.cuda.device(dev):
            t = g.bwd
            t.use_on_current_stream()
            gx = torch.empty((g.n_src, K), dtype=grad.dtype, device=dev)
            _lib.check(lib.gglb200_spmm_max_edge_bwd(_ptr(t.rowptr), _ptr(t.col), _ptr(t.perm), _ptr(weight), _ptr(grad), _ptr(ctx.arg),
                                                     _ptr(gx), g.n_src, K, _stream(dev)), 'spmm_max_edge_bwd')
        return None, None, gx.view(ctx.x_shape), None


class _BSpMM(torch.autograd.Function):
    """BSpMMSum (torch_ext/src/gspmm.cpp:204-260): grad for x AND for the per-head weight."""

    @staticmethod
    def forward(ctx, index, weight, x):
        _check_spmm_args(index, weight, x)
        if x.dim() != 3:
            raise ValueError('bspmm expects x of shape [N, heads, channels]')
        E = index.shape[1]
        N, H, C = x.shape
        if weight.dim() == 1 and H == 1:
            weight = weight.view(E, 1)
        if tuple(weight.shape) != (E, H):
            raise ValueError('bspmm expects weight of shape [num_edges, heads]')
        x = x.contiguous()
        weight = weight.contiguous()
        g = graph_of(index, N)
        out, _ = spmm_csr(g.fwd, g.fwd.col, x.view(N, H * C), _lib.GGL_SUM, w=weight, wperm=g.fwd.perm, heads=H)
        ctx.graph = g
        ctx.save_for_backward(weight, x)
        return out.view(N, H, C)

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        weight, x = ctx.saved_tensors
        g = ctx.graph
        N, H, C = x.shape
        grad = grad_out.contiguous().view(N, H * C)
        dev = grad.device
        gx = gw = None
        with torch.cuda.device(dev):
            if ctx.needs_input_grad[2]:
                gx, _ = spmm_csr(g.bwd, g.bwd.col, grad, _lib.GGL_SUM, w=weight, wperm=g.bwd.perm, heads=H)
                gx = gx.view(N, H, C)
            if ctx.needs_input_grad[1]:
                if x.dtype != torch.float32 or grad.dtype != torch.float32:
                    raise _lib.GglError('bspmm weight gradient is float32 only')
                gw = torch.empty((g.E, H), dtype=torch.float32, device=dev)
                g.fwd.use_on_current_stream()
                _lib.check(lib.gglb200_bspmm_bwd_w(_ptr(g.fwd.rowptr), _ptr(g.fwd.col), _ptr(g.fwd.perm), _ptr(g.fwd.plan),
                                                   _ptr(x), _ptr(grad), _ptr(gw), N, g.E, H, C, _stream(dev)), 'bspmm_bwd_w')
        return None, gw, gx


class _SDDMM(torch.autograd.Function):
    """e[edge, h] = <a[src, h, :], b[dst, h, :]> per edge and head (sampled dense-dense product) -- the logits of
    dot-product attention layers (AGNNConv: agnn_conv.py:64-70 cos(x_src, x_dst); the same kernel is bspmm's weight
    gradient, cpu/bspmm_sum_cpu.cpp:95-107).  No [E, H, C] gather is materialised.  Backward: two per-head weighted
    SpMMs (grad a[s] = sum_e g[e] b[dst], grad b[d] = sum_e g[e] a[src])."""

    @staticmethod
    def forward(ctx, index, a, b):
        lib = _lib.load()
        _check_spmm_args(index, None, a)
        require_cuda(b, 'b')
        if a.dim() != 3 or tuple(a.shape) != tuple(b.shape) or a.dtype != torch.float32 or b.dtype != torch.float32:
            raise _lib.GglError('sddmm expects float32 a, b of one shape [N, heads, channels]')
        N, H, C = a.shape
        a, b = a.contiguous(), b.contiguous()
        g = graph_of(index, N)
        dev = a.device
        with torch.cuda.device(dev):
            _lib.tune_device(dev.index)
            f = g.fwd
            f.use_on_current_stream()
            e = torch.empty((g.E, H), dtype=torch.float32, device=dev)
            _lib.check(lib.gglb200_bspmm_bwd_w(_ptr(f.rowptr), _ptr(f.col), _ptr(f.perm), _ptr(f.plan), _ptr(a), _ptr(b), _ptr(e),
                                               N, g.E, H, C, _stream(dev)), 'sddmm')
        ctx.graph = g
        ctx.save_for_backward(a, b)
        return e

    @staticmethod
    def backward(ctx, ge):
        a, b = ctx.saved_tensors
        g = ctx.graph
        N, H, C = a.shape
        ge = ge.contiguous()
        ga = gb = None
        with torch.cuda.device(a.device):
            if ctx.needs_input_grad[1]:        # rows = sources: transposed CSR, gathers b[dst]
                ga, _ = spmm_csr(g.bwd, g.bwd.col, b.view(N, H * C), _lib.GGL_SUM, w=ge, wperm=g.bwd.perm, heads=H)
                ga = ga.view(N, H, C)
            if ctx.needs_input_grad[2]:        # rows = destinations: forward CSR, gathers a[src]
                gb, _ = spmm_csr(g.fwd, g.fwd.col, a.view(N, H * C), _lib.GGL_SUM, w=ge, wperm=g.fwd.perm, heads=H)
                gb = gb.view(N, H, C)
        return None, ga, gb


def sddmm(index, a, b):
    """e[edge, h] = sum_c a[src(edge), h, c] * b[dst(edge), h, c]; a, b [N, H, C] float32; differentiable in a and b."""
    return _SDDMM.apply(index, a, b)


def c_spmm_sum(index, weight, x, num_dst=None):
    return _SpMM.apply(index, weight, x, _lib.GGL_SUM, num_dst)


def c_spmm_mean(index, weight, x, num_dst=None):
    return _SpMM.apply(index, weight, x, _lib.GGL_MEAN, num_dst)


def c_spmm_max(index, weight, x, num_dst=None):
    return _SpMM.apply(index, weight, x, _lib.GGL_MAX, num_dst)


def gspmm_max_edge(index, weight=None, x=None, num_dst=None, return_arg=False):
    """Fused `message` + `unsorted_segment_max`: out[dst] = max over in-edges of weight[e] * x[src[e]] (empty rows =
    lowest(), like the compiled reference extension), the gradient credits the first maximal EDGE only.  This is what
    SAGEConv(aggr='pool') and aggr='max' layers compute (sage_conv.py:89-92), without the [E, K] message tensor.
    return_arg: also return the int64 [num_dst, ...] original edge ids of the winners (E for empty rows)."""
    out, arg = _SpMMMaxEdge.apply(index, weight, x, num_dst)
    return (out, arg) if return_arg else out


def c_bspmm_sum(index, weight, x):
    return _BSpMM.apply(index, weight, x)


class _GcnPropagate(torch.autograd.Function):
    """out = D_dst^a . A . D_src^b . x with the degree powers GCNConv uses (gcn_conv.py:88-102), evaluated as
    row-scale -> unweighted SpMM -> row-scale: no per-edge weight stream.  Same math as
    gspmm(edge_index, norm_src[src] * norm_dst[dst], x); rounding differs in the last ulp (one product per row instead
    of one per edge), so this path is checked at 1e-5 relative, not bit-exactly."""

    @staticmethod
    def forward(ctx, x, edge_index, norm, bias=None, relu=False):
        lib = _lib.load()
        require_cuda(x, 'x')
        require_cuda(edge_index, 'edge_index')
        if x.dtype not in (torch.float32, torch.bfloat16, torch.float16):
            raise _lib.GglError('gcn_propagate needs a floating x')
        bias_flat = None
        if bias is not None:
            require_cuda(bias, 'bias')
            bias_flat = bias.detach().reshape(-1).to(torch.float32).contiguous()
            if bias_flat.numel() != x[0].numel():
                raise ValueError('bias must have %d elements, got %d' % (x[0].numel(), bias_flat.numel()))
        x = x.contiguous()
        N = x.shape[0]
        g = graph_of(edge_index, N)
        dev = x.device
        x2d = x.view(N, -1)
        K = x2d.shape[1]
        with torch.cuda.device(dev):
            ns = nd = None
            if norm == 'both_src':             # SAGEConv(aggr='gcn'): utils/norm.py:27-30 takes deg(src)^-1/2 on BOTH ends of an edge
                ns = torch.empty(N, dtype=torch.float32, device=dev)
                _lib.check(lib.gglb200_degree_power(_ptr(g.bwd.rowptr), _ptr(ns), N, -0.5, _stream(dev)), 'degree_power')
                nd = ns
            if norm in ('both', 'left'):       # source side: out-degree (degree(src), gcn_conv.py:89)
                ns = torch.empty(N, dtype=torch.float32, device=dev)
                _lib.check(lib.gglb200_degree_power(_ptr(g.bwd.rowptr), _ptr(ns), N, -0.5 if norm == 'both' else -1.0, _stream(dev)), 'degree_power')
            if norm in ('both', 'right'):      # destination side: in-degree (degree(dst), gcn_conv.py:97)
                nd = torch.empty(N, dtype=torch.float32, device=dev)
                _lib.check(lib.gglb200_degree_power(_ptr(g.fwd.rowptr), _ptr(nd), N, -0.5 if norm == 'both' else -1.0, _stream(dev)), 'degree_power')
            xs = x2d
            if ns is not None:
                xs = torch.empty_like(x2d)
                _lib.check(lib.gglb200_scale_rows(_ptr(x2d), _ptr(ns), _ptr(xs), N, K, _dtype_code(x2d), _stream(dev)), 'scale_rows')
            out, _ = spmm_csr(g.fwd, g.fwd.col, xs, _lib.GGL_SUM, out_scale=nd, bias=bias_flat, relu=relu)
        ctx.graph, ctx.ns, ctx.nd, ctx.shape = g, ns, nd, tuple(x.shape)
        ctx.relu, ctx.bias_shape = bool(relu), None if bias is None else tuple(bias.shape)
        ctx.bias_dtype = None if bias is None else bias.dtype
        if relu:
            ctx.save_for_backward(out)
        return out.view(x.shape)

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        g = ctx.graph
        grad = grad_out.contiguous().view(g.N, -1)
        if ctx.relu:
            (out,) = ctx.saved_tensors
            grad = torch.ops.aten.threshold_backward(grad, out.view_as(grad), 0.0)   # one pass: grad where out > 0, else 0
        gbias = None
        if ctx.bias_shape is not None:
            gbias = grad.sum(dim=0, dtype=torch.float32).to(ctx.bias_dtype).reshape(ctx.bias_shape)
        K = grad.shape[1]
        dev = grad.device
        with torch.cuda.device(dev):
            gs = grad
            if ctx.nd is not None:
                gs = torch.empty_like(grad)
                _lib.check(lib.gglb200_scale_rows(_ptr(grad), _ptr(ctx.nd), _ptr(gs), g.N, K, _dtype_code(grad), _stream(dev)), 'scale_rows')
            gx, _ = spmm_csr(g.bwd, g.bwd.col, gs, _lib.GGL_SUM, out_scale=ctx.ns)
        return gx.view(ctx.shape), None, None, gbias, None


def gcn_propagate(edge_index, x, norm='both', bias=None, relu=False):
    """GCNConv's normalised propagate with unit edge weights (gcn_conv.py:78-104 after the Linear), weight-stream free.
    `bias` ([K] or [1, K]) and `relu` fuse `out + bias` (gcn_conv.py:104-106) and the model's ReLU into the row write.
    norm='both_src' is calc_gcn_norm's weighting (utils/norm.py:27-30: deg(src)^-1/2 at both ends), what SAGEConv(aggr='gcn') uses."""
    if norm not in ('both', 'left', 'right', 'none', 'both_src'):
        raise ValueError('Invalid norm value. Must be either "none", "both", "right" or "left".')
    return _GcnPropagate.apply(x, edge_index, norm, bias, relu)


# The unfused decomposition (`message` -> [E, K] tensor -> segment reduce) exists for call shapes no fused kernel covers
# (differentiable weights under mean / max, user-defined message functions).  At ogbn-products scale that tensor is 49.5 GB:
# refuse instead of silently allocating it (GGLB200_MAX_MESSAGE_BYTES raises or lowers the bar; 0 disables the check).
MAX_MESSAGE_BYTES = int(os.environ.get('GGLB200_MAX_MESSAGE_BYTES', str(8 << 30)))


def check_message_bytes(num_edges, row_numel, elem_size, what):
    need = int(num_edges) * int(row_numel) * int(elem_size)
    if MAX_MESSAGE_BYTES and need > MAX_MESSAGE_BYTES:
        raise _lib.GglError('%s would materialise a [%d, %d] message tensor (%.1f GB > GGLB200_MAX_MESSAGE_BYTES = %.1f GB): '
                            'this call shape has no fused kernel; use fixed (non-differentiable) edge weights, aggr sum / mean / '
                            'max on 2-D features, or raise the limit' % (what, num_edges, row_numel, need / 1e9, MAX_MESSAGE_BYTES / 1e9))


def _linear_tf32x3(a, w, bflat, relu):
    """Launch gglb200_linear_tf32x3 on contiguous fp32 a [M, K], w [K, N] (no autograd)."""
    lib = _lib.load()
    M, K = a.shape
    N = w.shape[1]
    dev = a.device
    with torch.cuda.device(dev):
        out = torch.empty((M, N), dtype=torch.float32, device=dev)
        nb = lib.gglb200_linear_workspace_bytes(K, N)
        ws = torch.empty(nb, dtype=torch.uint8, device=dev)
        _lib.check(lib.gglb200_linear_tf32x3(_ptr(a), _ptr(w), _ptr(bflat), int(bool(relu)), _ptr(out), M, K, N, _ptr(ws), nb,
                                             _stream(dev)), 'linear_tf32x3')
    return out


# input gradient of the Linear, g @ W^T: 'tc' = the same tcgen05 kernel on the transposed weight, 'lib' = torch.matmul
_LINEAR_BWD = os.environ.get('GGLB200_LINEAR_BWD', 'tc')     # measured at C3 (256 -> 100): 1.37 ms vs 2.91 ms


class _LinearTC(torch.autograd.Function):
    """out = act(A @ W + bias) on the tcgen05 tensor cores with fp32 accuracy (csrc/linear_tc.cu, 3xTF32, TMEM accumulator):
    the Linear + bias + ReLU that follows the aggregation when a layer aggregates first (gcn_conv.py:79,104-106).  The weight
    gradient (a reduction over the rows) is a library GEMM; the input gradient is the same kind of GEMM as the forward
    (GGLB200_LINEAR_BWD)."""

    @staticmethod
    def forward(ctx, a, w, bias, relu):
        a = a.contiguous()
        w = w.contiguous()
        bflat = None if bias is None else bias.detach().reshape(-1).to(torch.float32).contiguous()
        out = _linear_tf32x3(a, w, bflat, relu)
        ctx.relu = bool(relu)
        ctx.bias_shape = None if bias is None else tuple(bias.shape)
        ctx.save_for_backward(a, w, out if relu else None)
        return out

    @staticmethod
    def backward(ctx, g):
        a, w, out = ctx.saved_tensors
        g = g.contiguous()
        if ctx.relu:
            g = torch.ops.aten.threshold_backward(g, out, 0.0)   # one pass: g where out > 0, else 0
        ga = None
        if ctx.needs_input_grad[0]:
            wt = w.t()
            if _LINEAR_BWD == 'tc' and g.dtype == torch.float32 and linear_tc_supported(g, wt):
                ga = _linear_tf32x3(g, wt.contiguous(), None, False)
            else:
                ga = g @ wt
        gw = a.t() @ g if ctx.needs_input_grad[1] else None
        gb = g.sum(dim=0).reshape(ctx.bias_shape) if ctx.bias_shape is not None and ctx.needs_input_grad[2] else None
        return ga, gw, gb, None


def linear_tc_supported(a, w):
    return (a.is_cuda and a.dtype == torch.float32 and w.dtype == torch.float32 and a.dim() == 2 and w.dim() == 2
            and a.shape[1] == w.shape[0] and a.shape[1] % 4 == 0 and a.shape[1] <= 256 and 0 < w.shape[1] <= 256)


def linear_tc(a, w, bias=None, relu=False):
    """act(a @ w + bias) for a [M, K <= 256] fp32 (K a multiple of 4), w [K, N <= 256] fp32 on the tcgen05 tensor cores at fp32
    accuracy; raises for other shapes (see linear_tc_supported)."""
    require_cuda(a, 'a')
    require_cuda(w, 'w')
    if not linear_tc_supported(a, w):
        raise _lib.GglError('linear_tc: unsupported shapes / dtypes %s @ %s' % (tuple(a.shape), tuple(w.shape)))
    return _LinearTC.apply(a, w, bias, relu)


def weight_needs_grad(weight):
    """True when autograd must reach an edge weight: gspmm cannot provide that gradient (no grad for weight in the reference
    operator either, src/gspmm.cpp:30,79), the callers then use bspmm (sum) or the message + aggregate decomposition."""
    return weight is not None and weight.requires_grad and torch.is_grad_enabled()


def gspmm(index, weight=None, x=None, reduce='sum', num_dst=None):
    """Generalized SpMM: out[dst] (+|mean|max)= weight[e] * x[src]  (torch.py:302-351).
    num_dst (extension; None = x.shape[0] like the reference): output rows of a bipartite block whose sources are the
    rows of x -- SAGEConv's (src_feat, dst_feat) call and sampled mini-batch blocks run fused instead of materialising
    the [E, K] message."""
    if reduce == 'sum':
        return c_spmm_sum(index, weight, x, num_dst)
    elif reduce == 'mean':
        return c_spmm_mean(index, weight, x, num_dst)
    elif reduce == 'max':
        return c_spmm_max(index, weight, x, num_dst)
    raise Exception("Unsupported reduce type, please choose from ['sum', 'mean', 'max'].")


def bspmm(index, weight=None, x=None, reduce='sum'):
    """Per-head weighted SpMM used by GAT: out[dst,h,:] += weight[e,h] * x[src,h,:]  (torch.py:354-365)."""
    if weight is None:
        weight = torch.ones(size=(index.shape[1], x.shape[1]), dtype=torch.float32, device=index.device)
    if reduce == 'sum':
        return c_bspmm_sum(index, weight, x)
    raise Exception("Unsupported reduce type, please choose from ['sum'].")


# ------------------------------------------------------------------------------------------------
# fused edge softmax (gammagl/utils/softmax.py:10-36)
# ------------------------------------------------------------------------------------------------
# segment_softmax has two forms of its segment statistics:
#   csr      one pass over the segment CSR (rows gathered through perm): serial-order sums, long rows combined chunk-wise -- the
#            denominators of 10^6-edge hub segments stay within 1e-6 of exact; deterministic
#   oneshot  sort-free streaming passes with L2 reductions (csrc/oneshot.cu): no CSR, 1.3x / 1.7x faster (C4: 7.8 -> 6.0 ms forward,
#            11.0 -> 6.3 ms backward), sums in arrival order -- the accuracy of the reference's own serial fp32 loop (|sum alpha - 1|
#            up to ~5e-5 on a 2 * 10^6-edge hub), not better
# GGLB200_SOFTMAX_PATH = auto (default: forward csr -- alpha is a model output --, backward's Jacobian row term oneshot; everything
# csr under torch.use_deterministic_algorithms) | oneshot | csr.
_SOFTMAX_PATH = os.environ.get('GGLB200_SOFTMAX_PATH', 'auto')


def _softmax_oneshot(backward):
    if _SOFTMAX_PATH == 'csr':
        return False
    if _SOFTMAX_PATH == 'oneshot':
        return True
    return backward and not torch.are_deterministic_algorithms_enabled()


class _EdgeSoftmax(torch.autograd.Function):
    @staticmethod
    def _workspace(csr, n_rows, H, dev):
        n_chunks = csr.n_chunks if (csr is not None and csr.has_long) else 0
        need = _lib.load().gglb200_edge_softmax_workspace_bytes(n_rows, H, n_chunks)
        return torch.empty(need, dtype=torch.uint8, device=dev), need

    @staticmethod
    def forward(ctx, data, segment_ids, num_segments):
        lib = _lib.load()
        require_cuda(data, 'data')
        require_cuda(segment_ids, 'segment_ids')
        if data.dtype != torch.float32:
            raise _lib.GglError('segment_softmax is float32 only')
        E = data.shape[0]
        data = data.contiguous()
        H = data.numel() // max(E, 1)
        alpha = torch.empty_like(data)
        ctx.shape = tuple(data.shape)
        ctx.n_rows = int(num_segments)
        if E == 0:
            ctx.empty = True
            return alpha
        ctx.empty = False
        ids = segment_ids if segment_ids.dtype == torch.int64 and segment_ids.is_contiguous() else segment_ids.to(torch.int64).contiguous()
        if _softmax_oneshot(False):
            validate_ids(segment_ids, num_segments)
            csr = None
        else:
            csr = segment_csr_of(segment_ids, int(num_segments))
            csr.use_on_current_stream()
        dev = data.device
        with torch.cuda.device(dev):
            _lib.tune_device(dev.index)
            ws, nb = _EdgeSoftmax._workspace(csr, ctx.n_rows, H, dev)
            _lib.check(lib.gglb200_edge_softmax_fwd(_ptr(csr.rowptr) if csr else None, _ptr(csr.perm) if csr else None,
                                                    _ptr(csr.plan) if csr else None, _ptr(ids), _ptr(data), _ptr(alpha), ctx.n_rows, E, H,
                                                    _ptr(ws), nb, _stream(dev)), 'edge_softmax_fwd')
        ctx.csr, ctx.H, ctx.ids = csr, H, ids
        ctx.save_for_backward(alpha)
        return alpha

    @staticmethod
    def backward(ctx, galpha):
        if ctx.empty:
            return torch.zeros(ctx.shape, dtype=galpha.dtype, device=galpha.device), None, None
        lib = _lib.load()
        (alpha,) = ctx.saved_tensors
        galpha = galpha.contiguous()
        gl = torch.empty_like(alpha)
        dev = alpha.device
        csr = None if _softmax_oneshot(True) else ctx.csr
        if csr is None and not _softmax_oneshot(True):      # forward ran sort-free, backward asked for the CSR form
            csr = segment_csr_of(ctx.ids, ctx.n_rows)
        E = alpha.shape[0]
        with torch.cuda.device(dev):
            ws, nb = _EdgeSoftmax._workspace(csr, ctx.n_rows, ctx.H, dev)
            _lib.check(lib.gglb200_edge_softmax_bwd(_ptr(csr.rowptr) if csr else None, _ptr(csr.perm) if csr else None,
                                                    _ptr(csr.plan) if csr else None, _ptr(ctx.ids), _ptr(alpha), _ptr(galpha), _ptr(gl),
                                                    ctx.n_rows, E, ctx.H, _ptr(ws), nb, _stream(dev)), 'edge_softmax_bwd')
        return gl, None, None


# fused GAT backward: 'merged' = one streaming pass over the transposed CSR, grad s_dst by L2 reductions (arrival-order sums);
# 'two_pass' = the deterministic form (second pass over the forward CSR).  torch.use_deterministic_algorithms(True) forces two_pass.
_GAT_BWD = os.environ.get('GGLB200_GAT_BWD', 'merged')


def gat_fusable(H, C):
    lh = C // 4
    return C % 4 == 0 and lh >= 1 and (lh & (lh - 1)) == 0 and H * C <= 128


class _GatAggregate(torch.autograd.Function):
    """Fused GAT aggregation (csrc/gat_fused.cu): logits, edge softmax and the weighted neighbourhood sum in one streaming
    pass per direction; no [E, H] tensor is ever materialised."""

    @staticmethod
    def _ws(csr, H, C, dev):
        if not csr.has_long:
            return None, 0
        nb = _lib.load().gglb200_gat_workspace_bytes(csr.n_chunks, H, C)
        return torch.empty(nb, dtype=torch.uint8, device=dev), nb

    @staticmethod
    def forward(ctx, x, s_src, s_dst, edge_index, slope):
        lib = _lib.load()
        for t, n in ((x, 'x'), (s_src, 's_src'), (s_dst, 's_dst'), (edge_index, 'edge_index')):
            require_cuda(t, n)
        N, H, C = x.shape
        if x.dtype != torch.float32 or not gat_fusable(H, C):
            raise _lib.GglError('gat_aggregate needs float32 x [N,H,C] with C %% 4 == 0, C/4 a power of two, H*C <= 128')
        K = H * C
        Hp = (H + 3) // 4 * 4
        dev = x.device
        g = graph_of(edge_index, N)
        with torch.cuda.device(dev):
            _lib.tune_device(dev.index)
            xcat = torch.zeros((N, K + Hp), dtype=torch.float32, device=dev) if Hp != H else torch.empty((N, K + Hp), dtype=torch.float32, device=dev)
            xcat[:, :K] = x.reshape(N, K)
            xcat[:, K:K + H] = s_src
            sd = s_dst.contiguous().float()
            out = torch.empty((N, K), dtype=torch.float32, device=dev)
            m = torch.empty((N, H), dtype=torch.float32, device=dev)
            s = torch.empty((N, H), dtype=torch.float32, device=dev)
            g.fwd.use_on_current_stream()
            ws, nb = _GatAggregate._ws(g.fwd, H, C, dev)
            _lib.check(lib.gglb200_gat_fwd(_ptr(g.fwd.rowptr), _ptr(g.fwd.col), _ptr(g.fwd.plan), _ptr(xcat), K + Hp, _ptr(sd),
                                           float(slope), _ptr(out), _ptr(m), _ptr(s), N, g.E, H, C, _ptr(ws), nb, _stream(dev)),
                       'gat_fwd')
        ctx.graph, ctx.dims, ctx.slope = g, (N, H, C, Hp), float(slope)
        ctx.save_for_backward(xcat, sd, out, m, s)
        return out.view(N, H, C)

    @staticmethod
    def backward(ctx, grad_out):
        lib = _lib.load()
        xcat, sd, out, m, s = ctx.saved_tensors
        N, H, C, Hp = ctx.dims
        K = H * C
        g = ctx.graph
        dev = xcat.device
        gout = grad_out.contiguous().view(N, K).float()
        with torch.cuda.device(dev):
            D = (out * gout).view(N, H, C).sum(-1)                           # row term of the softmax Jacobian, per node
            merged = _GAT_BWD == 'merged' and not torch.are_deterministic_algorithms_enabled()
            W = K + 4 * H + (4 if merged else 0)
            gcat = torch.empty((N, W), dtype=torch.float32, device=dev)
            gcat[:, :K] = gout
            gcat[:, K:K + 4 * H] = torch.stack([sd, m, 1.0 / s, D], dim=2).view(N, 4 * H)
            gx = torch.empty((N, K), dtype=torch.float32, device=dev)
            gs_src = torch.empty((N, H), dtype=torch.float32, device=dev)
            gs_dst = torch.empty((N, H), dtype=torch.float32, device=dev)
            g.bwd.use_on_current_stream()
            g.fwd.use_on_current_stream()
            ws, nb = _GatAggregate._ws(g.bwd, H, C, dev)
            if merged:
                # ONE pass: every edge adds its d loss / d pre to grad s_dst[i] with an L2 reduction; the destination id rides in the row
                gcat[:, K + 4 * H:] = 0
                gcat[:, K + 4 * H] = torch.arange(N, dtype=torch.int32, device=dev).view(torch.float32)
                _lib.check(lib.gglb200_gat_bwd(_ptr(g.bwd.rowptr), _ptr(g.bwd.col), _ptr(g.bwd.plan), _ptr(gcat), W, _ptr(xcat), K + Hp,
                                               ctx.slope, _ptr(gx), _ptr(gs_src), _ptr(gs_dst), N, g.E, H, C, _ptr(ws), nb, _stream(dev)),
                           'gat_bwd')
                return gx.view(N, H, C), gs_src, gs_dst, None, None
            _lib.check(lib.gglb200_gat_bwd_src(_ptr(g.bwd.rowptr), _ptr(g.bwd.col), _ptr(g.bwd.plan), _ptr(gcat), K + 4 * H, _ptr(xcat),
                                               K + Hp, ctx.slope, _ptr(gx), _ptr(gs_src), N, g.E, H, C, _ptr(ws), nb, _stream(dev)),
                       'gat_bwd_src')
            ws, nb = _GatAggregate._ws(g.fwd, H, C, dev)
            _lib.check(lib.gglb200_gat_bwd_dst(_ptr(g.fwd.rowptr), _ptr(g.fwd.col), _ptr(g.fwd.plan), _ptr(xcat), K + Hp, _ptr(gcat),
                                               K + 4 * H, ctx.slope, _ptr(gs_dst), N, g.E, H, C, _ptr(ws), nb, _stream(dev)),
                       'gat_bwd_dst')
        return gx.view(N, H, C), gs_src, gs_dst, None, None


def gat_head_groups(H, C):
    """Heads per fused launch: the kernel holds H_g * C <= 128 columns per row; heads are independent in GAT, so wider layers
    (the reference default-shaped 8 x 64) run as ceil(H / H_g) launches over column slices.  0 = not fusable at all."""
    lh = C // 4
    if C % 4 != 0 or lh < 1 or (lh & (lh - 1)) != 0 or C > 128:
        return 0
    return min(H, 128 // C)


def gat_fusable_grouped(H, C):
    return gat_head_groups(H, C) > 0


def gat_aggregate(edge_index, x, s_src, s_dst, negative_slope=0.2):
    """out[i,h,:] = sum over in-edges (j -> i) of softmax_i(leaky_relu(s_src[j,h] + s_dst[i,h])) * x[j,h,:]
    -- GATConv's attention + propagate (gat_conv.py:103-112) as one fused op; gradients for x, s_src and s_dst.
    H * C > 128 runs head group by head group (gat_head_groups)."""
    N, H, C = x.shape
    if gat_fusable(H, C):
        return _GatAggregate.apply(x, s_src, s_dst, edge_index, negative_slope)
    hg = gat_head_groups(H, C)
    if hg == 0:
        raise _lib.GglError('gat_aggregate needs float32 x [N,H,C] with C %% 4 == 0, C/4 a power of two, C <= 128 (got H=%d C=%d)' % (H, C))
    outs = []
    for h0 in range(0, H, hg):
        h1 = min(H, h0 + hg)
        outs.append(_GatAggregate.apply(x[:, h0:h1].contiguous(), s_src[:, h0:h1].contiguous(), s_dst[:, h0:h1].contiguous(),
                                        edge_index, negative_slope))
    return torch.cat(outs, dim=1)


def segment_softmax(data, segment_ids, num_segments):
    """softmax of `data` [E, ...] over the edges that share a segment id (gammagl/utils/softmax.py:10-36)."""
    return _EdgeSoftmax.apply(data, segment_ids, num_segments)


__all__ = ['use_ext', 'unsorted_segment_sum', 'unsorted_segment_mean', 'unsorted_segment_max', 'segment_sum',
           'segment_mean', 'segment_max', 'gspmm', 'bspmm', 'c_segment_sum', 'c_segment_mean', 'c_segment_max',
           'c_spmm_sum', 'c_spmm_mean', 'c_spmm_max', 'c_bspmm_sum', 'segment_softmax', 'gcn_propagate', 'gat_aggregate',
           'gspmm_max_edge', 'linear_tc', 'sddmm']
