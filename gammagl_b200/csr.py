"""CSR structures the kernels consume, built on the GPU from the int64 COO the reference API delivers.

Host-side mirror of the bookkeeping in /root/reference/gammagl/sparse/storage.py:22-297 (row / rowptr /
col / csr2csc caches), reduced to what the aggregation path needs:

  Csr        rows over `key` (stable by original edge id), `col` gathered from `other`, `perm` = edge ids
  Graph      the forward CSR (rows = destination i, col = source j) and, lazily, the transposed CSR
             (rows = source, col = destination) of one edge_index[2,E]
  caches     an LRU keyed on the *identity of the memory* (storage pointer, offset, shape, strides,
             tensor._version) that also holds a strong reference to the keyed tensor, so the address
             cannot be recycled for different contents while the entry lives.  `_version` is the ONLY
             invalidation signal: writes through `.data`, DLPack aliases or raw CUDA kernels do not bump it --
             call clear_caches() after such a write.  Tensors created under torch.inference_mode() have no
             version counter and are never cached (the CSR is rebuilt per call).
  streams    a Csr is built on the stream current at build time; use_on_current_stream() makes any other
             stream wait for that build, and the long-row partials buffer is kept per stream.
"""
import collections
import ctypes

import torch

from . import _lib


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError('%s must be a torch.Tensor (got %s)' % (name, type(t).__name__))
    if not t.is_cuda:
        raise _lib.GglError('%s is on %s: gammagl_b200 is a CUDA-only backend (no CPU fallback)' % (name, t.device))


class Csr:
    """rowptr[N+1] i32, col[E] i32 (or None), perm[E] i32, plan (long-row work list), has_long."""
    __slots__ = ('rowptr', 'col', 'perm', 'plan', 'n_rows', 'E', 'has_long', 'n_chunks', 'device', '_ws', 'build_event')

    def workspace(self, K, reduce):
        """Partials buffer for split rows; empty when the graph has no row longer than LONG_ROW."""
        if not self.has_long:
            return None, 0
        need = _lib.load().gglb200_spmm_workspace_bytes_for_chunks(self.n_chunks, K, reduce)
        key = torch.cuda.current_stream(self.device).cuda_stream       # one buffer per stream: concurrent streams never share partials
        ws = self._ws.get(key)
        if ws is None or ws.numel() < need:
            ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            self._ws[key] = ws
        return ws, ws.numel()

    def degrees(self):
        return self.rowptr[1:] - self.rowptr[:-1]

    def use_on_current_stream(self):
        """Order the current stream after the build of this CSR (no-op cost when it is the build stream)."""
        ev = getattr(self, 'build_event', None)
        if ev is not None:
            torch.cuda.current_stream(self.device).wait_event(ev)


def build_csr(key, other, n_rows, n_other=None):
    """key/other: int64 [E] CUDA tensors (other may be None) -> Csr.  One host sync (long-row count + range errors).
    n_other: number of rows `other` may address (None: n_rows; < 0: unchecked, `other` carries an encoding).
    Raises IndexError when a key is outside [0, n_rows) or an `other` id outside [0, n_other) -- the reference's CPU
    operators raise as well (TORCH_CHECK_INDEX idx < N, cpu/segment_max_cpu.cpp:50); nothing is clamped silently."""
    lib = _lib.load()
    require_cuda(key, 'index')
    if key.dtype != torch.int64:
        key = key.to(torch.int64)
    key = key.contiguous()
    if other is not None:
        if other.dtype != torch.int64:
            other = other.to(torch.int64)
        other = other.contiguous()
    E = key.numel()
    dev = key.device
    with torch.cuda.device(dev):
        c = Csr()
        c.n_rows, c.E, c.device, c._ws = int(n_rows), int(E), dev, {}
        c.rowptr = torch.empty(n_rows + 1, dtype=torch.int32, device=dev)
        c.perm = torch.empty(E, dtype=torch.int32, device=dev)
        c.col = torch.empty(E, dtype=torch.int32, device=dev) if other is not None else None
        c.plan = torch.empty(lib.gglb200_csr_plan_ints(E), dtype=torch.int32, device=dev)
        ws_bytes = lib.gglb200_csr_build_workspace_bytes(E, n_rows)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        n_other = int(n_rows) if n_other is None else int(n_other)
        _lib.check(lib.gglb200_csr_build(_ptr(key), _ptr(other), E, n_rows, n_other, _ptr(c.rowptr), _ptr(c.col), _ptr(c.perm),
                                         _ptr(c.plan), _ptr(ws), ws_bytes, _stream(dev)), 'csr_build')
        head = c.plan[:8].tolist()                                      # the one host sync of the build
        if head[6] or head[7]:
            raise IndexError('index out of range: %d row ids outside [0, %d), %d column ids outside [0, %d)'
                             % (head[6], n_rows, head[7], n_other))
        c.n_chunks, n_long = head[0], head[1]
        c.has_long = n_long > 0
        c.build_event = torch.cuda.Event()
        c.build_event.record(torch.cuda.current_stream(dev))
    return c


class Graph:
    """Forward (by destination) and transposed (by source) CSR of one edge_index[2,E].
    Bipartite graphs (SAGEConv's (src_feat, dst_feat) blocks, sage_conv.py:75-92; sampled mini-batch blocks): n_src source
    rows and n_dst destination rows; N (= n_dst) keeps the square-graph meaning for the callers that only know one size."""

    def __init__(self, edge_index, num_nodes, num_src=None):
        require_cuda(edge_index, 'edge_index')
        if edge_index.dim() != 2 or edge_index.shape[0] != 2:
            raise ValueError('edge_index must have shape [2, E], got %s' % (tuple(edge_index.shape),))
        self.edge_index = edge_index
        self.N = self.n_dst = int(num_nodes)
        self.n_src = self.n_dst if num_src is None else int(num_src)
        self.E = int(edge_index.shape[1])
        self._fwd = None
        self._bwd = None

    @property
    def fwd(self):  # rows = dst (edge_index[1]), col = src
        if self._fwd is None:
            self._fwd = build_csr(self.edge_index[1], self.edge_index[0], self.n_dst, n_other=self.n_src)
        return self._fwd

    @property
    def bwd(self):  # rows = src, col = dst
        if self._bwd is None:
            self._bwd = build_csr(self.edge_index[0], self.edge_index[1], self.n_src, n_other=self.n_dst)
        return self._bwd


class _Lru:
    """Identity-keyed cache (storage pointer + version counter).  `by_content`: on an identity miss, an entry whose tensor has the
    same shape / dtype / device and EQUAL contents (torch.equal: one streaming compare, ~0.7 ms for the 2 GB edge list of C3, against
    two 6 ms sorts) is reused -- reference layers rebuild their edge lists every forward (sage_conv.py:83-85 add_self_loops, model
    code that concatenates self loops), which would otherwise mean a CSR build per call.  At most two candidates are compared; an
    entry whose tensor was modified in place after insertion (version moved on) is never matched.  GGLB200_CONTENT_CACHE=0 disables."""

    def __init__(self, capacity, by_content=False):
        self.capacity = capacity
        self.by_content = by_content and _os.environ.get('GGLB200_CONTENT_CACHE', '1') != '0'
        self.d = collections.OrderedDict()
        self.content_hits = 0
        self.match_cpu = False     # product code only ever keys CUDA tensors; the CPU unit test of the matching rule sets this

    @staticmethod
    def key_of(t, extra):
        return (t.untyped_storage().data_ptr(), t.storage_offset(), tuple(t.shape), tuple(t.stride()), t._version,
                t.dtype, t.device.index, extra)

    def get(self, t, extra, make):
        if t.is_inference():
            # inference tensors carry no version counter (t._version raises): build per call, never cache
            return make()
        k = self.key_of(t, extra)
        hit = self.d.get(k)
        if hit is not None:
            self.d.move_to_end(k)
            return hit[1]
        val = None
        if self.by_content and (t.is_cuda or self.match_cpu):
            tried = 0
            for k2 in reversed(self.d):
                t2 = self.d[k2][0]
                if k2[2:4] + k2[5:] != k[2:4] + k[5:] or t2._version != k2[4]:
                    continue
                tried += 1
                if torch.equal(t2, t):
                    val = self.d[k2][1]
                    self.content_hits += 1
                    break
                if tried == 2:
                    break
        if val is None:
            val = make()
        self.d[k] = (t, val)  # strong ref to `t` pins its storage => the key cannot alias new contents
        while len(self.d) > self.capacity:
            self.d.popitem(last=False)
        return val

    def clear(self):
        self.d.clear()


import os as _os

# hetero models keep one graph per relation alive: capacities are configurable (GGLB200_GRAPH_CACHE / GGLB200_SEGMENT_CACHE)
_graphs = _Lru(int(_os.environ.get('GGLB200_GRAPH_CACHE', '16')), by_content=True)
_segments = _Lru(int(_os.environ.get('GGLB200_SEGMENT_CACHE', '16')), by_content=True)


def graph_of(edge_index, num_nodes, num_src=None):
    """Cached Graph for an edge_index tensor (built on first use; CSR build cost is paid once per graph).
    num_src: number of source rows when it differs from num_nodes (= destination rows): bipartite blocks."""
    n_src = int(num_nodes) if num_src is None else int(num_src)
    return _graphs.get(edge_index, (int(num_nodes), n_src), lambda: Graph(edge_index, num_nodes, n_src))


def segment_csr_of(segment_ids, num_segments):
    """Cached CSR over segment ids (col is None: the kernels gather through perm)."""
    return _segments.get(segment_ids, int(num_segments), lambda: build_csr(segment_ids, None, num_segments))


# ids seen by the sort-free path: key -> number of sightings (the CSR is built from the second sighting on)
_sightings = _Lru(int(_os.environ.get('GGLB200_SEGMENT_CACHE', '16')) * 4)


def segment_csr_if_cached(segment_ids, num_segments):
    """The cached CSR of these ids, or None (never builds)."""
    if segment_ids.is_inference():
        return None
    hit = _segments.d.get(_Lru.key_of(segment_ids, int(num_segments)))
    if hit is None:
        return None
    _segments.d.move_to_end(_Lru.key_of(segment_ids, int(num_segments)))
    return hit[1]


def note_sighting(segment_ids, num_segments):
    """How many times (including this one) these very ids have been reduced without a CSR: 1 = first time."""
    if segment_ids.is_inference():
        return 1
    box = _sightings.get(segment_ids, int(num_segments), lambda: [0])
    box[0] += 1
    return box[0]


_validated = _Lru(64)


def validate_ids(ids, n):
    """Raise IndexError unless every id lies in [0, n).  One min/max pass + host sync per ids tensor (cached by tensor identity like
    the CSRs): the sort-free kernels skip out-of-range ids instead of dereferencing them, this makes the error loud."""
    if ids.numel() == 0:
        return True

    def check():
        mn, mx = torch.aminmax(ids)
        mn, mx = int(mn.item()), int(mx.item())
        if mn < 0 or mx >= int(n):
            raise IndexError('index out of range: ids span [%d, %d], expected [0, %d)' % (mn, mx, int(n)))
        return True
    return _validated.get(ids, int(n), check)


_derived = _Lru(int(_os.environ.get('GGLB200_DERIVED_CACHE', '8')))   # edge lists derived from a caller's edge_index (self loops ...)


def derived_of(edge_index, tag, make):
    """A tensor / tuple derived from `edge_index` (same identity + version => same value): built once, so the graph cache that is
    keyed on tensor identity sees the SAME derived edge list on every forward instead of a fresh one (= a CSR rebuild per call)."""
    return _derived.get(edge_index, tag, make)


def clear_caches():
    _graphs.clear()
    _segments.clear()
    _sightings.clear()
    _validated.clear()
    _derived.clear()
