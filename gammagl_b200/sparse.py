"""Host-side mirror of the sparse structures the path touches (`gammagl.sparse`, `gammagl.ops.sparse`), on the GPU.

    ind2ptr / ptr2ind     /root/reference/gammagl/ops/sparse/sparse.py:38-75 (cpu/convert.cpp:58-128)
    CSRAdj.from_edges     /root/reference/gammagl/sparse/sparse_adj.py:47-67   (+ the `matmul` the reference leaves
                          commented out, sparse_adj.py:122-129: reduced_h = segment_sum(h[col] * w, row))
    SparseGraph           /root/reference/gammagl/sparse/graph.py:14-132 / storage.py:22-297: COO <-> CSR <-> CSC views
                          (row, rowptr, col, value, colptr, csr2csc) and `t()`, plus `matmul` = the SpMM north_star asks
                          to be added on these types.

SparseGraph orders its elements by `row * ncols + col` like SparseStorage (storage.py:124-140, a host numpy argsort there),
with the GPU two-key radix sort of gglb200_sort_edges; equal (row, col) pairs keep their input order.  CSRAdj sorts by
the first index only, as sparse_adj.py:47-67 does.  int64 views are materialised lazily from the int32 arrays the kernels use.
"""
import torch

from . import _lib, mpops, utils
from .csr import Graph, _ptr, _stream, build_csr, require_cuda


def ind2ptr(ind, M):
    """CSR pointer [M+1] (int64) of a sorted row index."""
    require_cuda(ind, 'ind')
    ind = ind.to(torch.int64).contiguous()
    ptr = torch.empty(M + 1, dtype=torch.int64, device=ind.device)
    with torch.cuda.device(ind.device):
        _lib.check(_lib.load().gglb200_ind2ptr(_ptr(ind), ind.numel(), M, _ptr(ptr), _stream(ind.device)), 'ind2ptr')
    return ptr


def ptr2ind(ptr, E):
    """Row id (int64) of each of the E elements of a CSR pointer."""
    require_cuda(ptr, 'ptr')
    ptr = ptr.to(torch.int64).contiguous()
    ind = torch.empty(E, dtype=torch.int64, device=ptr.device)
    with torch.cuda.device(ptr.device):
        _lib.check(_lib.load().gglb200_ptr2ind(_ptr(ptr), ptr.numel() - 1, E, _ptr(ind), _stream(ptr.device)), 'ptr2ind')
    return ind


class CSRAdj(object):
    """sparse_adj.py:47-67: adjacency sorted by the FIRST index `u` (rows = u, columns = v)."""

    @classmethod
    def from_edges(cls, u, v, num_nodes):
        self = cls()
        require_cuda(u, 'u')
        require_cuda(v, 'v')
        self.num_nodes = int(num_nodes)
        self._csr = build_csr(u, v, self.num_nodes)                      # stable: ties keep edge order (numpy argsort does not promise it)
        self._u, self._v = u, v
        return self

    @property
    def degree(self):
        return self._csr.degrees().to(torch.int64)

    @property
    def _indptr(self):
        return self._csr.rowptr.to(torch.int64)

    @property
    def _sorted_eid(self):
        return self._csr.perm.to(torch.int64)

    @property
    def _sorted_u(self):
        return ptr2ind(self._indptr, self._csr.E)

    @property
    def _sorted_v(self):
        return self._csr.col.to(torch.int64)

    def matmul(self, h, edge_weight=None, reduce='sum'):
        """sparse_adj @ h: out[u] = reduce over edges (u, v) of w * h[v]  (the reference's commented-out matmul)."""
        require_cuda(h, 'h')
        h2 = h.contiguous().view(h.shape[0], -1)
        w = None if edge_weight is None else edge_weight.contiguous().view(-1).float()
        out, _ = mpops.spmm_csr(self._csr, self._csr.col, h2, mpops._REDUCE[reduce], w=w,
                                wperm=self._csr.perm if w is not None else None, n_rows=self.num_nodes)
        return out.view((self.num_nodes,) + tuple(h.shape[1:]))


class SparseGraph(object):
    """graph.py:14-132 over storage.py:22-297, reduced to the views the aggregation path uses."""

    def __init__(self, row, col, value=None, sparse_sizes=None):
        require_cuda(row, 'row')
        require_cuda(col, 'col')
        M = int(row.max().item()) + 1 if sparse_sizes is None else int(sparse_sizes[0])
        N = int(col.max().item()) + 1 if sparse_sizes is None else int(sparse_sizes[1])
        self._sizes = (M, N)
        self._row, self._col, self._value = row, col, value
        self._by_row = None
        self._by_col = None

    @classmethod
    def from_edge_index(cls, edge_index, edge_attr=None, sparse_sizes=None):
        return cls(edge_index[0], edge_index[1], edge_attr, sparse_sizes)

    def sparse_sizes(self):
        return self._sizes

    def nnz(self):
        return int(self._row.numel())

    @staticmethod
    def _sorted_csr(major, minor, n_major, n_minor):
        sm, sn, perm, _, _ = utils.sort_pairs(major, minor, n_major, n_minor)   # by (major, minor), stable
        c = build_csr(sm, sn, n_major, n_other=n_minor)   # input already in order: the stable row sort keeps it (its perm is the identity)
        c.perm = perm                            # sorted position -> input element
        return c

    def _rows(self):
        if self._by_row is None:
            self._by_row = self._sorted_csr(self._row, self._col, self._sizes[0], self._sizes[1])
        return self._by_row

    def _cols(self):
        if self._by_col is None:
            self._by_col = self._sorted_csr(self._col, self._row, self._sizes[1], self._sizes[0])
        return self._by_col

    def coo(self):
        c = self._rows()
        value = None if self._value is None else self._value[c.perm.to(torch.int64)]
        return ptr2ind(c.rowptr.to(torch.int64), c.E), c.col.to(torch.int64), value

    def csr(self):
        c = self._rows()
        value = None if self._value is None else self._value[c.perm.to(torch.int64)]
        return c.rowptr.to(torch.int64), c.col.to(torch.int64), value

    def csc(self):
        c = self._cols()
        value = None if self._value is None else self._value[c.perm.to(torch.int64)]
        return c.rowptr.to(torch.int64), c.col.to(torch.int64), value

    def csr2csc(self):
        """Permutation that takes CSR element order to CSC element order (storage.py:200-217)."""
        inv = torch.empty_like(self._rows().perm, dtype=torch.int64)
        inv[self._rows().perm.to(torch.int64)] = torch.arange(self.nnz(), device=inv.device)
        return inv[self._cols().perm.to(torch.int64)]

    def t(self):
        g = SparseGraph(self._col, self._row, self._value, (self._sizes[1], self._sizes[0]))
        g._by_row, g._by_col = self._by_col, self._by_row
        return g

    def matmul(self, x, reduce='sum'):
        """A @ x with A[row, col] = value: out[row] = reduce of value * x[col]."""
        require_cuda(x, 'x')
        c = self._rows()
        x2 = x.contiguous().view(x.shape[0], -1)
        w = None if self._value is None else self._value.contiguous().view(-1).float()
        out, _ = mpops.spmm_csr(c, c.col, x2, mpops._REDUCE[reduce], w=w, wperm=c.perm if w is not None else None,
                                n_rows=self._sizes[0])
        return out.view((self._sizes[0],) + tuple(x.shape[1:]))
