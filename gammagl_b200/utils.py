"""Host-side mirror of the edge-list utilities that sit immediately before the aggregation path,
on the GPU two-key sort of `gglb200_sort_edges` instead of a host argsort.

    sort_edge_index   /root/reference/gammagl/utils/sort_edge_index.py:5-45
    coalesce          /root/reference/gammagl/utils/coalesce.py:8-78
    to_undirected     /root/reference/gammagl/utils/undirected.py:62-109 (concat both directions, coalesce)
    is_undirected     /root/reference/gammagl/utils/undirected.py:6-59   (row-sorted == transposed column-sorted, incl. attributes)
    maybe_num_nodes   /root/reference/gammagl/utils/num_nodes.py

Same names, argument meaning and return values.  Differences, all deliberate:
  * CUDA tensors only, results stay on the device (the reference round-trips through numpy);
  * the sort is STABLE: duplicate edges keep their input order (the reference's argsort leaves it unspecified), so
    duplicate attributes are merged in ascending input order;
  * `coalesce` merges every tensor of a list `edge_attr` (the reference only merges a single tensor, coalesce.py:75-76)
    and, like the reference, always merges by sum whatever `reduce` says (coalesce.py:76 calls segment_sum).
"""
import torch

from . import _lib, mpops
from .csr import _ptr, _stream, require_cuda


def maybe_num_nodes(edge_index, num_nodes=None):
    if num_nodes is not None:
        return int(num_nodes)
    return int(edge_index.max().item()) + 1 if edge_index.numel() > 0 else 0


def sort_pairs(major, minor, n_major, n_minor, want_sorted=True, want_groups=False):
    """Stable sort of the pairs (major[i], minor[i]) by major * n_minor + minor.
    Returns (sorted_major, sorted_minor, perm int32, group int64 | None, n_unique device scalar | None)."""
    lib = _lib.load()
    require_cuda(major, 'major')
    require_cuda(minor, 'minor')
    major = major.to(torch.int64).contiguous()
    minor = minor.to(torch.int64).contiguous()
    E, dev = major.numel(), major.device
    with torch.cuda.device(dev):
        sm = torch.empty(E, dtype=torch.int64, device=dev) if want_sorted else None
        sn = torch.empty(E, dtype=torch.int64, device=dev) if want_sorted else None
        perm = torch.empty(E, dtype=torch.int32, device=dev)
        group = torch.empty(E, dtype=torch.int64, device=dev) if want_groups else None
        n_unique = torch.zeros(1, dtype=torch.int64, device=dev) if want_groups else None
        n_bad = torch.empty(1, dtype=torch.int64, device=dev)
        ws_bytes = lib.gglb200_sort_edges_workspace_bytes(E, n_major, n_minor)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(lib.gglb200_sort_edges(_ptr(major), _ptr(minor), E, int(n_major), int(n_minor), _ptr(sm), _ptr(sn), _ptr(perm),
                                          _ptr(group), _ptr(n_unique), _ptr(n_bad), _ptr(ws), ws_bytes, _stream(dev)), 'sort_edges')
        bad = int(n_bad.item())                                # host sync: range errors are raised, never returned as data
        if bad:
            raise IndexError('sort_pairs: %d ids outside [0, %d) x [0, %d)' % (bad, int(n_major), int(n_minor)))
    return sm, sn, perm, group, n_unique


def permute_rows(t, perm):
    """t[perm] along dim 0 for an int32 permutation (gglb200_permute_rows; rows of any width)."""
    require_cuda(t, 'edge_attr')
    t = t.contiguous()
    out = torch.empty_like(t)
    n = t.shape[0]
    if t.numel() == 0:
        return out
    row_bytes = t.numel() // n * t.element_size()
    with torch.cuda.device(t.device):
        _lib.check(_lib.load().gglb200_permute_rows(_ptr(t), _ptr(perm), _ptr(out), n, row_bytes, _stream(t.device)), 'permute_rows')
    return out


def unique_pairs(sorted_major, sorted_minor, group, n_unique):
    """First pair of every duplicate group of a sorted pair list (gglb200_unique_edges; coalesce.py:69)."""
    dev, E = sorted_major.device, sorted_major.numel()
    um = torch.empty(n_unique, dtype=torch.int64, device=dev)
    un = torch.empty(n_unique, dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().gglb200_unique_edges(_ptr(sorted_major), _ptr(sorted_minor), _ptr(group), E, _ptr(um), _ptr(un),
                                                    _stream(dev)), 'unique_edges')
    return um, un


def _map_attr(edge_attr, fn):
    if edge_attr is None:
        return None
    if isinstance(edge_attr, torch.Tensor):
        return fn(edge_attr)
    return [fn(e) for e in edge_attr]


def sort_edge_index(edge_index, edge_attr=None, num_nodes=None, sort_by_row=True):
    """Row-wise (or column-wise) sort of edge_index[2,E] (sort_edge_index.py:31-45)."""
    require_cuda(edge_index, 'edge_index')
    N = maybe_num_nodes(edge_index, num_nodes)
    a, b = (0, 1) if sort_by_row else (1, 0)
    sm, sn, perm, _, _ = sort_pairs(edge_index[a], edge_index[b], N, N)
    out = torch.empty((2, sm.numel()), dtype=torch.int64, device=edge_index.device)
    out[a], out[b] = sm, sn
    if edge_attr is None:
        return out
    return out, _map_attr(edge_attr, lambda t: permute_rows(t, perm))


def coalesce(edge_index, edge_attr=None, num_nodes=None, reduce='add', is_sorted=False, sort_by_row=True):
    """Sort edge_index and merge duplicate edges, summing their attributes (coalesce.py:41-78).
    `is_sorted` is accepted for signature parity; the stable sort leaves sorted input unchanged."""
    require_cuda(edge_index, 'edge_index')
    N = maybe_num_nodes(edge_index, num_nodes)
    a, b = (0, 1) if sort_by_row else (1, 0)
    sm, sn, perm, group, n_unique = sort_pairs(edge_index[a], edge_index[b], N, N, want_groups=True)
    E, dev = sm.numel(), edge_index.device
    U = int(n_unique.item())                                   # output size: the one host sync (coalesce.py:64 has it too)
    attr = _map_attr(edge_attr, lambda t: permute_rows(t, perm))
    out = torch.empty((2, U), dtype=torch.int64, device=dev)
    if U == E:                                                 # no duplicates (coalesce.py:63-66)
        out[a], out[b] = sm, sn
        return out if edge_attr is None else (out, attr)
    out[a], out[b] = unique_pairs(sm, sn, group, U)
    if edge_attr is None:
        return out
    return out, _map_attr(attr, lambda t: mpops.unsorted_segment_sum(t, group, U))   # coalesce.py:76


def to_undirected(edge_index, edge_attr=None, num_nodes=None, reduce='add'):
    """Add the reverse of every edge and merge duplicates (undirected.py:96-109): coalesce(cat(ei, ei[::-1]))."""
    require_cuda(edge_index, 'edge_index')
    if isinstance(edge_attr, int):                             # undirected.py:96-98: to_undirected(edge_index, num_nodes)
        edge_attr, num_nodes = None, edge_attr
    row, col = edge_index[0], edge_index[1]
    both = torch.stack([torch.cat([row, col]), torch.cat([col, row])])
    attr = _map_attr(edge_attr, lambda t: torch.cat([t, t], dim=0))
    return coalesce(both, attr, num_nodes, reduce)


def is_undirected(edge_index, edge_attr=None, num_nodes=None):
    """True iff the row-sorted edge list equals the transposed column-sorted one, attributes included (undirected.py:35-59)."""
    require_cuda(edge_index, 'edge_index')
    N = maybe_num_nodes(edge_index, num_nodes)
    attrs = [] if edge_attr is None else ([edge_attr] if isinstance(edge_attr, torch.Tensor) else list(edge_attr))
    ei1, attr1 = sort_edge_index(edge_index, attrs, N, sort_by_row=True)
    ei2, attr2 = sort_edge_index(ei1, attr1, N, sort_by_row=False)
    return (bool((ei1[0] == ei2[1]).all()) and bool((ei1[1] == ei2[0]).all())
            and all(bool((a == b).all()) for a, b in zip(attr1, attr2)))
