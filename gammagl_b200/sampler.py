"""GPU neighbour sampling for the mini-batch path: host-side mirror of

    gammagl.ops.sparse.sample_adj        /root/reference/gammagl/ops/sparse/sparse.py + cpu/sample.cpp:10-135
    gammagl.ops.sparse.neighbor_sample   /root/reference/gammagl/ops/sparse/sparse.py:76-93 + cpu/neighbor_sample.cpp:22-125
    gammagl.loader.NeighborSampler       /root/reference/gammagl/loader/neighbor_sampler.py:29-109 (sample())

on the kernels of csrc/sampler.cu (COUNT + scan + PICK per hop).  Same names, argument meaning and return layout; CUDA
tensors only, results stay on the device.  Differences, all deliberate:
  * the random stream is this library's counter-based generator (`seed` argument / torch's default generator): the SET of
    possible outputs and their distribution are the reference's (every neighbour when fanout < 0 or >= degree, uniform with
    replacement, Floyd's algorithm without), the concrete draws are not;
  * new nodes are numbered in ascending global id behind the seeds (the reference numbers them in first-encounter order of
    its serial loop): local ids differ by a permutation of the non-seed nodes, `n_id` maps them back either way.
The blocks a sampler returns are bipartite (size = (n_src, n_dst)); layers consume them through the fused bipartite
aggregation (mpops.gspmm(..., num_dst=n_dst)) -- SAGEConv((x_src, x_dst), edge_index) as in the reference's examples.
"""
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _lib
from .csr import _ptr, _stream, build_csr, require_cuda


def _next_seed(seed):
    if seed is not None:
        return int(seed) & ((1 << 64) - 1)
    return int(torch.randint(0, 2 ** 62, (1,)).item())           # torch's default CPU generator: torch.manual_seed() controls it


def _pick(rowptr, col, idx, fanout, replace, seed):
    """One hop: (out_ptr [n+1], e_id [total], nbr [total]) for the seeds idx (int64, CUDA)."""
    lib = _lib.load()
    n = idx.numel()
    dev = idx.device
    n_nodes = rowptr.numel() - 1
    with torch.cuda.device(dev):
        out_ptr = torch.empty(n + 1, dtype=torch.int64, device=dev)
        n_bad = torch.empty(1, dtype=torch.int64, device=dev)
        nb = lib.gglb200_sample_workspace_bytes(n)
        ws = torch.empty(nb, dtype=torch.uint8, device=dev)
        _lib.check(lib.gglb200_sample_count(_ptr(rowptr), _ptr(idx), n, n_nodes, int(fanout), int(bool(replace)), _ptr(out_ptr),
                                            _ptr(n_bad), _ptr(ws), nb, _stream(dev)), 'sample_count')
        total, bad = int(out_ptr[-1].item()), int(n_bad.item())       # output size: the one host sync per hop
        if bad:
            raise IndexError('sample: %d seed ids outside [0, %d)' % (bad, n_nodes))
        e_id = torch.empty(total, dtype=torch.int64, device=dev)
        nbr = torch.empty(total, dtype=torch.int64, device=dev)
        _lib.check(lib.gglb200_sample_pick(_ptr(rowptr), _ptr(col), _ptr(idx), n, n_nodes, int(fanout), int(bool(replace)),
                                           _next_seed(seed), _ptr(out_ptr), _ptr(e_id), _ptr(nbr), _stream(dev)), 'sample_pick')
    return out_ptr, e_id, nbr


def _relabel(idx, nbr):
    """Local ids: seeds keep their positions 0..n-1, every other sampled node follows in ascending global id.
    Returns (n_id, local ids of nbr)."""
    n = idx.numel()
    sidx, order = torch.sort(idx)
    u = torch.unique(nbr)
    if n:
        pos = torch.searchsorted(sidx, u).clamp_(max=n - 1)
        new = u[sidx[pos] != u]
    else:
        new = u
    n_id = torch.cat([idx, new])
    if nbr.numel() == 0:
        return n_id, nbr.clone()
    if n:
        p1 = torch.searchsorted(sidx, nbr).clamp_(max=n - 1)
        seed_hit = sidx[p1] == nbr
    else:
        p1 = torch.zeros_like(nbr)
        seed_hit = torch.zeros_like(nbr, dtype=torch.bool)
    p2 = torch.searchsorted(new, nbr).clamp_(max=max(new.numel() - 1, 0)) if new.numel() else torch.zeros_like(nbr)
    local = torch.where(seed_hit, order[p1] if n else p1, n + p2)
    return n_id, local


def sample_adj(rowptr, col, idx, num_neighbors, replace=False, seed=None):
    """One hop of neighbour sampling over an int64 CSR (ops/sparse/cpu/sample.cpp:10-135).
    Returns (out_rowptr [len(idx)+1], out_col [E_s] LOCAL ids into n_id, n_id [n_s] global ids with n_id[:len(idx)] == idx,
    e_id [E_s] positions in `col`)."""
    for t, nm in ((rowptr, 'rowptr'), (col, 'col'), (idx, 'idx')):
        require_cuda(t, nm)
    rowptr, col, idx = (t.to(torch.int64).contiguous() for t in (rowptr, col, idx))
    out_ptr, e_id, nbr = _pick(rowptr, col, idx.reshape(-1), num_neighbors, replace, seed)
    n_id, local = _relabel(idx.reshape(-1), nbr)
    return out_ptr, local, n_id, e_id


def neighbor_sample(colptr, row, input_node, num_neighbors, replace=False, directed=True, seed=None):
    """Multi-hop sampling (ops/sparse/cpu/neighbor_sample.cpp:22-125): hop l samples the in-neighbours of the nodes added by
    hop l-1 (hop 0: input_node).  Returns (samples, rows, cols, edges): global node ids (input nodes first), local source /
    destination ids of the sampled edges, their positions in `row`.  directed=False returns instead every edge of the graph
    between two sampled nodes (neighbor_sample.cpp:102-118)."""
    for t, nm in ((colptr, 'colptr'), (row, 'row'), (input_node, 'input_node')):
        require_cuda(t, nm)
    colptr, row, samples = (t.to(torch.int64).contiguous() for t in (colptr, row, input_node.reshape(-1)))
    rows, cols, edges = [], [], []
    begin, end = 0, samples.numel()
    base = None if seed is None else int(seed)
    for hop, fanout in enumerate(num_neighbors):
        frontier = samples[begin:end]
        out_ptr, e_id, nbr = _pick(colptr, row, frontier, fanout, replace, None if base is None else base + 7919 * (hop + 1))
        # relabel against ALL samples so far (a neighbour may be an older sample)
        n_id, local = _relabel(samples, nbr)
        if directed:
            seed_of_edge = torch.repeat_interleave(torch.arange(begin, end, device=samples.device), out_ptr[1:] - out_ptr[:-1])
            rows.append(local)
            cols.append(seed_of_edge)
            edges.append(e_id)
        samples = n_id
        begin, end = end, samples.numel()
    if not directed:
        out_ptr, e_id, nbr = _pick(colptr, row, samples, -1, False, 0)
        ssorted, order = torch.sort(samples)
        pos = torch.searchsorted(ssorted, nbr).clamp_(max=max(samples.numel() - 1, 0))
        keep = ssorted[pos] == nbr if samples.numel() else torch.zeros_like(nbr, dtype=torch.bool)
        dst_local = torch.repeat_interleave(torch.arange(samples.numel(), device=samples.device), out_ptr[1:] - out_ptr[:-1])
        return samples, order[pos][keep], dst_local[keep], e_id[keep]
    cat = lambda xs: torch.cat(xs) if xs else samples.new_empty(0)
    return samples, cat(rows), cat(cols), cat(edges)


@dataclass
class EdgeIndex:            # loader/neighbor_sampler.py:16-20
    edge_index: torch.Tensor
    e_id: Optional[torch.Tensor]
    size: Tuple[int, int]


class NeighborSampler:
    """loader/neighbor_sampler.py:29-109 for a CUDA edge_index: sample(batch) -> (batch_size, n_id, adjs), adjs ordered from the
    outermost hop to the seeds; adjs[l].edge_index is LOCAL [2, E_l] (row 0 = source in n_id space of that hop, row 1 =
    destination = position among that hop's target nodes), size = (n_src, n_dst): a bipartite block for
    SAGEConv((x[:n_src], x[:n_dst]), edge_index)."""

    def __init__(self, edge_index, sample_lists, num_nodes=None, return_e_id=True, seed=None):
        require_cuda(edge_index, 'edge_index')
        self.sizes = list(sample_lists)
        self.return_e_id = return_e_id
        N = int(edge_index.max().item()) + 1 if num_nodes is None else int(num_nodes)
        self.num_nodes = N
        c = build_csr(edge_index[1], edge_index[0], N)                  # rows = destination: the in-neighbour lists (adj_t)
        self.rowptr = c.rowptr.to(torch.int64)
        self.col = c.col.to(torch.int64)
        self.perm = c.perm.to(torch.int64)                              # CSR position -> original edge id
        self._seed = seed
        self._calls = 0

    def sample(self, batch):
        batch = torch.as_tensor(batch, device=self.rowptr.device).to(torch.int64).reshape(-1)
        adjs, n_id = [], batch
        for hop, size in enumerate(self.sizes):
            self._calls += 1
            seed = None if self._seed is None else self._seed + 104729 * self._calls
            out_ptr, local, new_n_id, pos = sample_adj(self.rowptr, self.col, n_id, size, replace=False, seed=seed)
            n_dst = n_id.numel()
            dst = torch.repeat_interleave(torch.arange(n_dst, device=batch.device), out_ptr[1:] - out_ptr[:-1])
            e_id = self.perm[pos] if self.return_e_id else None
            adjs.append(EdgeIndex(torch.stack([local, dst]), e_id, (int(new_n_id.numel()), n_dst)))
            n_id = new_n_id
        adjs = adjs[0] if len(adjs) == 1 else adjs[::-1]
        return batch.numel(), n_id, adjs
