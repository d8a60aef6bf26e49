"""1-D node partition of full-graph aggregation across the GPUs of one node.

New functionality -- the reference has no distributed code at all (no collective call site anywhere under
/root/reference/gammagl).  It exists for graphs that do not fit one GPU (BASELINE.json configs[4], the
ogbn-papers100M shape); graphs that fit run as independent replicas instead (bench.py --gpus N).

Layout: rank g owns the nodes [g*B, min((g+1)*B, N)) with B = ceil(N / P): their feature rows, their output rows
and all edges that point INTO them.  One aggregation pass

    out_g = A_g . x          needs x[j] for every distinct remote source j of A_g   (the halo)

so each pass has exactly one exchange step, an all-to-all(v) of feature rows:

    plan (once per graph)   per peer: sorted unique remote source ids -> request lists; the peers learn what to
                            send (all_to_all of counts + ids); local `col` ids are remapped into the index space
                            [local rows | halo rows] and an ordinary CSR pair is built over it (gammagl_b200.csr)
    forward                 pack x[send_ids] -> all_to_all_single (NCCL over NVLink 5 / NVSwitch) -> SpMM over [x | halo]
    backward                transposed SpMM -> grad of [x | halo] -> reverse all_to_all_single of the halo part ->
                            index_add into the owned rows (sum over peers in ascending rank order: deterministic)

The local reduce is the single-GPU CUDA path (mpops.spmm_csr).  `reduce_fn` exists so that the world-size-2 gloo
tests, which run without a GPU, can exercise plan + exchange with a CPU checker plugged in by the test; the
package itself never supplies a CPU implementation.  With a gloo process group the exchange buffers are staged
through host memory (gloo has no CUDA all-to-all); with NCCL they stay on the device.
"""
import torch
import torch.distributed as dist


def rows_per_rank(N, world):
    return (N + world - 1) // world


def owned_range(N, world, rank):
    B = rows_per_rank(N, world)
    return min(rank * B, N), min((rank + 1) * B, N)


def shard_edges_by_dst(edge_index, N, world, rank):
    """Edges of a replicated edge_index whose destination is owned by `rank` (global ids, original order)."""
    lo, hi = owned_range(N, world, rank)
    m = (edge_index[1] >= lo) & (edge_index[1] < hi)
    return edge_index[:, m]


def _all_to_all(out, inp, out_splits, in_splits, group):
    """all_to_all_single on `inp`'s device with NCCL, staged through the host with gloo."""
    backend = dist.get_backend(group)
    if backend == 'nccl' or not inp.is_cuda:
        dist.all_to_all_single(out, inp, out_splits, in_splits, group=group)
        return out
    o, i = out.cpu(), inp.cpu()
    dist.all_to_all_single(o, i, out_splits, in_splits, group=group)
    out.copy_(o)
    return out


class HaloPlan:
    """Exchange plan + local CSR pair of one rank.  edge_index_local: [2, E_g] GLOBAL ids, every dst owned by us."""

    def __init__(self, edge_index_local, N, group=None):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.N = int(N)
        self.B = rows_per_rank(self.N, self.world)
        self.lo, self.hi = owned_range(self.N, self.world, self.rank)
        self.n_local = self.hi - self.lo
        dev = edge_index_local.device
        src, dst = edge_index_local[0], edge_index_local[1]
        if src.numel() and (int(dst.min()) < self.lo or int(dst.max()) >= self.hi):
            raise ValueError('edge_index_local contains destinations this rank does not own')
        owner = torch.div(src, self.B, rounding_mode='floor')
        # ---- what we need from every peer (sorted unique global ids) ----
        need, need_counts = [], []
        for p in range(self.world):
            ids = torch.unique(src[owner == p]) if p != self.rank else src.new_empty(0)
            need.append(ids)
            need_counts.append(int(ids.numel()))
        self.recv_counts = need_counts                                  # rows received from peer p per pass
        self.n_halo = sum(need_counts)
        # ---- tell the peers: counts, then the id lists ----
        cnt_in = torch.tensor(need_counts, dtype=torch.int64, device=dev)
        cnt_out = torch.empty_like(cnt_in)
        _all_to_all(cnt_out, cnt_in, None, None, group)
        self.send_counts = [int(c) for c in cnt_out.tolist()]           # rows peer p wants from us per pass
        req_in = torch.cat(need) if self.n_halo else src.new_empty(0)
        req_out = src.new_empty(sum(self.send_counts))
        _all_to_all(req_out, req_in, self.send_counts, self.recv_counts, group)
        self.send_ids = (req_out - self.lo).contiguous()                # local row ids, grouped by destination peer
        if self.send_ids.numel() and (int(self.send_ids.min()) < 0 or int(self.send_ids.max()) >= self.n_local):
            raise RuntimeError('halo plan: a peer requested rows this rank does not own')
        # ---- remap sources into [local | halo] ----
        col = torch.empty_like(src)
        local = owner == self.rank
        col[local] = src[local] - self.lo
        off = self.n_local
        for p in range(self.world):
            if need_counts[p]:
                m = owner == p
                col[m] = off + torch.searchsorted(need[p], src[m])
                off += need_counts[p]
        self.n_src = self.n_local + self.n_halo
        self.edge_index_ext = torch.stack([col, dst - self.lo]).contiguous()
        self.E = int(src.numel())
        self._csr = None

    # local CSR pair (rows = owned destinations over n_src sources; transposed: rows = n_src sources)
    def csr(self):
        if self._csr is None:
            from .csr import build_csr
            fwd = build_csr(self.edge_index_ext[1], self.edge_index_ext[0], self.n_local, n_other=self.n_src)
            bwd = build_csr(self.edge_index_ext[0], self.edge_index_ext[1], self.n_src, n_other=self.n_local)
            self._csr = (fwd, bwd)
        return self._csr

    # ---- exchange steps ----
    def gather_halo(self, x_local):
        """x_local [n_local, K] -> x_ext [n_local + n_halo, K] (one all-to-all of feature rows)."""
        K = x_local.shape[1]
        x_ext = x_local.new_empty((self.n_src, K))
        x_ext[:self.n_local] = x_local
        send = x_local.index_select(0, self.send_ids)
        _all_to_all(x_ext[self.n_local:], send, [c for c in self.recv_counts], [c for c in self.send_counts], self.group)
        return x_ext

    def scatter_halo_grad(self, g_ext):
        """g_ext [n_src, K] -> grad of x_local: own part + halo parts returned to their owners and summed there."""
        K = g_ext.shape[1]
        g_local = g_ext[:self.n_local].clone()
        recv = g_ext.new_empty((sum(self.send_counts), K))
        _all_to_all(recv, g_ext[self.n_local:].contiguous(), [c for c in self.send_counts], [c for c in self.recv_counts], self.group)
        g_local.index_add_(0, self.send_ids, recv)
        return g_local


def _cuda_reduce(plan, x_ext, weight, transposed):
    from . import _lib, mpops
    fwd, bwd = plan.csr()
    c = bwd if transposed else fwd
    out, _ = mpops.spmm_csr(c, c.col, x_ext, _lib.GGL_SUM, w=weight, wperm=c.perm if weight is not None else None)
    return out


class _DistSpMMSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_local, weight, plan, reduce_fn):
        x_ext = plan.gather_halo(x_local.contiguous())
        out = reduce_fn(plan, x_ext, weight, False)
        ctx.plan, ctx.reduce_fn = plan, reduce_fn
        ctx.save_for_backward(weight) if weight is not None else None
        ctx.has_w = weight is not None
        return out

    @staticmethod
    def backward(ctx, grad_out):
        weight = ctx.saved_tensors[0] if ctx.has_w else None
        g_ext = ctx.reduce_fn(ctx.plan, grad_out.contiguous(), weight, True)
        return ctx.plan.scatter_halo_grad(g_ext), None, None, None


def dist_gspmm_sum(plan, x_local, weight=None, reduce_fn=None):
    """out_local[i] = sum over in-edges (j -> i) of weight[e] * x[j], x partitioned by rows over the group.
    weight: [E_g] fp32 aligned with the rank's local edge list (or None)."""
    return _DistSpMMSum.apply(x_local, weight, plan, reduce_fn or _cuda_reduce)


# ----------------------------------------------------------------------------------------------------------------------
# Peer-memory variant: ONE kernel does gather + transfer + reduce (no pack / all-to-all / halo buffer)
# ----------------------------------------------------------------------------------------------------------------------
OWNER_SHIFT = 27    # col = owner << 27 | local row : up to 134 M rows per GPU, 16 GPUs


def _encode(ids, B):
    return (torch.div(ids, B, rounding_mode='floor') << OWNER_SHIFT) | (ids % B)


class PeerGraph:
    """1-D node partition whose aggregation kernel reads remote feature rows straight from the owner's HBM over
    NVLink (gglb200_spmm_csr_peer): the bulk async copies that stage rows into shared memory simply carry peer
    addresses, so the transfer overlaps the reduction tile by tile.  Feature shards live in symmetric memory
    (torch.distributed._symmetric_memory: CUDA IPC mappings of every rank's buffer); the only synchronisation per
    pass is a device-side barrier before (all shards written) and after (all shards read).

    Forward uses the in-edges of the owned nodes (rows = local dst, col = encoded global src); backward uses their
    out-edges (rows = local src, col = encoded global dst), obtained once at plan time by an all-to-all of the
    edge list -- so the backward pass is a pull as well and needs no scatter / reduction across GPUs.

    Halo cache (`cache_rows`): a remote row that at least two local edges reference is pulled ONCE per pass into a
    cache region behind the rank's own shard (gglb200_gather_rows_peer: remote reads, no pack on the owner, no
    collective) and the aggregation kernel finds it in local HBM -- `col` of those edges is rewritten to
    owner = self, row = B + slot.  Rows referenced once keep their direct peer address.  NVLink bytes per pass drop
    from (remote edges) x row to (distinct remote rows) x row; the per-row edge order, hence the result bits, do not
    change.  cache_rows: None = every remote row referenced >= 2 times (env GGLB200_PEER_CACHE_ROWS caps it; 0 = off)."""

    def __init__(self, edge_index_local, N, group=None, cache_rows=None):
        import os
        from .csr import build_csr
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.N = int(N)
        self.B = rows_per_rank(self.N, self.world)
        if self.B >= (1 << OWNER_SHIFT) or self.world > 16:
            raise ValueError('PeerGraph supports up to 16 GPUs and 2^27 rows per GPU')
        self.lo, self.hi = owned_range(self.N, self.world, self.rank)
        self.n_local = self.hi - self.lo
        src, dst = edge_index_local[0], edge_index_local[1]
        dev = src.device
        self.device = dev
        self.E = int(src.numel())
        if self.E and (int(src.min()) < 0 or int(src.max()) >= self.N):
            raise IndexError('PeerGraph: source ids outside [0, %d)' % self.N)
        if cache_rows is None:
            env = os.environ.get('GGLB200_PEER_CACHE_ROWS')
            cache_rows = int(env) if env is not None else (1 << OWNER_SHIFT) - self.B - 1
        cache_rows = max(0, min(int(cache_rows), (1 << OWNER_SHIFT) - self.B - 1))
        col_f, self.cache_fwd, self.stats_fwd = self._with_cache(_encode(src, self.B), cache_rows)
        self.fwd = build_csr(dst - self.lo, col_f, self.n_local, n_other=-1)   # col = owner << 27 | row (validated above)
        del col_f
        # out-edges of the owned nodes: bucket the local edge list by the owner of its source and exchange
        owner = torch.div(src, self.B, rounding_mode='floor')
        order = torch.argsort(owner, stable=True)
        counts = torch.bincount(owner, minlength=self.world)
        cnt_out = torch.empty_like(counts)
        _all_to_all(cnt_out, counts, None, None, self.group)
        send_counts = [int(c) for c in counts.tolist()]
        recv_counts = [int(c) for c in cnt_out.tolist()]
        pairs = torch.stack([src[order], dst[order]], dim=1).contiguous()          # [E, 2]
        del owner, order
        got = pairs.new_empty((sum(recv_counts), 2))
        _all_to_all(got, pairs, recv_counts, send_counts, self.group)
        del pairs
        self.E_out = int(got.shape[0])
        col_b, self.cache_bwd, self.stats_bwd = self._with_cache(_encode(got[:, 1], self.B), cache_rows)
        self.bwd = build_csr(got[:, 0] - self.lo, col_b, self.n_local, n_other=-1)
        del col_b, got
        # symmetric buffers hold [B shard rows | cache rows]; every rank allocates the same (maximal) cache capacity
        cap = torch.tensor([self.cache_fwd.numel(), self.cache_bwd.numel()], dtype=torch.int64, device=dev)
        if self.world > 1:
            if dist.get_backend(self.group) == 'nccl':
                dist.all_reduce(cap, op=dist.ReduceOp.MAX, group=self.group)
            else:
                c = cap.cpu()
                dist.all_reduce(c, op=dist.ReduceOp.MAX, group=self.group)
                cap = c
        self.cache_cap = {'x': int(cap[0]), 'g': int(cap[1])}
        self._bufs = {}

    def _with_cache(self, enc, cache_rows):
        """enc: encoded (owner << shift | row) source ids of the local edges.  Returns (col, cache_ids int32 sorted, stats):
        col = enc with every cached remote id replaced by (rank << shift | B + slot)."""
        remote = (enc >> OWNER_SHIFT) != self.rank
        n_remote = int(remote.sum())
        stats = {'edges': int(enc.numel()), 'remote_edges': n_remote, 'cached_rows': 0, 'cached_edges': 0,
                 'direct_remote_edges': n_remote}
        empty = torch.empty(0, dtype=torch.int32, device=enc.device)
        if cache_rows <= 0 or n_remote == 0:
            return enc, empty, stats
        uniq, cnt = torch.unique(enc[remote], return_counts=True)                # sorted ascending = by (owner, row)
        keep = cnt >= 2
        if int(keep.sum()) > cache_rows:                                          # over capacity: keep the hottest rows
            thr = torch.topk(cnt, cache_rows).values[-1]
            keep = cnt > thr
            room = cache_rows - int(keep.sum())
            if room > 0:
                ties = torch.nonzero(cnt == thr).view(-1)[:room]
                keep[ties] = True
        cached = uniq[keep]
        if cached.numel() == 0:
            return enc, empty, stats
        pos = torch.searchsorted(cached, enc).clamp_(max=cached.numel() - 1)
        hit = remote & (cached[pos] == enc)
        col = torch.where(hit, (self.rank << OWNER_SHIFT) | (self.B + pos), enc)
        stats.update(cached_rows=int(cached.numel()), cached_edges=int(hit.sum()),
                     direct_remote_edges=n_remote - int(hit.sum()))
        return col, cached.to(torch.int32).contiguous(), stats

    def _symm(self, tag, K, dtype):
        """Symmetric [B + cache capacity, K] buffer + handle (allocated once per (tag, K, dtype); collective call)."""
        import torch.distributed._symmetric_memory as symm
        key = (tag, K, dtype)
        if key not in self._bufs:
            t = symm.empty((self.B + self.cache_cap[tag], K), dtype=dtype, device=self.device)
            hdl = symm.rendezvous(t, self.group)
            self._bufs[key] = (t, hdl)
        return self._bufs[key]

    def nvlink_bytes_per_pass(self, K, elem_size, tag='x'):
        """Bytes this rank pulls over NVLink in one pass: one row per cached row + one per uncached remote edge."""
        st = self.stats_fwd if tag == 'x' else self.stats_bwd
        return (st['cached_rows'] + st['direct_remote_edges']) * K * elem_size

    def _pull(self, csr, tag, x_local, in_scale, out_scale):
        from . import _lib, mpops
        from .csr import _ptr, _stream
        lib = _lib.load()
        K = x_local.shape[1]
        buf, hdl = self._symm(tag, K, x_local.dtype)
        cache_ids = self.cache_fwd if tag == 'x' else self.cache_bwd
        dev = x_local.device
        with torch.cuda.device(dev):
            _lib.tune_device(dev.index)
            shard = buf[:self.n_local]
            if in_scale is not None:
                _lib.check(lib.gglb200_scale_rows(_ptr(x_local), _ptr(in_scale), _ptr(shard), self.n_local, K,
                                                  mpops._dtype_code(x_local), _stream(dev)), 'scale_rows')
            else:
                shard.copy_(x_local)
            hdl.barrier(channel=0)                       # every shard is complete
            out = torch.empty((self.n_local, K), dtype=x_local.dtype, device=dev)
            ws, nb = csr.workspace(K, _lib.GGL_SUM)
            if mpops.PROFILE is not None:
                ev0 = torch.cuda.Event(enable_timing=True)
                ev0.record()
            table = ctypes_ptr(hdl.buffer_ptrs_dev)
            if cache_ids.numel():                        # deduplicated halo: one NVLink read per cached row
                _lib.check(lib.gglb200_gather_rows_peer(table, OWNER_SHIFT, _ptr(cache_ids), _ptr(buf[self.B:]), cache_ids.numel(),
                                                        K * x_local.element_size(), _stream(dev)), 'gather_rows_peer')
            _lib.check(lib.gglb200_spmm_csr_peer(_lib.GGL_SUM, _ptr(csr.rowptr), _ptr(csr.col), _ptr(csr.plan), None, None, 1,
                                                 table, OWNER_SHIFT, _ptr(out), _ptr(out_scale),
                                                 self.n_local, csr.E, K, mpops._dtype_code(x_local), _ptr(ws), nb, _stream(dev)),
                       'spmm_csr_peer')
            if mpops.PROFILE is not None:
                ev1 = torch.cuda.Event(enable_timing=True)
                ev1.record()
                mpops.PROFILE.append((self.n_local, ev0, ev1))
            hdl.barrier(channel=1)                       # every rank has finished reading the shards
        return out

    def degree_norms(self, power=-0.5):
        """(deg_out^power, deg_in^power) of the owned nodes (0 for isolated ones), from the two local CSRs (computed once per power)."""
        hit = self._norms.get(power) if hasattr(self, '_norms') else None
        if hit is not None:
            return hit
        from . import _lib
        from .csr import _ptr, _stream
        lib = _lib.load()
        with torch.cuda.device(self.device):
            ns = torch.empty(self.n_local, dtype=torch.float32, device=self.device)
            nd = torch.empty(self.n_local, dtype=torch.float32, device=self.device)
            _lib.check(lib.gglb200_degree_power(_ptr(self.bwd.rowptr), _ptr(ns), self.n_local, power, _stream(self.device)), 'degree_power')
            _lib.check(lib.gglb200_degree_power(_ptr(self.fwd.rowptr), _ptr(nd), self.n_local, power, _stream(self.device)), 'degree_power')
        if not hasattr(self, '_norms'):
            self._norms = {}
        self._norms[power] = (ns, nd)
        return ns, nd


def ctypes_ptr(v):
    import ctypes
    return ctypes.c_void_p(int(v))


class _PeerGcnPropagate(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_local, graph, ns, nd):
        ctx.graph, ctx.ns, ctx.nd = graph, ns, nd
        return graph._pull(graph.fwd, 'x', x_local.contiguous(), ns, nd)

    @staticmethod
    def backward(ctx, grad_out):
        g = ctx.graph
        return g._pull(g.bwd, 'g', grad_out.contiguous(), ctx.nd, ctx.ns), None, None, None


def peer_gcn_propagate(graph, x_local, norm='both'):
    """out_local = (D_in^-1/2 A D_out^-1/2 x)[owned rows] for norm='both' (GCNConv with unit edge weights), or the plain
    neighbourhood sum for norm='none'; x is partitioned by rows over the group; forward and backward are single fused
    gather+reduce kernels over peer memory."""
    if norm == 'both':
        ns, nd = graph.degree_norms(-0.5)
    elif norm == 'none':
        ns = nd = None
    else:
        raise ValueError("norm must be 'both' or 'none'")
    return _PeerGcnPropagate.apply(x_local, graph, ns, nd)
