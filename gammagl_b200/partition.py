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
            fwd = build_csr(self.edge_index_ext[1], self.edge_index_ext[0], self.n_local)
            bwd = build_csr(self.edge_index_ext[0], self.edge_index_ext[1], self.n_src)
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
