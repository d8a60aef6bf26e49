"""GPU test of the partitioned (N>1) path with the real kernels: two ranks share cuda:0, the halo exchange runs over
a gloo group staged through host memory (NCCL refuses two ranks on one device; on a multi-GPU box bench.py
--workload c5part exercises the NCCL exchange).  Result = single-process oracle on the whole graph."""
import os
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, init_file, N, E, K, seed, q):
    from gammagl_b200 import partition as P
    torch.cuda.set_device(0)
    dist.init_process_group('gloo', init_method='file://' + init_file, rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(seed)
        ei = torch.randint(0, N, (2, E), generator=g)
        w = torch.rand(E, generator=g)
        x = torch.randn(N, K, generator=g)
        gout = torch.randn(N, K, generator=g)
        lo, hi = P.owned_range(N, world, rank)
        ei[1, :3000] = lo + 1                                           # a split row (> 1024 edges) on every rank
        m = (ei[1] >= lo) & (ei[1] < hi)
        plan = P.HaloPlan(ei[:, m].cuda(), N)
        x_local = x[lo:hi].cuda().requires_grad_(True)
        out = P.dist_gspmm_sum(plan, x_local, w[m].cuda())
        out.backward(gout[lo:hi].cuda())
        q.put((rank, out.detach().cpu().numpy(), x_local.grad.cpu().numpy(), ei[:, m].numpy(), w[m].numpy()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_partitioned_aggregation_two_ranks_one_gpu():
    from oracle import oracle as O
    N, E, K, world, seed = 5000, 120000, 100, 2, 5
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    with tempfile.TemporaryDirectory() as d:
        procs = [ctx.Process(target=_worker, args=(r, world, os.path.join(d, 'rdzv'), N, E, K, seed, q)) for r in range(world)]
        for p in procs:
            p.start()
        res = [q.get(timeout=300) for _ in range(world)]
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
    res.sort(key=lambda t: t[0])
    g = torch.Generator().manual_seed(seed)
    torch.randint(0, N, (2, E), generator=g)
    torch.rand(E, generator=g)
    x = torch.randn(N, K, generator=g)
    gout = torch.randn(N, K, generator=g)
    ei = np.concatenate([r[3] for r in res], axis=1)                  # the union of the shards is the graph
    w = np.concatenate([r[4] for r in res])
    ref, _ = O.spmm_fwd('sum', ei, w, x.numpy())
    gref = O.spmm_bwd('sum', ei, w, gout.numpy())
    out = np.concatenate([r[1] for r in res])
    gx = np.concatenate([r[2] for r in res])
    assert np.max(np.abs(out - ref)) <= 1e-5 * np.max(np.abs(ref))
    deg = np.bincount(ei[1], minlength=N)
    assert np.array_equal(out[deg <= 1024], ref[deg <= 1024])          # unsplit rows bit-exact across the partition
    assert np.max(np.abs(gx - gref)) <= 1e-5 * np.max(np.abs(gref))
