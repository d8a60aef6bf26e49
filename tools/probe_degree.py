"""A/B of the sort-free K = 1 sum (`unsorted_segment_sum(ones[E])`, the kernel under degree()/calc_gcn_norm()) on one B200:
output replicas x edges in flight per lane, on the C3 (uniform + RMAT) and C4 (RMAT) id streams.  One JSON object per line.
Usage: python tools/probe_degree.py"""
import json
import os
import sys

import torch

sys.path.insert(0, '.')
import bench  # noqa: E402
from gammagl_b200 import _lib, mpops  # noqa: E402

dev = torch.device('cuda', 0)


def timeit(fn, reps=10, warm=5):
    import statistics
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


PEAK = bench.measured_peak()[0]

for name, N, E, kind in (('c3_uniform', 2449029, 123718280, 'uniform'), ('c3_rmat', 2449029, 123718280, 'rmat'),
                         ('c4_rmat', 232965, 114615892, 'rmat')):
    ei = bench.gen_graph(torch, N, E, kind, 7, dev)
    dst = ei[1].contiguous()
    ones = torch.ones(E, 1, device=dev)
    ref = torch.bincount(dst, minlength=N).float()
    for rep in (1, 2, 4, 8):
        for unroll in (4, 8):
            os.environ['GGLB200_ONESHOT_REPLICAS'] = str(rep)
            os.environ['GGLB200_ONESHOT_UNROLL'] = str(unroll)
            out = mpops.segment_reduce_oneshot(ones, dst, N, _lib.GGL_SUM)[0]
            ok = bool(torch.equal(out.view(-1), ref))
            ms = timeit(lambda: mpops.segment_reduce_oneshot(ones, dst, N, _lib.GGL_SUM))
            b = E * 12 + N * 4
            print(json.dumps(dict(graph=name, replicas=rep, unroll=unroll, ms=round(ms, 4), exact=ok, alg_GBps=round(b / ms / 1e6, 1),
                                  frac_of_measured_hbm=round(b / ms / 1e6 / PEAK, 3))), flush=True)
