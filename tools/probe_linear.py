"""Where the time of the tcgen05 Linear goes: the kernel timed with parts switched off (GGLB200_LINEAR_DBG bits: 1 no W copies,
2 no output stores, 4 no A loads, 8 consecutive MMAs on different accumulators; results are wrong by construction, only the time is read).  One JSON line.
Usage: python tools/probe_linear.py"""
import json
import os
import statistics
import sys

import torch

sys.path.insert(0, '.')
from gammagl_b200 import mpops  # noqa: E402

dev = torch.device('cuda', 0)
M, K, N = 2449029, 100, 256
a = torch.randn(M, K, device=dev); w = torch.randn(K, N, device=dev); b = torch.randn(N, device=dev)


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return statistics.median(ts)


out = {}
for dbg in (0,):
    os.environ['GGLB200_LINEAR_DBG'] = str(dbg)
    out['dbg_%d_ms' % dbg] = round(timeit(lambda: mpops.linear_tc(a, w, b, relu=True)), 4)
os.environ['GGLB200_LINEAR_DBG'] = '0'
out['note'] = 'bit 1: no W copies, bit 2: no output stores, bit 4: no A loads; M=%d K=%d N=%d' % (M, K, N)
print(json.dumps(out))
