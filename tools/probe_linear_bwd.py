"""Input gradient of the Linear behind the aggregation, g [M, 256] @ W^T [256, 100] at the C3 row count: torch.matmul (cuBLAS
fp32) against the tcgen05 3xTF32 kernel on the transposed weight.  One JSON line.  Usage: python tools/probe_linear_bwd.py"""
import json
import statistics
import sys

import torch

sys.path.insert(0, '.')
from gammagl_b200 import mpops  # noqa: E402

dev = torch.device('cuda', 0)
M, K, N = 2449029, 100, 256
g = torch.randn(M, N, device=dev); w = torch.randn(K, N, device=dev)
wt = w.t().contiguous()


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


ref = g[:4096].double() @ w.t().double()
a = mpops._linear_tf32x3(g, wt, None, False)
b = g @ w.t()
out = {'M': M, 'K_in': N, 'N_out': K,
       'tcgen05_ms': round(timeit(lambda: mpops._linear_tf32x3(g, wt, None, False)), 4),
       'cublas_fp32_ms': round(timeit(lambda: g @ w.t()), 4),
       'rel_err_tcgen05': float((a[:4096].double() - ref).abs().max() / ref.abs().max()),
       'rel_err_cublas': float((b[:4096].double() - ref).abs().max() / ref.abs().max())}
print(json.dumps(out))
