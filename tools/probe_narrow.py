"""Probe: segment-sum forward over narrow rows (K = 4..32 fp32) -- rows/s of the bulk-copy ring vs the register-staged
kernel (run once with GGLB200_RING=none).  Usage: python tools/probe_narrow.py [c4|c3]"""
import json
import os
import sys

import torch

sys.path.insert(0, '.')
import bench  # noqa: E402
from gammagl_b200 import _lib, mpops  # noqa: E402
from gammagl_b200.csr import segment_csr_of  # noqa: E402
from tools.microbench import timeit  # noqa: E402

dev = torch.device('cuda', 0)
cfg = sys.argv[1] if len(sys.argv) > 1 else 'c4'
N, E = (232965, 114615892) if cfg == 'c4' else (2449029, 123718280)
ei = bench.gen_graph(torch, N, E, 'rmat', 7, dev)
dst = ei[1].contiguous()
csr = segment_csr_of(dst, N)
sorted_dst, _ = torch.sort(dst)
csr_sorted = segment_csr_of(sorted_dst, N)
for K in (4, 8, 16, 32):
    x = torch.randn(E, K, device=dev)
    for name, c in (('random order', csr), ('dst-sorted edges', csr_sorted)):
        ms = timeit(lambda: mpops.spmm_csr(c, c.perm, x, _lib.GGL_SUM))
        print(json.dumps(dict(K=K, order=name, ring=os.environ.get('GGLB200_RING', 'bulk'), ms=round(ms, 3), rows_per_s=E / ms * 1e3)), flush=True)
    del x
