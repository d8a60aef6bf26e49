"""Diagnostic (not part of the product): run gspmm forward for several feature widths so that ncu can report DRAM /
L2 sectors per gathered row for each copy mechanism (GGLB200_RING=bulk|ldgsts|none)."""
import sys
import torch
sys.path.insert(0, '.')
from gammagl_b200 import mpops

dev = 'cuda:0'
import ctypes, os
torch.zeros(1, device=dev)
gran = int(os.environ.get('L2GRAN', '0'))
rt = ctypes.CDLL('libcudart.so.12')
if gran:
    print('set limit ->', rt.cudaDeviceSetLimit(5, ctypes.c_size_t(gran)))
v = ctypes.c_size_t(0)
rt.cudaDeviceGetLimit(ctypes.byref(v), 5)
print('cudaLimitMaxL2FetchGranularity =', v.value)
N, E = 1_000_000, 20_000_000
g = torch.Generator(device=dev).manual_seed(0)
ei = torch.randint(0, N, (2, E), generator=g, device=dev)
w = torch.rand(E, device=dev)
for K in [int(a) for a in sys.argv[1:]] or [96, 100, 104, 128]:
    x = torch.randn(N, K, device=dev)
    for _ in range(2):
        out = mpops.gspmm(ei, w, x, 'sum')
    torch.cuda.synchronize()
    print('K', K, 'done')
