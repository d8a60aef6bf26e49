"""Small invocations of every ring / sub-warp kernel configuration for compute-sanitizer (memcheck, racecheck, synccheck).
Usage (on a GPU box): compute-sanitizer --tool racecheck python tools/sanitize_cases.py"""
import sys

import torch

sys.path.insert(0, '.')
from gammagl_b200 import mpops  # noqa: E402

dev = torch.device('cuda', 0)
g = torch.Generator().manual_seed(0)
N, E = 700, 9000
ei = torch.randint(0, N, (2, E), generator=g)
ei[1, :2500] = 5          # one split row (> 1024 edges)
ei = ei.to(dev)
w = torch.rand(E, generator=g).to(dev)
for K in (4, 8, 16, 32, 64, 100, 128, 47):
    x = torch.randn(N, K, generator=g).to(dev).requires_grad_(True)
    for red in ('sum', 'mean', 'max'):
        mpops.gspmm(ei, w, x, red).sum().backward()
    m = torch.randn(E, K, generator=g).to(dev).requires_grad_(True)
    for f in (mpops.unsorted_segment_sum, mpops.unsorted_segment_mean, mpops.unsorted_segment_max):
        f(m, ei[1], N).sum().backward()
    if K in (64, 100):
        mpops.gspmm(ei, w, x.detach().bfloat16(), 'sum')
for H in (1, 2, 4, 8):
    lg = torch.randn(E, H, generator=g).to(dev).requires_grad_(True)
    mpops.segment_softmax(lg, ei[1], N).square().sum().backward()
for H, C in ((8, 8), (4, 16), (1, 64), (2, 4), (4, 25), (3, 5)):
    x3 = torch.randn(N, H, C, generator=g).to(dev).requires_grad_(True)
    wh = torch.rand(E, H, generator=g).to(dev).requires_grad_(True)
    mpops.bspmm(ei, wh, x3).sum().backward()
    if mpops.gat_fusable(H, C):
        ss = torch.randn(N, H, generator=g).to(dev).requires_grad_(True)
        sd = torch.randn(N, H, generator=g).to(dev).requires_grad_(True)
        mpops.gat_aggregate(ei, x3, ss, sd, 0.2).square().sum().backward()
# round-2 kernels: spanning lanes (C % 4 != 0), edge-id max, bipartite rows, sort-free segment reduce, tcgen05 GEMM, sampler, merged GAT backward
from gammagl_b200 import sampler  # noqa: E402
for H, C in ((4, 25), (3, 12)):
    x3 = torch.randn(N, H, C, generator=g).to(dev).requires_grad_(True)
    wh = torch.rand(E, H, generator=g).to(dev).requires_grad_(True)
    mpops.bspmm(ei, wh, x3).sum().backward()
xb = torch.randn(N, 20, generator=g).to(dev).requires_grad_(True)
mpops.gspmm_max_edge(ei, w, xb).sum().backward()
eib = torch.stack([ei[0], ei[1] % 300])
mpops.gspmm(eib, w, xb, 'mean', num_dst=300).sum().backward()
for K in (1, 2, 4, 8, 33):
    m = torch.randn(E, K, generator=g).to(dev).requires_grad_(True)
    for code in (0, 1, 2):
        mpops.segment_reduce_oneshot(m.detach(), ei[1], N, code)
a = torch.randn(300, 100, generator=g).to(dev)
wl = torch.randn(100, 192, generator=g).to(dev)
mpops.linear_tc(a, wl, torch.randn(192, generator=g).to(dev), True)
ns = sampler.NeighborSampler(ei, [5, 3], num_nodes=N, seed=1)
ns.sample(torch.arange(50))
sampler.neighbor_sample(ns.rowptr, ns.col, torch.arange(20, device=dev), [4, -1], directed=False, seed=2)
torch.cuda.synchronize()
print('sanitize cases done')
