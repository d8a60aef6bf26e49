"""Kernel-level breakdown of one 2-layer GCN training step (100 -> 256 -> 47, C3 graph) through gammagl_b200.layers:
torch.profiler CUDA times per kernel, top 30.  Usage: python tools/profile_model_step.py"""
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, '.')
import bench  # noqa: E402
from gammagl_b200.layers import GCNConv  # noqa: E402

dev = torch.device('cuda', 0)
N, E, F, H = 2449029, 123718280, 100, 256
ei = bench.gen_graph(torch, N, E, 'rmat', 7, dev)
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(N, F, generator=g, device=dev)
y = torch.randint(0, 47, (N,), generator=g, device=dev)
torch.manual_seed(0)
l1 = GCNConv(F, H, aggregate_first=True, fused_activation='relu').to(dev)
l2 = GCNConv(H, 47).to(dev)


def step():
    l1.zero_grad(set_to_none=True); l2.zero_grad(set_to_none=True)
    loss = torch.nn.functional.cross_entropy(l2(l1(x, ei), ei), y)
    loss.backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=30, max_name_column_width=70))
