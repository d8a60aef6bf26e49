"""One GCN layer (Linear + normalised propagate + bias + ReLU) at the C3 shape, forward, four ways -- and the effect of the
32-byte L2 fetch granularity (gglb200_set_l2_fetch_granularity, set by tune_device for the row gathers) on the library GEMM:

  transform-first   x @ W (cuBLAS fp32, N x 100 -> 256)  then propagate K = 256 with the bias / ReLU epilogue
  aggregate-first   propagate K = 100, then (a) x @ W + b, relu as torch ops (cuBLAS fp32)   (b) mpops.linear_tc (tcgen05 3xTF32)

Usage: python tools/bench_layer.py [F_out]"""
import json
import sys

import torch

sys.path.insert(0, '.')
import bench  # noqa: E402
from gammagl_b200 import _lib, mpops  # noqa: E402
from tools.microbench import timeit  # noqa: E402

dev = torch.device('cuda', 0)
N, E, F = 2449029, 123718280, 100
H = int(sys.argv[1]) if len(sys.argv) > 1 else 256
lib = _lib.load()
ei = bench.gen_graph(torch, N, E, 'rmat', 1000, dev)
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(N, F, generator=g, device=dev)
W = torch.randn(F, H, generator=g, device=dev) * 0.1
b = torch.randn(H, generator=g, device=dev)
mpops.gcn_propagate(ei, x, 'both')           # builds + caches the CSRs, tunes the device (32-byte fetch granularity)
out = {}
for gran in (32, 64, 128):
    _lib.check(lib.gglb200_set_l2_fetch_granularity(gran, None))
    out['cublas_x@W_ms_l2fetch%d' % gran] = timeit(lambda: x @ W)
    out['propagate_K100_ms_l2fetch%d' % gran] = timeit(lambda: mpops.gcn_propagate(ei, x, 'both'))
_lib.check(lib.gglb200_set_l2_fetch_granularity(32, None))
out['linear_tc_ms'] = timeit(lambda: mpops.linear_tc(x, W, b, True))
out['cublas_linear_bias_relu_ms'] = timeit(lambda: torch.relu(x @ W + b))
xw = x @ W
out['transform_first_ms'] = timeit(lambda: mpops.gcn_propagate(ei, x @ W, 'both', bias=b, relu=True))
out['aggregate_first_cublas_ms'] = timeit(lambda: torch.relu(mpops.gcn_propagate(ei, x, 'both') @ W + b))
out['aggregate_first_linear_tc_ms'] = timeit(lambda: mpops.linear_tc(mpops.gcn_propagate(ei, x, 'both'), W, b, True))
ref = torch.relu(mpops.gcn_propagate(ei, x, 'both').double() @ W.double() + b.double())
got = mpops.linear_tc(mpops.gcn_propagate(ei, x, 'both'), W, b, True)
lib32 = torch.relu(mpops.gcn_propagate(ei, x, 'both') @ W + b)
out['rel_err_linear_tc_vs_fp64'] = float((got.double() - ref).abs().max() / ref.abs().max())
out['rel_err_cublas_fp32_vs_fp64'] = float((lib32.double() - ref).abs().max() / ref.abs().max())
# ---- a whole model step: 2-layer GCN (100 -> 256 -> 47 classes, ReLU between), forward + backward, layers of gammagl_b200.layers ----
from gammagl_b200.layers import GCNConv  # noqa: E402
del ref, got, lib32, xw
torch.cuda.empty_cache()
y = torch.randint(0, 47, (N,), generator=g, device=dev)


def build_model(aggregate_first):
    torch.manual_seed(0)
    l1 = GCNConv(F, H, aggregate_first=aggregate_first, fused_activation='relu').to(dev)
    l2 = GCNConv(H, 47).to(dev)
    return l1, l2


def model_step(layers):
    l1, l2 = layers
    for l in layers:
        l.zero_grad(set_to_none=True)
    logits = l2(l1(x, ei), ei)
    loss = torch.nn.functional.cross_entropy(logits, y)
    loss.backward()
    return loss


ma, mb = build_model(False), build_model(True)
la, lb = float(model_step(ma)), float(model_step(mb))
out['gcn2_transform_first_fwd_bwd_ms'] = timeit(lambda: model_step(ma), reps=5, warm=2)
out['gcn2_aggregate_first_linear_tc_fwd_bwd_ms'] = timeit(lambda: model_step(mb), reps=5, warm=2)
out['gcn2_loss_transform_first'], out['gcn2_loss_aggregate_first'] = la, lb
out['gcn2_note'] = 'GCN 100 -> 256 -> 47 on the C3 graph, cross-entropy, forward + backward (no optimizer); layer 1 aggregate-first runs the tcgen05 Linear'
out['shape'] = {'N': N, 'E': E, 'F_in': F, 'F_out': H}
print(json.dumps(out))
