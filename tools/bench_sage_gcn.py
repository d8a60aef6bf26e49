"""SAGEConv(aggr='gcn') forward + backward at the C3 shape (F 100 -> 100): normalisation folded into the aggregation kernel
(mpops.gcn_propagate norm='both_src', bias in the row write) against the literal form (add_self_loops -> calc_gcn_norm ->
weighted gspmm -> bias pass).  One JSON line.  Usage: python tools/bench_sage_gcn.py"""
import json
import statistics
import sys

import torch

sys.path.insert(0, '.')
import bench  # noqa: E402
from gammagl_b200 import layers as L  # noqa: E402

dev = torch.device('cuda', 0)
N, E, F = 2449029, 123718280, 100
ei = bench.gen_graph(torch, N, E, 'rmat', 7, dev)
ei[0, 0] = N - 1            # sage_conv.py:83 sizes the self loops by max(src) + 1
x = torch.randn(N, F, device=dev, requires_grad=True)
go = torch.randn(N, F, device=dev)


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


out = {}
for name, fold in (('folded', True), ('literal', False)):
    torch.manual_seed(0)
    layer = L.SAGEConv(F, F, activation=torch.relu, aggr='gcn', fold_norm=fold).to(dev)

    def step():
        x.grad = None
        layer.zero_grad(set_to_none=True)
        layer(x, ei).backward(go)

    out[name + '_fwd_bwd_ms'] = round(timeit(step), 3)
    with torch.no_grad():
        out[name + '_fwd_ms'] = round(timeit(lambda: layer(x, ei)), 3)
out['shape'] = {'N': N, 'E': E, 'F': F}
print(json.dumps(out))
