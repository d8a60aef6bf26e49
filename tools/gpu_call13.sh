#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_linear_tc.py -q -m gpu -k matches_fp64 -rA 2>&1 | grep -E "PASSED|FAILED|passed|failed" | head -40 > $O/pytest_gpu_r2m_linear_tc.txt
timeout 100 python - > $O/linear_debug.txt 2>&1 <<'PY'
import torch, sys
sys.path.insert(0, '.')
from gammagl_b200 import mpops
dev = 'cuda:0'
for (M, K, N) in [(128, 8, 64), (128, 56, 64), (128, 56, 128), (128, 56, 192), (128, 56, 256), (128, 100, 64), (128, 100, 128), (256, 8, 128)]:
    a = torch.randn(M, K, device=dev); w = torch.randn(K, N, device=dev)
    out = mpops.linear_tc(a, w)
    ref = a.double() @ w.double()
    err = ((out.double() - ref).abs().amax(0) / ref.abs().max()).cpu()
    print((M, K, N), 'max rel err per 64-col chunk:', [round(float(err[c:c + 64].max()), 8) for c in range(0, N, 64)])
PY
cat $O/pytest_gpu_r2m_linear_tc.txt; cat $O/linear_debug.txt
