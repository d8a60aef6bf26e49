#!/bin/bash
# round-2 call 29 (1 GPU): warp-specialised GEMM: W ring depth A/B
O=gpurun_out; mkdir -p $O
export GGLB200_LINEAR=ws
for r in 2 3 4; do echo ring $r; GGLB200_LINEAR_RING=$r timeout 100 python tools/probe_linear.py 2>&1 | tee $O/probe_linear_r2_ws_ring$r.json; done
GGLB200_LINEAR_RING=4 timeout 100 python -m pytest tests/test_gpu_linear_tc.py -q -m gpu 2>&1 | tail -2
GGLB200_LINEAR_RING=2 timeout 100 python -m pytest tests/test_gpu_linear_tc.py -q -m gpu 2>&1 | tail -2
