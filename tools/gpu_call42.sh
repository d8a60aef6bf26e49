#!/bin/bash
# round-2 call 42 (1 GPU): 32-column plan of the warp-specialised GEMM (A ring 64 KB, 5-slot W ring) against the shipped plan
O=gpurun_out; mkdir -p $O
GGLB200_LINEAR_PLAN=deep timeout 150 python -m pytest tests/test_gpu_linear_tc.py -q -m gpu 2>&1 | tail -3 | tee $O/pytest_gpu_r2_linear_deep.txt
echo default; timeout 100 python tools/probe_linear.py 2>&1 | tee $O/probe_linear_r2_ws_default.json
for r in 3 4 5 6; do echo deep ring $r; GGLB200_LINEAR_PLAN=deep GGLB200_LINEAR_RING=$r timeout 100 python tools/probe_linear.py 2>&1 | tee $O/probe_linear_r2_ws_deep_ring$r.json; done
