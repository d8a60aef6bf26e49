#!/bin/bash
# round-2 call 33 (1 GPU): input gradient of the Linear on tcgen05 vs cuBLAS; extended smoke; model step with both
O=gpurun_out; mkdir -p $O
timeout 200 python tools/probe_linear_bwd.py 2>&1 | tee $O/probe_linear_bwd_r2.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/smoke_r2_final.txt
GGLB200_LINEAR_BWD=tc timeout 200 python -m pytest tests/test_gpu_linear_tc.py -q -m gpu 2>&1 | tail -2
