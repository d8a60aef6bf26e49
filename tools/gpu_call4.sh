#!/bin/bash
# round-2 call 4 (1 GPU): full GPU suite with the new kernels (spanning lanes, edge-id max, bipartite, one-shot segment reduce),
# the tcgen05 GEMM in its own short-fused process, atomics probe, C3 microbench
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x --ignore=tests/test_gpu_linear_tc.py 2>&1 | tail -25 > $O/pytest_gpu_r2d.txt
timeout 240 python -m pytest tests/test_gpu_linear_tc.py -q -m gpu -x 2>&1 | tail -25 > $O/pytest_gpu_r2d_linear_tc.txt
timeout 300 build/probe_atomics > $O/probe_atomics_r2.jsonl 2>&1
timeout 600 python tools/microbench.py c3 > $O/microbench_r2d_c3.jsonl 2> $O/mb_r2d_c3.err
tail -12 $O/pytest_gpu_r2d.txt; tail -12 $O/pytest_gpu_r2d_linear_tc.txt; cat $O/probe_atomics_r2.jsonl; grep -i "bspmm\|SDDMM\|degree" $O/microbench_r2d_c3.jsonl; tail -3 $O/mb_r2d_c3.err
