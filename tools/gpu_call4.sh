#!/bin/bash
# round-2 call 4 (1 GPU): full GPU suite with the new kernels (spanning lanes, edge-id max, bipartite), atomics probe, C3 microbench
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -25 > $O/pytest_gpu_r2d.txt
timeout 300 build/probe_atomics > $O/probe_atomics_r2.jsonl 2>&1
timeout 600 python tools/microbench.py c3 > $O/microbench_r2d_c3.jsonl 2> $O/mb_r2d_c3.err
tail -12 $O/pytest_gpu_r2d.txt; cat $O/probe_atomics_r2.jsonl; grep -i "bspmm\|SDDMM" $O/microbench_r2d_c3.jsonl; tail -3 $O/mb_r2d_c3.err
