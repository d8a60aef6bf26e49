#!/bin/bash
# round-2 call 4 (1 GPU): full GPU suite with the new kernels (spanning lanes, edge-id max, bipartite, one-shot segment reduce, sampler,
# GATv2 / AGNN), the tcgen05 GEMM in its own short-fused process, the lean narrow-row kernel A/B, probes, microbenchmarks, layer bench
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x --ignore=tests/test_gpu_linear_tc.py 2>&1 | tail -25 > $O/pytest_gpu_r2d.txt
timeout 240 python -m pytest tests/test_gpu_linear_tc.py -q -m gpu -x 2>&1 | tail -25 > $O/pytest_gpu_r2d_linear_tc.txt
GGLB200_NARROW=ldg timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_layers.py tests/test_gpu_oneshot.py -q -m gpu -x 2>&1 | tail -8 > $O/pytest_gpu_r2d_narrow_ldg.txt
timeout 300 build/probe_atomics > $O/probe_atomics_r2.jsonl 2>&1
GGLB200_NARROW=ring timeout 300 python tools/probe_narrow.py c4 > $O/probe_narrow_r2d_ring.jsonl 2> $O/probe_narrow_r2d_ring.err
GGLB200_NARROW=ldg timeout 300 python tools/probe_narrow.py c4 > $O/probe_narrow_r2d_ldg.jsonl 2> $O/probe_narrow_r2d_ldg.err
timeout 600 python tools/microbench.py c3 > $O/microbench_r2d_c3.jsonl 2> $O/mb_r2d_c3.err
timeout 600 python tools/microbench.py c4 > $O/microbench_r2d_c4.jsonl 2> $O/mb_r2d_c4.err
GGLB200_NARROW=ldg timeout 600 python tools/microbench.py c4 > $O/microbench_r2d_c4_narrow_ldg.jsonl 2> $O/mb_r2d_c4_ldg.err
timeout 400 python tools/bench_layer.py 256 > $O/bench_layer_r2d.json 2> $O/bench_layer_r2d.err
tail -12 $O/pytest_gpu_r2d.txt; tail -12 $O/pytest_gpu_r2d_linear_tc.txt; tail -4 $O/pytest_gpu_r2d_narrow_ldg.txt; cat $O/probe_atomics_r2.jsonl
cat $O/probe_narrow_r2d_ring.jsonl $O/probe_narrow_r2d_ldg.jsonl; cat $O/bench_layer_r2d.json; tail -2 $O/bench_layer_r2d.err
