#!/bin/bash
# round-2 call 3 (1 GPU): TMA row-gather probe (bulk vs gather4 vs ldg), plugin-on-CUDA tests, N=1 bench with the configs section
O=gpurun_out; mkdir -p $O
timeout 300 build/probe_tma 02 > $O/probe_tma_r2.jsonl 2> $O/probe_tma_r2.err
timeout 200 build/probe_tma 1 1 > $O/probe_tma_r2_gather4_box1.jsonl 2>&1
timeout 200 build/probe_tma 1 4 > $O/probe_tma_r2_gather4_box4.jsonl 2>&1
timeout 600 python -m pytest tests/test_gpu_plugin.py tests/test_gpu_robustness.py -q -m gpu 2>&1 | tail -25 > $O/pytest_gpu_r2c.txt
timeout 850 python bench.py --steps 20 --warmup 5 > $O/bench_r2c_rmat.json 2> $O/bench_r2c_rmat.err
cat $O/probe_tma_r2.jsonl $O/probe_tma_r2_gather4_box1.jsonl $O/probe_tma_r2_gather4_box4.jsonl; tail -8 $O/pytest_gpu_r2c.txt; tail -c 1500 $O/bench_r2c_rmat.json; tail -3 $O/bench_r2c_rmat.err
