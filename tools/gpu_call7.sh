#!/bin/bash
# round-2 call 7 (1 GPU): verification of the sort-free softmax and the batched-load tcgen05 GEMM, their numbers, final N=1 bench line
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x --ignore=tests/test_gpu_linear_tc.py 2>&1 | tail -6 > $O/pytest_gpu_r2g.txt
timeout 240 python -m pytest tests/test_gpu_linear_tc.py -q -m gpu -x 2>&1 | tail -6 > $O/pytest_gpu_r2g_linear_tc.txt
cat $O/pytest_gpu_r2g.txt $O/pytest_gpu_r2g_linear_tc.txt | tail -8
timeout 400 python tools/bench_layer.py 256 > $O/bench_layer_r2g.json 2> $O/bench_layer_r2g.err
timeout 600 python tools/microbench.py c4 > $O/microbench_r2g_c4.jsonl 2> $O/mb_r2g_c4.err
GGLB200_SOFTMAX_PATH=csr timeout 600 python tools/microbench.py c4 2>/dev/null | grep softmax > $O/microbench_r2g_c4_softmax_csr.jsonl
timeout 600 python tools/microbench.py c3 > $O/microbench_r2g_c3.jsonl 2> $O/mb_r2g_c3.err
timeout 850 python bench.py --steps 20 --warmup 5 > $O/bench_r2g_rmat.json 2> $O/bench_r2g_rmat.err
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_cases.py > $O/sanitizer_r2g_memcheck.log 2>&1
cat $O/bench_layer_r2g.json; grep softmax $O/microbench_r2g_c4.jsonl $O/microbench_r2g_c4_softmax_csr.jsonl | cut -c1-220; tail -c 600 $O/bench_r2g_rmat.json; tail -2 $O/sanitizer_r2g_memcheck.log
