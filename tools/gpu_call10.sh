#!/bin/bash
# round-2 call 10 (1 GPU): tcgen05 GEMM with K passes (two CTAs per SM, K <= 256): tests, layer bench, full suite, final N=1 line
O=gpurun_out; mkdir -p $O
timeout 240 python -m pytest tests/test_gpu_linear_tc.py -q -m gpu -x 2>&1 | tail -6 > $O/pytest_gpu_r2j_linear_tc.txt
cat $O/pytest_gpu_r2j_linear_tc.txt | tail -3
timeout 400 python tools/bench_layer.py 256 > $O/bench_layer_r2j.json 2> $O/bench_layer_r2j.err
cat $O/bench_layer_r2j.json
timeout 900 python -m pytest tests -q -m gpu --ignore=tests/test_gpu_linear_tc.py 2>&1 | tail -4 > $O/pytest_gpu_r2j.txt
cat $O/pytest_gpu_r2j.txt
timeout 400 ncu --set full --clock-control none -k regex:linear_tf32x3 -c 1 -o $O/prof_linear python tools/bench_layer.py 256 > $O/ncu_linear.log 2>&1
ncu -i $O/prof_linear.ncu-rep --page raw --csv > $O/ncu_r2_linear_tf32x3_v3_raw.csv 2>/dev/null; rm -f $O/prof_linear.ncu-rep
timeout 850 python bench.py --steps 20 --warmup 5 > $O/bench_r2j_rmat.json 2> $O/bench_r2j_rmat.err
tail -c 400 $O/bench_r2j_rmat.json
