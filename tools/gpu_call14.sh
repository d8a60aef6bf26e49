#!/bin/bash
# round-2 call 14 (1 GPU): tcgen05 GEMM with the bulk-copy-fed W ring: tests (own process, short fuse), layer + model bench, ncu
O=gpurun_out; mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_linear_tc.py -q -m gpu 2>&1 | tail -12 > $O/pytest_gpu_r2n_linear_tc.txt
cat $O/pytest_gpu_r2n_linear_tc.txt | tail -4
timeout 500 python tools/bench_layer.py 256 > $O/bench_layer_r2n.json 2> $O/bench_layer_r2n.err
cat $O/bench_layer_r2n.json; tail -3 $O/bench_layer_r2n.err
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_linear_tc.py -q -m gpu -x -k "matches_fp64" > $O/sanitizer_r2n_linear_memcheck.log 2>&1
tail -3 $O/sanitizer_r2n_linear_memcheck.log
timeout 400 ncu --set full --clock-control none -k regex:linear_tf32x3 -c 1 -o $O/prof_linear python tools/bench_layer.py 256 > $O/ncu_linear.log 2>&1
ncu -i $O/prof_linear.ncu-rep --page raw --csv > $O/ncu_r2_linear_tf32x3_v4_raw.csv 2>/dev/null; rm -f $O/prof_linear.ncu-rep
