#!/bin/bash
# round-2 call 16 (1 GPU): tcgen05 GEMM with the new epilogue: tests, layer bench, ncu raw + source
O=gpurun_out; mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_linear_tc.py -q -m gpu 2>&1 | tail -12 > $O/pytest_gpu_r2o_linear_tc.txt
cat $O/pytest_gpu_r2o_linear_tc.txt | tail -4
timeout 500 python tools/bench_layer.py 256 > $O/bench_layer_r2o.json 2> $O/bench_layer_r2o.err
cat $O/bench_layer_r2o.json; tail -3 $O/bench_layer_r2o.err
timeout 400 ncu --set full --import-source on --clock-control none -k regex:linear_tf32x3 -c 1 -o $O/prof_linear python tools/bench_layer.py 256 > $O/ncu_linear.log 2>&1
ncu -i $O/prof_linear.ncu-rep --page source --csv > $O/ncu_r2_linear_tf32x3_v5_source.csv 2>/dev/null
ncu -i $O/prof_linear.ncu-rep --page raw --csv > $O/ncu_r2_linear_tf32x3_v5_raw.csv 2>/dev/null; rm -f $O/prof_linear.ncu-rep
