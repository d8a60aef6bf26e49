#!/bin/bash
# round-2 call 21 (1 GPU): tcgen05 GEMM with the warp-uniform MMA role (elected lane, descriptors in uniform registers)
O=gpurun_out; mkdir -p $O
timeout 120 python -m pytest tests/test_gpu_linear_tc.py -q -m gpu 2>&1 | tail -12 > $O/pytest_gpu_r2s_linear_tc.txt
cat $O/pytest_gpu_r2s_linear_tc.txt | tail -4
timeout 200 python tools/probe_linear.py > $O/probe_linear_r2f.json 2>&1; cat $O/probe_linear_r2c.json
timeout 500 python tools/bench_layer.py 256 > $O/bench_layer_r2s.json 2> $O/bench_layer_r2s.err
cat $O/bench_layer_r2s.json; tail -3 $O/bench_layer_r2s.err
timeout 400 ncu --set full --import-source on --clock-control none -k regex:linear_tf32x3 -c 1 -o $O/prof_linear python tools/bench_layer.py 256 > $O/ncu_linear.log 2>&1
ncu -i $O/prof_linear.ncu-rep --page source --csv > $O/ncu_r2_linear_tf32x3_v9_source.csv 2>/dev/null
ncu -i $O/prof_linear.ncu-rep --page raw --csv > $O/ncu_r2_linear_tf32x3_v9_raw.csv 2>/dev/null; rm -f $O/prof_linear.ncu-rep
