#!/bin/bash
# round-2 call 15 (1 GPU): source-level ncu of the tcgen05 GEMM; K = 1 sort-free sum A/B (replicas x unroll)
O=gpurun_out; mkdir -p $O
timeout 300 python tools/probe_degree.py > $O/probe_degree_r2.jsonl 2> $O/probe_degree_r2.err
cat $O/probe_degree_r2.jsonl; tail -3 $O/probe_degree_r2.err
timeout 400 ncu --set full --import-source on --clock-control none -k regex:linear_tf32x3 -c 1 -o $O/prof_linear python tools/bench_layer.py 256 > $O/ncu_linear.log 2>&1
ncu -i $O/prof_linear.ncu-rep --page source --csv > $O/ncu_r2_linear_tf32x3_v4_source.csv 2>/dev/null
ncu -i $O/prof_linear.ncu-rep --page raw --csv > $O/ncu_r2_linear_tf32x3_v4_raw.csv 2>/dev/null; rm -f $O/prof_linear.ncu-rep
ls -la $O | tail -5
