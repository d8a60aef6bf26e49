#!/bin/bash
# round-2 call 31 (1 GPU): warp-specialised GEMM with the bulk-copied A tile: tests, staged vs per-thread loads, layer bench, ncu
O=gpurun_out; mkdir -p $O
export GGLB200_LINEAR=ws
timeout 120 python -m pytest tests/test_gpu_linear_tc.py -q -m gpu 2>&1 | tail -3
echo staged; timeout 100 python tools/probe_linear.py 2>&1 | tee $O/probe_linear_r2_ws_staged.json
echo per-thread loads; GGLB200_LINEAR_STAGE=0 timeout 100 python tools/probe_linear.py 2>&1 | tee $O/probe_linear_r2_ws_ldg.json
timeout 300 python tools/bench_layer.py 256 > $O/bench_layer_r2_ws3.json 2>/dev/null; cut -c1-1000 $O/bench_layer_r2_ws3.json
timeout 400 ncu --set full --import-source on --clock-control none -k regex:linear_ws -c 1 -o $O/prof_linear python tools/bench_layer.py 256 > $O/ncu_linear.log 2>&1
ncu -i $O/prof_linear.ncu-rep --page source --csv > $O/ncu_r2_linear_ws_source.csv 2>/dev/null
ncu -i $O/prof_linear.ncu-rep --page raw --csv > $O/ncu_r2_linear_ws_raw.csv 2>/dev/null; rm -f $O/prof_linear.ncu-rep
