#!/bin/bash
# round-2 call 28 (1 GPU): warp-specialised GEMM: stand-alone time, ncu raw + source
O=gpurun_out; mkdir -p $O
export GGLB200_LINEAR=ws
timeout 200 python tools/probe_linear.py > $O/probe_linear_r2_ws.json 2>&1; cat $O/probe_linear_r2_ws.json
timeout 400 ncu --set full --import-source on --clock-control none -k regex:linear_ws -c 1 -o $O/prof_linear python tools/bench_layer.py 256 > $O/ncu_linear.log 2>&1
ncu -i $O/prof_linear.ncu-rep --page source --csv > $O/ncu_r2_linear_ws_source.csv 2>/dev/null
ncu -i $O/prof_linear.ncu-rep --page raw --csv > $O/ncu_r2_linear_ws_raw.csv 2>/dev/null; rm -f $O/prof_linear.ncu-rep
for r in 2 3 4; do echo ring $r; GGLB200_LINEAR_RING=$r timeout 100 python tools/probe_linear.py 2>&1 | tee $O/probe_linear_r2_ws_ring$r.json; done
GGLB200_LINEAR_RING=4 timeout 100 python -m pytest tests/test_gpu_linear_tc.py -q -m gpu 2>&1 | tail -2
GGLB200_LINEAR_RING=2 timeout 100 python -m pytest tests/test_gpu_linear_tc.py -q -m gpu 2>&1 | tail -2
