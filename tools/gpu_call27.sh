#!/bin/bash
# round-2 call 27 (1 GPU): derived-edge-list cache + content match of rebuilt edge lists: tests, SAGE-gcn layer timing, full suite
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 > $O/pytest_gpu_r2_final2.txt; cat $O/pytest_gpu_r2_final2.txt
timeout 400 python tools/bench_sage_gcn.py > $O/bench_sage_gcn_r2.json 2> $O/bench_sage_gcn_r2.err; cat $O/bench_sage_gcn_r2.json; tail -3 $O/bench_sage_gcn_r2.err
GGLB200_CONTENT_CACHE=0 timeout 400 python tools/bench_sage_gcn.py > $O/bench_sage_gcn_r2_no_content_cache.json 2>/dev/null; cat $O/bench_sage_gcn_r2_no_content_cache.json
# warp-specialised persistent GEMM (opt-in until verified): tests with a short fuse, timing
GGLB200_LINEAR=ws timeout 150 python -m pytest tests/test_gpu_linear_tc.py -q -m gpu 2>&1 | tail -8 > $O/pytest_gpu_r2_linear_ws.txt; cat $O/pytest_gpu_r2_linear_ws.txt
GGLB200_LINEAR=ws timeout 200 python tools/bench_layer.py 256 > $O/bench_layer_r2_ws.json 2> /dev/null; cut -c1-900 $O/bench_layer_r2_ws.json
