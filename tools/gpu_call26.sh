#!/bin/bash
# round-2 call 26 (1 GPU): SAGEConv(aggr='gcn') with the folded normalisation: tests + layer timing
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_layers.py tests/test_gpu_zz_reference_layers.py -q -m gpu 2>&1 | tail -6 > $O/pytest_gpu_r2_sage_fold.txt; cat $O/pytest_gpu_r2_sage_fold.txt
timeout 400 python tools/bench_sage_gcn.py > $O/bench_sage_gcn_r2.json 2> $O/bench_sage_gcn_r2.err; cat $O/bench_sage_gcn_r2.json; tail -3 $O/bench_sage_gcn_r2.err
