#!/bin/bash
# round-2 call 32 (1 GPU): final tree with the warp-specialised GEMM as default: full suite, smoke, bench, sanitizers on the GEMM
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 > $O/pytest_gpu_r2_final.txt; cat $O/pytest_gpu_r2_final.txt
GGLB200_LINEAR=tile timeout 120 python -m pytest tests/test_gpu_linear_tc.py -q -m gpu 2>&1 | tail -2 | tee $O/pytest_gpu_r2_linear_tile_form.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke_r2_final.txt 2>&1; tail -2 $O/smoke_r2_final.txt
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_linear_tc.py -q -m gpu -x -k "matches_fp64" > $O/sanitizer_r2_linear_memcheck.log 2>&1; tail -2 $O/sanitizer_r2_linear_memcheck.log
timeout 400 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_linear_tc.py -q -m gpu -x -k "matches_fp64 and (1000 or 300)" > $O/sanitizer_r2_linear_racecheck.log 2>&1; tail -2 $O/sanitizer_r2_linear_racecheck.log
timeout 850 python bench.py > $O/bench_r2_final.json 2> $O/bench_r2_final.err; tail -c 600 $O/bench_r2_final.json
