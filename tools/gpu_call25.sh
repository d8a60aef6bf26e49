#!/bin/bash
# round-2 call 25 (1 GPU): verification of the final tree: full GPU suite, smoke(), the driver's default bench command,
# the reference arm, sanitizer on the GEMM, SASS-level check that the shipped library is the one measured
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 > $O/pytest_gpu_r2_final.txt; cat $O/pytest_gpu_r2_final.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke_r2_final.txt 2>&1; tail -2 $O/smoke_r2_final.txt
timeout 850 python bench.py > $O/bench_r2_final.json 2> $O/bench_r2_final.err; tail -c 1500 $O/bench_r2_final.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_r2_final_reference.json 2> /dev/null; tail -c 600 $O/bench_r2_final_reference.json
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest tests/test_gpu_linear_tc.py -q -m gpu -x -k "matches_fp64" > $O/sanitizer_r2_linear_memcheck.log 2>&1; tail -2 $O/sanitizer_r2_linear_memcheck.log
timeout 300 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest tests/test_gpu_linear_tc.py -q -m gpu -x -k "matches_fp64 and 1000" > $O/sanitizer_r2_linear_racecheck.log 2>&1; tail -2 $O/sanitizer_r2_linear_racecheck.log
