#!/bin/bash
# round-2 call 1: lock-step lane groups -- tests, narrow-row probe, microbenchmarks, sanitizer logs
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > $O/pytest_gpu_r2a.txt
timeout 300 python tools/probe_narrow.py c4 > $O/probe_narrow_r2a.jsonl 2> $O/probe_narrow_r2a.err
timeout 600 python tools/microbench.py c4 > $O/microbench_r2a_c4.jsonl 2> $O/mb_r2a_c4.err
timeout 600 python tools/microbench.py c3 > $O/microbench_r2a_c3.jsonl 2> $O/mb_r2a_c3.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_r2a_rmat.json 2> $O/bench_r2a_rmat.err
for t in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $t --print-limit 30 python tools/sanitize_cases.py > $O/sanitizer_r2a_$t.log 2>&1
done
tail -3 $O/pytest_gpu_r2a.txt; cat $O/probe_narrow_r2a.jsonl; tail -4 $O/sanitizer_r2a_*.log
