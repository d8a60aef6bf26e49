#!/bin/bash
# round-2 call 41 (1 GPU): last verification of the final tree
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 > $O/pytest_gpu_r2_final.txt; cat $O/pytest_gpu_r2_final.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/smoke_r2_final.txt
timeout 850 python bench.py > $O/bench_r2_final.json 2> $O/bench_r2_final.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_final.json').read().strip().split('\n')[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['roofline']['uniform_graph']['frac'], d['gpu_launches'], d['clocks'], list(d['configs'].keys()))
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_r2_final_reference.json 2>/dev/null; tail -c 300 $O/bench_r2_final_reference.json
