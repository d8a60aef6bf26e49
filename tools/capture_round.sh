#!/bin/bash
# One-shot evidence capture on a B200 box (run under gpurun from the repo root): tests, bench lines, microbenchmarks,
# ncu launch list and ncu --set full exports, sanitizer logs.  Everything lands in gpurun_out/ (copied into profiles/ afterwards).
R=${1:-r2}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 > $O/pytest_gpu_$R.txt
timeout 850 python bench.py --steps 20 --warmup 5 > $O/bench_${R}_rmat.json 2> $O/bench_${R}_rmat.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_${R}_reference.json 2> $O/bench_${R}_reference.err
timeout 600 python tools/microbench.py c3 > $O/microbench_${R}_c3_rmat.jsonl 2> $O/micro_c3.err
timeout 600 python tools/microbench.py c4 > $O/microbench_${R}_c4_rmat.jsonl 2> $O/micro_c4.err
GGLB200_GAT_BWD=two_pass timeout 300 python tools/microbench.py c4 2> /dev/null | grep gat_aggregate > $O/microbench_${R}_c4_gat_two_pass.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"spmm|ggl|sddmm|softmax|oneshot|linear|gat_" -c 60 --csv --log-file $O/launches_${R}_final.csv \
    python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --no-configs > $O/ncu_bench.log 2>&1
for g in rmat uniform; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmm_ring -s 4 -c 1 -o $O/prof_ring_$g \
      python bench.py --graph $g --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-configs > $O/ncu_full_$g.log 2>&1
  ncu -i $O/prof_ring_$g.ncu-rep --page raw --csv > $O/ncu_${R}_spmm_ring_${g}_raw.csv 2>/dev/null
  ncu -i $O/prof_ring_$g.ncu-rep --page details --csv > $O/ncu_${R}_spmm_ring_${g}_details.csv 2>/dev/null
  # stamp: bench.py only trusts a capture taken from the kernel sources it was built from (roofline.traffic)
  python -c "import json,hashlib;from gammagl_b200 import build as b;print(json.dumps({'kernel_digest':b.kernel_digest(),'source_digest':b._digest(),'lib_sha256':hashlib.sha256(open(b.LIB,'rb').read()).hexdigest(),'round':'$R','graph':'$g'}))" > $O/ncu_${R}_spmm_ring_${g}_raw.csv.stamp.json
  rm -f $O/prof_ring_$g.ncu-rep
done
# tcgen05 GEMM and the one-pass GAT backward: one full capture each (tensor-pipe / L2 reduction evidence)
timeout 400 ncu --set full --clock-control none -k regex:linear_tf32x3 -c 1 -o $O/prof_linear python tools/bench_layer.py 256 > $O/ncu_linear.log 2>&1
ncu -i $O/prof_linear.ncu-rep --page raw --csv > $O/ncu_${R}_linear_tf32x3_raw.csv 2>/dev/null; rm -f $O/prof_linear.ncu-rep
timeout 400 ncu --set full --clock-control none -k regex:gat_bwd_src -c 1 -o $O/prof_gatbwd python tools/microbench.py c4 > $O/ncu_gatbwd.log 2>&1
ncu -i $O/prof_gatbwd.ncu-rep --page raw --csv > $O/ncu_${R}_gat_bwd_merged_raw.csv 2>/dev/null; rm -f $O/prof_gatbwd.ncu-rep
for t in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $t --print-limit 30 python tools/sanitize_cases.py > $O/sanitizer_${R}_$t.log 2>&1
done
timeout 300 python bench.py --steps 20 --warmup 5 --no-configs > $O/bench_${R}_rmat_after_capture.json 2> /dev/null   # traffic now stamped
timeout 200 python examples/sage_minibatch.py --steps 20 > $O/example_sage_minibatch_$R.txt 2>&1
ls -la $O | tail -40
