#!/bin/bash
# One-shot evidence capture on a B200 box (run under gpurun from the repo root): tests, bench lines, microbenchmarks,
# ncu launch list and ncu --set full exports.  Everything lands in gpurun_out/ (copied into profiles/ afterwards).
R=${1:-r1}
O=gpurun_out
mkdir -p $O
python -m pytest tests -q -m gpu 2>&1 | tail -3 > $O/pytest_gpu_$R.txt
python bench.py --steps 20 --warmup 5 > $O/bench_${R}_rmat.json 2> $O/bench_${R}_rmat.err
python bench.py --graph uniform --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_${R}_uniform.json 2> $O/bench_${R}_uniform.err
python bench.py --workload c2 --steps 20 --warmup 5 --no-cpu-baseline > $O/bench_${R}_c2_rmat.json 2>&1
python bench.py --dtype bf16 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > $O/bench_${R}_rmat_bf16.json 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_${R}_reference.json 2> $O/bench_${R}_reference.err
python tools/microbench.py c3 > $O/microbench_${R}_c3_rmat.jsonl 2> $O/micro_c3.err
python tools/microbench.py c4 > $O/microbench_${R}_c4_rmat.jsonl 2> $O/micro_c4.err
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"spmm|ggl|sddmm|softmax" -c 40 --csv --log-file $O/launches_${R}_final.csv \
    python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_bench.log 2>&1
for g in rmat uniform; do
  ncu --set full --clock-control none --import-source on -k regex:spmm_ring -s 4 -c 1 -o $O/prof_ring_$g \
      python bench.py --graph $g --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > $O/ncu_full_$g.log 2>&1
  ncu -i $O/prof_ring_$g.ncu-rep --page raw --csv > $O/ncu_${R}_spmm_ring_${g}_raw.csv 2>/dev/null
  ncu -i $O/prof_ring_$g.ncu-rep --page details --csv > $O/ncu_${R}_spmm_ring_${g}_details.csv 2>/dev/null
  # stamp: bench.py only trusts a capture taken from the kernel sources it was built from (roofline.traffic)
  python -c "import json,hashlib;from gammagl_b200 import build as b;print(json.dumps({'source_digest':b._digest(),'lib_sha256':hashlib.sha256(open(b.LIB,'rb').read()).hexdigest(),'round':'$R','graph':'$g'}))" > $O/ncu_${R}_spmm_ring_${g}_raw.csv.stamp.json
  rm -f $O/prof_ring_$g.ncu-rep
done
# per-kernel durations of the GAT / softmax / segment-backward / ingest kernels (duration only: one pass per kernel)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"softmax|gather_rows|expand_rows|sddmm|gat_|pack_pairs|narrow_keys|rowptr_kernel|finalize" \
    -c 300 --csv --log-file $O/launches_${R}_c4_ops.csv python tools/microbench.py c4 > $O/ncu_gat.log 2>&1
ncu --set full --clock-control none -k regex:softmax_stats -c 1 -o $O/prof_gat \
    python tools/microbench.py c4 >> $O/ncu_gat.log 2>&1
ncu -i $O/prof_gat.ncu-rep --page raw --csv > $O/ncu_${R}_gat_kernels_raw.csv 2>/dev/null
rm -f $O/prof_gat.ncu-rep
ls -la $O
