#!/bin/bash
# round-2 call 6 (1 GPU): verification of the changes after call 4 (merged GAT backward, one-shot replicas + policy, sampler hub copy),
# then the round's evidence capture
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_layers.py tests/test_gpu_oneshot.py tests/test_gpu_sampler.py tests/test_gpu_plugin.py tests/test_gpu_fullsize.py tests/test_gpu_zz_reference_layers.py -q -m gpu -x 2>&1 | tail -15 > $O/pytest_gpu_r2f_quick.txt
cat $O/pytest_gpu_r2f_quick.txt | tail -5
bash tools/capture_round.sh r2 > $O/capture_r2.log 2>&1
tail -4 $O/pytest_gpu_r2.txt; tail -c 1200 $O/bench_r2_rmat.json; grep -h "gat_aggregate\|degree\|one-shot" $O/microbench_r2_c4_rmat.jsonl $O/microbench_r2_c3_rmat.jsonl $O/microbench_r2_c4_gat_two_pass.jsonl | cut -c1-200; tail -3 $O/sanitizer_r2_*.log; tail -5 $O/example_sage_minibatch_r2.txt
