#!/bin/bash
# round-2 call 8 (1 GPU): full suite after the softmax default change (forward CSR / backward sort-free), softmax numbers per path
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -8 > $O/pytest_gpu_r2h.txt
cat $O/pytest_gpu_r2h.txt
for p in auto oneshot csr; do
  GGLB200_SOFTMAX_PATH=$p timeout 300 python tools/microbench.py c4 2>/dev/null | grep softmax > $O/microbench_r2h_c4_softmax_$p.jsonl
done
cat $O/microbench_r2h_c4_softmax_*.jsonl | cut -c1-160
