#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 > $O/pytest_gpu_r2_final.txt; cat $O/pytest_gpu_r2_final.txt
timeout 400 python tools/bench_layer.py 256 > $O/bench_layer_r2_final.json 2>/dev/null; cat $O/bench_layer_r2_final.json | python -c "import json,sys; d=json.load(sys.stdin); print({k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k.startswith(('gcn2','linear','aggregate','transform'))})"
timeout 300 python tools/profile_model_step.py > $O/profile_model_step_r2.txt 2>&1; grep -E "Self CUDA time total|spmm_ring|linear_ws|sgemm|nll_loss_forward_reduce" $O/profile_model_step_r2.txt | cut -c1-200 | head -12
