#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_linear_tc.py -q -m gpu 2>&1 | tail -3 | tee $O/pytest_gpu_r2_linear_tc_final.txt
timeout 400 python tools/bench_layer.py 256 > $O/bench_layer_r2_final.json 2>/dev/null; cat $O/bench_layer_r2_final.json | python -c "import json,sys; d=json.load(sys.stdin); print({k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k.startswith(('gcn2','linear','aggregate','transform'))})"
GGLB200_LAYER_LINEAR=lib GGLB200_LINEAR_BWD=lib timeout 400 python tools/bench_layer.py 256 > $O/bench_layer_r2_final_lib_linear.json 2>/dev/null; cat $O/bench_layer_r2_final_lib_linear.json | python -c "import json,sys; d=json.load(sys.stdin); print({k:(round(v,3) if isinstance(v,float) else v) for k,v in d.items() if k.startswith(('gcn2'))})"
