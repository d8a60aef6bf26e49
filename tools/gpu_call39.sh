#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 6 --warmup 3 --no-configs > $O/bench_r2_final_2gpu.json 2> $O/bench_r2_final_2gpu.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_final_2gpu.json').read().strip().split('\n')[-1])
p=d['partitioned']
print(d['value'], p['parity_ok'], {k:(v.get('gcn_3_layer'), v.get('value')) for k,v in p.items() if isinstance(v,dict)})
PY
tail -2 $O/bench_r2_final_2gpu.err
