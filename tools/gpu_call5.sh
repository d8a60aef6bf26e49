#!/bin/bash
# round-2 call 5 (8 GPUs): the driver's scaling command at N=8 (replicas + e2e + partitioned section with parity gate)
O=gpurun_out; mkdir -p $O
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench_r2i_8gpu.json 2> $O/bench_r2i_8gpu.err
nvidia-smi topo -m > $O/topo_r2i.txt 2>&1
tail -c 2500 $O/bench_r2i_8gpu.json; tail -4 $O/bench_r2i_8gpu.err
