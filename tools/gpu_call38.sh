#!/bin/bash
# round-2 call 38 (2 GPUs): final tree: partition tests over NVLink + the driver's 2-GPU bench command
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_partition.py -q -m gpu 2>&1 | tail -3 | tee $O/pytest_gpu_r2_final_2gpu_partition.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_r2_final_2gpu.json 2> $O/bench_r2_final_2gpu.err; tail -c 1200 $O/bench_r2_final_2gpu.json; tail -3 $O/bench_r2_final_2gpu.err
