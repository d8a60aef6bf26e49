#!/bin/bash
# round-2 call 11 (4 GPUs): the scaling command at N=4 and N=2 (strong-scaling curve of the partitioned path: 1 / 2 / 4 / 8)
O=gpurun_out; mkdir -p $O
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 20 --warmup 5 > $O/bench_r2k_4gpu.json 2> $O/bench_r2k_4gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus 2 --steps 20 --warmup 5 > $O/bench_r2k_2gpu.json 2> $O/bench_r2k_2gpu.err
timeout 200 python -m pytest tests/test_gpu_partition.py -q -m gpu 2>&1 | tail -3 > $O/pytest_gpu_r2k_partition.txt
tail -c 1500 $O/bench_r2k_4gpu.json; tail -2 $O/bench_r2k_4gpu.err; cat $O/pytest_gpu_r2k_partition.txt
