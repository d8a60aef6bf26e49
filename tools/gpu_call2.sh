#!/bin/bash
# round-2 call 2 (2 GPUs): peer-memory tests, robustness tests, bench --gpus 2 with the partitioned section (cache on / off)
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_partition.py tests/test_gpu_robustness.py -q -m gpu 2>&1 | tail -15 > $O/pytest_gpu_r2b.txt
timeout 850 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > $O/bench_r2b_2gpu.json 2> $O/bench_r2b_2gpu.err
GGLB200_PEER_CACHE_ROWS=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-c5 > $O/bench_r2b_2gpu_nocache.json 2> $O/bench_r2b_2gpu_nocache.err
tail -5 $O/pytest_gpu_r2b.txt; tail -c 3000 $O/bench_r2b_2gpu.json; tail -5 $O/bench_r2b_2gpu.err
