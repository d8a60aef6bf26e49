#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_layers.py tests/test_gpu_sampler.py -q -m gpu 2>&1 | tail -3 | tee $O/pytest_gpu_r2_sage_aggfirst.txt
timeout 200 python examples/sage_minibatch.py --steps 40 2>&1 | tail -2 | tee $O/example_sage_minibatch_r2_final.txt
timeout 200 python examples/sage_minibatch.py --steps 40 --fuse-gather 2>&1 | tail -2 | tee $O/example_sage_minibatch_r2_final_fused_gather.txt
timeout 200 python examples/sage_minibatch.py --steps 40 --nodes 2449029 --edges 61859140 --batch 4096 2>&1 | tail -2 | tee $O/example_sage_minibatch_r2_products_shape.txt
timeout 200 python examples/sage_minibatch.py --steps 40 --nodes 2449029 --edges 61859140 --batch 4096 --fuse-gather 2>&1 | tail -2 | tee $O/example_sage_minibatch_r2_products_shape_fused_gather.txt
