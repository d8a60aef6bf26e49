#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 200 python -m pytest tests/test_gpu_linear_tc.py -q -m gpu 2>&1 | tail -3 | tee $O/pytest_gpu_r2_linear_tc_final.txt
GGLB200_LAYER_LINEAR=tc timeout 600 python -m pytest tests/test_gpu_layers.py tests/test_gpu_zz_reference_layers.py tests/test_gpu_plugin.py -q -m gpu 2>&1 | tail -3 | tee $O/pytest_gpu_r2_layers_tc_linear.txt
