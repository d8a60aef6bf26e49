#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3 > $O/pytest_gpu_r2_final.txt; cat $O/pytest_gpu_r2_final.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee $O/smoke_r2_final.txt
