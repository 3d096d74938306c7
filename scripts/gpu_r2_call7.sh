#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
echo "=== train tc tests"; timeout 600 python -m pytest tests/test_gpu_train_tc.py -m gpu -q --tb=short -rA 2>&1 | tail -60 | tee gpurun_out/t_train_tc.log
