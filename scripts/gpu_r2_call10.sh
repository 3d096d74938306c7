#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
echo "=== A/B"; LIBS="ab_libs/lib_head.so neural-volume-super-resolution_b200/libnvsr_b200.so" bash scripts/gpu_ab_mlp.sh 2>&1 | grep -E "^==|NVSR_DBG"
echo "=== tests"; timeout 1200 python -m pytest tests/test_gpu_stages.py tests/test_gpu_e2e.py tests/test_gpu_parity_chain.py tests/test_gpu_train_tc.py tests/test_gpu_sr.py -m gpu -q -x --tb=short 2>&1 | tail -15
