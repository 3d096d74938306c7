#!/bin/bash
# bias-as-K-step decoder: A/B timing alone, tests, bench
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
echo "=== A/B"; LIBS="ab_libs/lib_prestore.so neural-volume-super-resolution_b200/libnvsr_b200.so" bash scripts/gpu_ab_mlp.sh 2>&1 | tee gpurun_out/ab_mlp.log
bash scripts/gpu_tests.sh
