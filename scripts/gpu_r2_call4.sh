#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
echo "=== tests"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/t_gpu.log; cat gpurun_out/t_gpu.log
echo "=== diag grads"; timeout 300 python tests/diag_train_grads.py 2>&1 | tail -100 > gpurun_out/diag_grads.log; cat gpurun_out/diag_grads.log
echo "=== train step"; timeout 600 python scripts/bench_train_step.py --steps 20 > gpurun_out/train_step.json 2> gpurun_out/train_step.err; echo "rc=$?"; tail -2 gpurun_out/train_step.err; cat gpurun_out/train_step.json
echo "=== umma wgrad"; (nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I neural-volume-super-resolution_b200/csrc -I include -o /tmp/umma_wgrad scripts/ubench/umma_wgrad.cu && for k in 48 128 144; do timeout 60 /tmp/umma_wgrad $k 64 0; done; timeout 60 /tmp/umma_wgrad 128 64 1) > gpurun_out/umma_wgrad.log 2>&1; cat gpurun_out/umma_wgrad.log
