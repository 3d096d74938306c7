#!/bin/bash
# Round-2 first GPU call: trace dumps for the parity attribution work, the never-run round-1 scripts, fp32 frame time.
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv | tail -2
echo "=== dumps"; timeout 600 python tests/diag_dump.py > gpurun_out/dump.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/dump.log
echo "=== bench"; timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; tail -2 gpurun_out/bench.err; cat gpurun_out/bench.json
echo "=== bench fp32"; timeout 900 python bench.py --precision fp32 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; echo "rc=$?"; tail -2 gpurun_out/bench_fp32.err; cat gpurun_out/bench_fp32.json
echo "=== torch frame"; timeout 600 python scripts/bench_torch_frame.py --frames 2 > gpurun_out/torch_frame.json 2> gpurun_out/torch_frame.err; echo "rc=$?"; tail -2 gpurun_out/torch_frame.err; cat gpurun_out/torch_frame.json
echo "=== train step"; timeout 600 python scripts/bench_train_step.py --steps 20 > gpurun_out/train_step.json 2> gpurun_out/train_step.err; echo "rc=$?"; tail -2 gpurun_out/train_step.err; cat gpurun_out/train_step.json
echo "=== umma wgrad"; (nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I neural-volume-super-resolution_b200/csrc -I include -o /tmp/umma_wgrad scripts/ubench/umma_wgrad.cu && for k in 48 128 144; do timeout 60 /tmp/umma_wgrad $k 64 0; done; timeout 60 /tmp/umma_wgrad 128 64 1) > gpurun_out/umma_wgrad.log 2>&1; cat gpurun_out/umma_wgrad.log
ls -la gpurun_out | head -50; du -sh gpurun_out
