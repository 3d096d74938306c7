#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_tc.py tests/test_gpu_next_rows.py -m gpu -q --tb=short 2>&1 | grep -E "passed|failed|^E  |FAILED" | head
timeout 600 python scripts/prof_train_cpu.py 2>&1 | grep "host enqueue"
timeout 600 python scripts/bench_train_step.py --steps 30 > gpurun_out/train_step.json 2> gpurun_out/train_step.err; tail -2 gpurun_out/train_step.err; cut -c1-2400 gpurun_out/train_step.json
