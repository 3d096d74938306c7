#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
echo "=== split parity + fp32 gather users"; timeout 1200 python -m pytest tests/test_gpu_parity_chain.py tests/test_gpu_stages.py tests/test_gpu_e2e.py -m gpu -q --tb=short -k "split or fp32 or planes_forward or gather" 2>&1 | grep -E "passed|failed|Error|assert|FAILED" | cut -c1-300 | tail -12
bash scripts/gpu_r2_split_bench.sh
