cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_stages.py tests/test_gpu_parity_chain.py tests/test_gpu_e2e.py tests/test_gpu_sr.py -m gpu -q -x -k "hilo or split or sparse or fp32_mode or tensor_core_modes" 2>&1 | tail -6
bash scripts/gpu_split_breakdown.sh
