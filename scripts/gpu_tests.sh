cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r1_gpu.txt
echo "=== fp32 tests" ; timeout 900 python -m pytest tests -m gpu -q -k "not bf16 and not full" --maxfail=40 --tb=short -p no:cacheprovider > gpurun_out/r1_t_fp32.log 2>&1; echo "rc=$?"; tail -40 gpurun_out/r1_t_fp32.log
echo "=== bf16 tests" ; timeout 900 python -m pytest tests -m gpu -q -k "bf16 or full" --maxfail=40 --tb=short -s -p no:cacheprovider > gpurun_out/r1_t_bf16.log 2>&1; echo "rc=$?"; tail -60 gpurun_out/r1_t_bf16.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
