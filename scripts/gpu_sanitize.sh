# compute-sanitizer (memcheck over the stage / training / split-chain tests, racecheck on the shared-memory kernels).
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
K='not full and not 800 and not large_bins and not constant_planes'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_stages.py tests/test_gpu_train_tc.py tests/test_gpu_sr.py -m gpu -q -x -k "$K and not tracks_fp32 and not train_step and not graphed" -p no:cacheprovider > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitize_memcheck.log | tail -8
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_parity_chain.py -m gpu -q -x -k "golden_chain and planes_det" -p no:cacheprovider > gpurun_out/sanitize_memcheck_chain.log 2>&1; echo "memcheck (render chains, 4 modes) rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitize_memcheck_chain.log | tail -4
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_stages.py -m gpu -q -x -k "composite or sample_pdf or volume_render" -p no:cacheprovider > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/sanitize_racecheck.log | tail -8
