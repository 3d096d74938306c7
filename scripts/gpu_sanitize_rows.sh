# compute-sanitizer memcheck over the kernels added late in round 2: row-list training backward, fused weight gradients, sparse /
# precise training forward, hi/lo gather, sort_cat, single-launch weight packing
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 python -m pytest tests/test_gpu_train_tc.py tests/test_gpu_stages.py -m gpu -q -x -k "row_list or dgrad or wgrad or frozen or sparse_training or pack_weights or precise_density or hilo or sort_cat" -p no:cacheprovider > gpurun_out/sanitize_rows_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitize_rows_memcheck.log | tail -6
