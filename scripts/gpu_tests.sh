cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
echo "=== fp32 tests" ; timeout 900 python -m pytest tests -m gpu -q -k "not bf16 and not fp16 and not 16bit and not full" --maxfail=40 --tb=short -s -p no:cacheprovider > gpurun_out/r1_t_fp32.log 2>&1; echo "rc=$?"; grep -v "^E    +\|Warning" gpurun_out/r1_t_fp32.log | tail -50
echo "=== 16-bit tests" ; timeout 900 python -m pytest tests -m gpu -q -k "bf16 or fp16 or 16bit or full" --maxfail=40 --tb=short -s -p no:cacheprovider > gpurun_out/r1_t_16.log 2>&1; echo "rc=$?"; grep -v "^E    +\|Warning" gpurun_out/r1_t_16.log | tail -70
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "=== bench"; timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err; echo "rc=$?"; tail -3 gpurun_out/r1_bench.err; cat gpurun_out/r1_bench.json
