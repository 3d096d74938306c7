#!/bin/bash
# Round-2 evidence pass: whole GPU suite, smoke, bench (all blocks), train step, ncu launch list, ncu --set full of the top kernels,
# compute-sanitizer on the new training kernels.
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; rm -f gpurun_out/parity_chain.jsonl
export NVSR_PARITY_REPORT=$GRAFT_REPO_ROOT/gpurun_out/parity_chain.jsonl
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv | tail -2
echo "=== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 --tb=short -p no:cacheprovider -rA > gpurun_out/t_gpu.log 2>&1; echo "rc=$?"; grep -E "passed|failed" gpurun_out/t_gpu.log | tail -3; grep -E "^(FAILED|ERROR)" gpurun_out/t_gpu.log | head
unset NVSR_PARITY_REPORT
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warn | tail -5
echo "=== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; tail -3 gpurun_out/bench.err; cut -c1-600 gpurun_out/bench.json
echo "=== train step"; timeout 600 python scripts/bench_train_step.py --steps 20 > gpurun_out/train_step.json 2> gpurun_out/train_step.err; echo "rc=$?"; tail -2 gpurun_out/train_step.err; cut -c1-400 gpurun_out/train_step.json
echo "=== sanitizer (training kernels)"; timeout 900 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_train_tc.py -m gpu -q -x -k "forward or dgrad or wgrad or row_list or pack_weights" 2>&1 | tail -8 > gpurun_out/sanitizer_train.txt; cat gpurun_out/sanitizer_train.txt
echo "=== ncu launch list (training step)"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_train.csv python scripts/bench_train_step.py --steps 1 --warmup 2 --no-torch > gpurun_out/launches_train.out 2>&1; echo "rc=$?"; wc -l gpurun_out/launches_train.csv
echo "=== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/launches.out 2>&1; echo "rc=$?"; wc -l gpurun_out/launches.csv
echo "=== ncu full"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"mlp_chain_tc|gather_tile|composite_kernel" -s 0 -c 12 -o gpurun_out/prof_full -f python scripts/prof_frame.py --rows 400 > gpurun_out/prof_full.out 2>&1; echo "rc=$?"; tail -3 gpurun_out/prof_full.out; ls -la gpurun_out | head -30
