# Full GPU pass: parity tests, smoke, bench, ncu launch list, ncu full captures.  Run under gpurun.
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv | tail -2
echo "=== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=40 --tb=short -p no:cacheprovider > gpurun_out/t_gpu.log 2>&1; echo "rc=$?"; grep -v "Warning" gpurun_out/t_gpu.log | tail -40
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "=== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
echo "=== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/launches.out 2>&1; echo "rc=$?"; wc -l gpurun_out/launches.csv
echo "=== ncu full"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"mlp_chain_tc|gather_tile|gather_rows|keep_rows|composite_kernel" -s 0 -c 12 -o gpurun_out/prof_full -f python scripts/prof_frame.py --rows 400 > gpurun_out/prof_full.out 2>&1; echo "rc=$?"; tail -3 gpurun_out/prof_full.out; ls -la gpurun_out
