#!/bin/bash
# Round-2 profile pass: TMEM micro-benchmark, ncu launch list of the bench command, ncu --set full of the top kernels.
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
echo "=== tmem ubench"; (nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tmem_bw scripts/ubench/tmem_bw.cu && timeout 60 /tmp/tmem_bw) 2>&1 | tee gpurun_out/tmem_bw.log
echo "=== ncu launch list"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/launches.out 2>&1; echo "rc=$?"; wc -l gpurun_out/launches.csv
echo "=== ncu full"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"mlp_chain_tc|gather_tile|composite_kernel" -s 0 -c 12 -o gpurun_out/prof_full -f python scripts/prof_frame.py --rows 400 > gpurun_out/prof_full.out 2>&1; echo "rc=$?"; tail -3 gpurun_out/prof_full.out; ls -la gpurun_out
