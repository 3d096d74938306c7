# ncu full capture of the hot kernels on one ray chunk (coarse + fine).  Run under gpurun.
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-mlp_chain_tc|gather_tile|gather_rows|keep_rows|composite_kernel}" -s ${SKIP:-0} -c ${COUNT:-12} -o gpurun_out/prof_full -f python scripts/prof_frame.py --rows ${ROWS:-400} > gpurun_out/prof_full.out 2>&1; echo "rc=$?"; tail -3 gpurun_out/prof_full.out
