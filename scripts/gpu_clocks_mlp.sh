# SM clock and power while the decoder kernel alone runs back to back (is it power/clock limited?)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,clocks_throttle_reasons.active --format=csv,noheader -lms 50 > gpurun_out/clocks_mlp.csv &
SMI=$!
sleep 0.5
REPS=${REPS:-1500} timeout 300 python scripts/time_mlp.py 2>&1 | grep -E "density|rgb"
sleep 0.3
kill $SMI
awk -F, '{print $1, $3}' gpurun_out/clocks_mlp.csv | sort | uniq -c | sort -k2,2n | tail -25
