# N-GPU bench (torchrun) + reference arm.  Run under gpurun --gpus N with N=${N:-2}.
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
N=${N:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"
grep -v -i "warn\|OMP_NUM\|\*\*\*\*" gpurun_out/bench_n$N.err | tail -15
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_n$N.json")); print(d["n_gpus"], "ms", d["ms_per_step"], "rays/s", d["value"], d["e2e"], d["clocks"], "launches", d["gpu_launches"])
except Exception as e: print("parse failed", e)
PY
if [ "${REF:-1}" = "1" ]; then timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-700 gpurun_out/bench_ref.json; fi
