# GPU parity tests + bench (no ncu).  Run under gpurun.
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
echo "=== gpu tests"; timeout 1500 python -m pytest tests -m gpu -q --maxfail=${MAXFAIL:-10} --tb=short -p no:cacheprovider ${PYTEST_ARGS:-} > gpurun_out/t_gpu.log 2>&1; echo "rc=$?"; grep -v "Warning" gpurun_out/t_gpu.log | tail -${TAIL:-60}
echo "=== bench"; timeout 900 python bench.py --steps 5 --warmup 3 ${BENCH_ARGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; tail -3 gpurun_out/bench.err; python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench.json'))
    print('ms/step', d['ms_per_step'], 'rays/s', d['value'], 'e2e', d['e2e']['value'])
    print('roofline', d['roofline'])
    for k,v in d['kernels'].items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items()})
    print('cpu', d['cpu_baseline'])
except Exception as e: print('bench parse failed', e)
PY
