cd $GRAFT_REPO_ROOT
timeout 600 python bench.py --precision fp16-split --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_split2.json 2> gpurun_out/bench_split2.err; tail -3 gpurun_out/bench_split2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_split2.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['clocks'])
for k,v in d['kernels'].items(): print(k, v['launches'], round(v['avg_ms'],3))
PY
