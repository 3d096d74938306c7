cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python bench.py --precision fp16-split --steps 4 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_split.json 2> gpurun_out/bench_split.err; tail -2 gpurun_out/bench_split.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_split.json').read().strip().splitlines()[-1])
print(round(d['ms_per_step'],2), 'sparse', round(d['sparse']['ms_per_step'],2))
print({k:(v['launches'], round(v['avg_ms'],3), round(v['launches']*v['avg_ms']/ (d['steps']+0),1)) for k,v in d['kernels'].items()})
PY
