#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_quick.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['per_step'], 'e2e', d['e2e']['ms_per_step'], 'sparse', d['sparse']['ms_per_step'], d['clocks'])
print('roofline frac', d['roofline']['frac'], d['roofline']['achieved'])
print({k:round(v['avg_ms'],3) for k,v in d['kernels'].items()})
PY
