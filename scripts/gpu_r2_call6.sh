#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
echo "=== A/B"; LIBS="ab_libs/lib_kstep.so neural-volume-super-resolution_b200/libnvsr_b200.so" bash scripts/gpu_ab_mlp.sh 2>&1 | tee gpurun_out/ab_mlp.log
echo "=== tests (decoder + e2e)"; timeout 900 python -m pytest tests/test_gpu_stages.py tests/test_gpu_e2e.py tests/test_gpu_parity_chain.py -m gpu -q -x 2>&1 | tail -5
echo "=== bench"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_quick.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_quick.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['clocks'], {k:round(v['avg_ms'],3) for k,v in d['kernels'].items()})
PY
