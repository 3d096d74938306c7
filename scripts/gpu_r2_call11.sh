#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
echo "=== time mlp"; timeout 300 python scripts/time_mlp.py 2>&1 | grep NVSR_DBG
echo "=== tests (mip / generic chain users + all decoder)"; timeout 1200 python -m pytest tests/test_gpu_stages.py tests/test_gpu_e2e.py tests/test_gpu_parity_chain.py -m gpu -q --tb=short 2>&1 | tail -6
echo "=== cfg3b"; timeout 600 python bench.py --config cfg3b --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_cfg3b.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_cfg3b.json').read().strip().splitlines()[-1])
print(d['config']['workload'], d['ms_per_step'], {k:round(v['avg_ms'],3) for k,v in d['kernels'].items()})
PY
