#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
echo "=== split parity"; timeout 900 python -m pytest tests/test_gpu_parity_chain.py -m gpu -q --tb=short -k "split" -rA 2>&1 | grep -E "passed|failed|Error|assert|^\{|FAILED|SKIPPED" | cut -c1-900 | tail -30
echo "=== bench (precision modes)"; timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_split.json 2> gpurun_out/bench_split.err; tail -3 gpurun_out/bench_split.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_split.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], json.dumps(d['precision_modes'], indent=0)[:1200])
PY
