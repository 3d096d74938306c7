#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train_tc.py tests/test_gpu_next_rows.py -m gpu -q --tb=short 2>&1 | grep -E "passed|failed|^E  |FAILED|Error" | head -20
timeout 600 python scripts/bench_train_step.py --steps 30 > gpurun_out/train_step.json 2> gpurun_out/train_step.err; tail -2 gpurun_out/train_step.err; python - <<PY
import json
r=json.load(open("gpurun_out/train_step.json"))
print({k:round(v,4) for k,v in r.items() if k.endswith("_ms") and k != "kernels_ms" or "rel_l2" in k})
print({k:round(v["total_ms"],3) for k,v in r["kernels_ms"].items()})
PY
