#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | grep -E "passed|failed|^E  |FAILED" | head
timeout 600 python scripts/prof_train_cpu.py 2>&1 | grep "host enqueue"
timeout 600 python scripts/bench_train_step.py --steps 30 > gpurun_out/train_step.json 2> gpurun_out/train_step.err; tail -2 gpurun_out/train_step.err; cut -c1-330 gpurun_out/train_step.json
timeout 600 python scripts/train_demo.py --steps 200 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:(v if not isinstance(v,dict) else {a:b for a,b in v.items() if a!='curve'}) for k,v in d.items()})"
bash scripts/gpu_r2_bench_quick.sh
