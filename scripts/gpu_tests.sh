#!/bin/bash
# Full GPU test pass + smoke; the parity-chain figures go to gpurun_out/parity_chain.jsonl.
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out; rm -f gpurun_out/parity_chain.jsonl
export NVSR_PARITY_REPORT=$GRAFT_REPO_ROOT/gpurun_out/parity_chain.jsonl
timeout 1500 python -m pytest tests -m gpu -q --maxfail=30 --tb=short -p no:cacheprovider -rA > gpurun_out/t_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|error" gpurun_out/t_gpu.log | tail -5; grep -E "^(FAILED|ERROR)" gpurun_out/t_gpu.log | head -40
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -v Warning | tail -6
echo "=== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
