#!/bin/bash
# Round-2 verification call: whole GPU suite (incl. the SR tests), smoke, bench.
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
echo "=== tests"; timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -25 > gpurun_out/t_gpu.log; cat gpurun_out/t_gpu.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "=== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "rc=$?"; tail -2 gpurun_out/bench.err; cut -c1-1500 gpurun_out/bench.json
