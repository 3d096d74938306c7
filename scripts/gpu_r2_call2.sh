#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
echo "=== dumps"; timeout 600 python tests/diag_dump.py > gpurun_out/dump.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/dump.log
echo "=== torch frame"; timeout 600 python scripts/bench_torch_frame.py --frames 2 > gpurun_out/torch_frame.json 2> gpurun_out/torch_frame.err; echo "rc=$?"; tail -2 gpurun_out/torch_frame.err; cat gpurun_out/torch_frame.json
du -sh gpurun_out
