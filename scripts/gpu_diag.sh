cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python tests/diag_e2e.py fp32 2>&1 | grep -v Warning | tail -80
