# Same-box A/B of library variants under ab_libs/:  STAGE=composite|gather|all bash scripts/gpu_ab.sh
cd $GRAFT_REPO_ROOT
for lib in ab_libs/lib_*.so; do
  NVSR_B200_LIB=$PWD/$lib timeout 300 python scripts/time_stage.py ${STAGE:-all} 2>&1 | grep -E "composite|gather|Error|error" | tail -4
done
