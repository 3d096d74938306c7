cd $GRAFT_REPO_ROOT
for i in 1 2; do
for v in ${VARIANTS:-V0 V1 V2}; do echo "lib $v"; NVSR_B200_LIB=$GRAFT_REPO_ROOT/ab_libs/lib$v.so python scripts/time_mlp.py 2>&1 | grep NVSR_DBG; done
done
