# Same-box A/B of decoder variants under ab_libs/ (scripts/time_mlp.py), run twice in alternation
cd $GRAFT_REPO_ROOT
for rep in 1 2; do
for lib in ${LIBS:-neural-volume-super-resolution_b200/libnvsr_b200.so ab_libs/lib_*.so}; do
  echo "== $lib"
  NVSR_B200_LIB=$PWD/$lib timeout 300 python scripts/time_mlp.py 2>&1 | grep -E "density|rgb|rror|warp" | grep -v Warning
done; done
