# Build a named variant of the library into ab_libs/ (git-ignored, shipped to the GPU box):
#   bash scripts/ab_build.sh <name> [extra nvcc flags, e.g. -DNVSR_COMP_MINB=8]
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p ab_libs
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared \
  -I include -I neural-volume-super-resolution_b200/csrc "$@" -o ab_libs/lib_$name.so neural-volume-super-resolution_b200/csrc/*.cu
echo built ab_libs/lib_$name.so
