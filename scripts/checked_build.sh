#!/bin/bash
# The sanitizer substitute (compute-sanitizer is closed on this pool): build the library with -DSLAFB_CHECKED (device-side bounds
# checks on every stage / table / output index, kernels.cuh) and run the whole GPU test-suite plus smoke() against it.
# Usage: scripts/checked_build.sh build   (here, no GPU)      scripts/checked_build.sh run <tag>   (on the GPU box)
set -u
cd "$(dirname "$0")/.."
if [ "${1:-build}" = build ]; then
  mkdir -p build_checked
  env -u CC -u CXX nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -DSLAFB_CHECKED -Xcompiler -fPIC -shared \
      -o build_checked/libslaf_b200_checked.so slaf_b200/csrc/slaf_b200.cu && echo built build_checked/libslaf_b200_checked.so
  exit $?
fi
TAG=${2:-r2}
mkdir -p gpurun_out
export SLAFB_LIB=$PWD/build_checked/libslaf_b200_checked.so
export SLAFB_CHECK_REPORT=$PWD/gpurun_out/checked_build_$TAG.txt
echo "checked build: $(ls -la $SLAFB_LIB)" > $SLAFB_CHECK_REPORT
timeout 1800 python -m pytest tests -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/pytest_checked_$TAG.log 2>&1
echo "pytest (checked build) rc=$?" >> $SLAFB_CHECK_REPORT
tail -4 gpurun_out/pytest_checked_$TAG.log >> $SLAFB_CHECK_REPORT
cat $SLAFB_CHECK_REPORT
