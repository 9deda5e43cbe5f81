#!/bin/bash
# full GPU suite on the product build and the checked build, smoke(), then the percentile / plain benches
mkdir -p gpurun_out
TAG=${1:-r2r}
timeout 1200 python -m pytest tests -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$TAG.log
bash scripts/checked_build.sh run $TAG | tail -4
SKIP_NCU=1 bash scripts/gpu_focus.sh $TAG "cfg3p:uint16 cfg3:uint16 cfg2:uint16"
