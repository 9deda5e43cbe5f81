#!/bin/bash
# closing check of a build: full GPU suite on the product and the checked build, smoke(), one driver-style bench line
mkdir -p gpurun_out
TAG=${1:-r2fin}
timeout 1200 python -m pytest tests -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu_$TAG.log
bash scripts/checked_build.sh run $TAG | tail -2
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"
timeout 400 python bench.py --steps 20 --warmup 5 --no-loader --no-cpu-baseline > gpurun_out/bench_cfg2_steps20_$TAG.json 2> gpurun_out/bench_cfg2_steps20_$TAG.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_cfg2_steps20_$TAG.json").read().strip().splitlines()[-1]); r=d["roofline"]; e=d["e2e"]
print("value %.1f M ms/step %.4f frac %.3f iso %.3f whole %.3f e2e %.3f M" % (d["value"]/1e6, d["ms_per_step"], r["frac"], r["frac_isolated"], r["whole_step_hbm_frac"], e["value"]/1e6))
PY
