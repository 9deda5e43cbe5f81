#!/bin/bash
# full GPU suite (product + checked build), then float32 percentile A/B (K2f vs general kernel) and the long-cell sweep rows
mkdir -p gpurun_out
TAG=${1:-r2u}
timeout 1200 python -m pytest tests -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_$TAG.log
bash scripts/checked_build.sh run $TAG | tail -3
show() {
  python - <<PY
import json
try:
    d=json.loads(open("$1").read().strip().splitlines()[-1]); r=d["roofline"]
    print("  value %.1f M  ms/step %.4f  K2 %.4f ms frac %s iso %.3f whole %s gen %s %s" % (d["value"]/1e6, d["ms_per_step"], r["avg_launch_ms"], r["frac"], r["frac_isolated"], r["whole_step_hbm_frac"], r["general_path"], r["kernel"]))
except Exception as ex:
    print("  (no line)", ex)
PY
}
SLAFB_F32_CTA=1 timeout 400 python bench.py --workload cfg3p --value-dtype float32 --no-loader --no-cpu-baseline --no-e2e --steps 20 > gpurun_out/bench_cfg3p_float32_general_$TAG.json 2> gpurun_out/bench_cfg3p_float32_general_$TAG.err
echo "bench cfg3p f32 general rc=$?"; show gpurun_out/bench_cfg3p_float32_general_$TAG.json
timeout 400 python bench.py --workload cfg3p --value-dtype float32 --no-loader --no-cpu-baseline --no-e2e --steps 40 > gpurun_out/bench_cfg3p_float32_$TAG.json 2> gpurun_out/bench_cfg3p_float32_$TAG.err
echo "bench cfg3p f32 warp rc=$?"; show gpurun_out/bench_cfg3p_float32_$TAG.json; tail -2 gpurun_out/bench_cfg3p_float32_$TAG.err
timeout 400 python bench.py --workload cfg3p --value-dtype uint16 --no-loader --no-cpu-baseline --no-e2e --steps 40 > gpurun_out/bench_cfg3p_uint16_$TAG.json 2> gpurun_out/bench_cfg3p_uint16_$TAG.err
echo "bench cfg3p u16 rc=$?"; show gpurun_out/bench_cfg3p_uint16_$TAG.json
