#!/bin/bash
# GPU visit: K2f (warp-per-cell kernel for float32 values) -- parity, A/B against the CTA-per-cell kernel, one ncu capture.
# Usage: scripts/gpu_f32warp.sh <tag> "<vcap> ..."
mkdir -p gpurun_out
TAG=${1:-r2s}
VCAPS=${2:-"2304"}
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_dropin_gpu.py -x -q -m gpu > gpurun_out/pytest_f32_$TAG.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_f32_$TAG.log
show() {
  python - <<PY
import json
try:
    d=json.loads(open("$1").read().strip().splitlines()[-1]); r=d["roofline"]
    print("  value %.1f M  ms/step %.4f  K2 %.4f ms frac %.3f iso %.3f whole %s gen %s" % (d["value"]/1e6, d["ms_per_step"], r["avg_launch_ms"], r["frac"], r["frac_isolated"], r["whole_step_hbm_frac"], r["general_path"]))
except Exception as ex:
    print("  (no line)", ex)
PY
}
for W in cfg2 cfg3; do
  SLAFB_F32_CTA=1 timeout 400 python bench.py --workload $W --value-dtype float32 --no-loader --no-cpu-baseline --no-e2e --steps 40 > gpurun_out/bench_${W}_float32_cta_$TAG.json 2> gpurun_out/bench_${W}_float32_cta_$TAG.err
  echo "bench $W cta rc=$?"; show gpurun_out/bench_${W}_float32_cta_$TAG.json
  for VC in $VCAPS; do
    SLAFB_F32_VCAP=$VC timeout 400 python bench.py --workload $W --value-dtype float32 --no-loader --no-cpu-baseline --no-e2e --steps 40 > gpurun_out/bench_${W}_float32_w${VC}_$TAG.json 2> gpurun_out/bench_${W}_float32_w${VC}_$TAG.err
    echo "bench $W warp vcap=$VC rc=$?"; show gpurun_out/bench_${W}_float32_w${VC}_$TAG.json; tail -1 gpurun_out/bench_${W}_float32_w${VC}_$TAG.err
  done
done
N=f32w_$TAG
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tokenize_f32_warp" -s 3 -c 1 -f -o gpurun_out/prof_$N \
    python scripts/prof_driver.py cfg2 3 32768 float32 > gpurun_out/ncu_full_$N.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/prof_$N.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$N.csv 2>/dev/null
ncu -i gpurun_out/prof_$N.ncu-rep --page source --csv > gpurun_out/ncu_source_$N.csv 2>/dev/null
ncu -i gpurun_out/prof_$N.ncu-rep --page details > gpurun_out/ncu_details_$N.txt 2>/dev/null
rm -f gpurun_out/prof_$N.ncu-rep
du -sh gpurun_out
