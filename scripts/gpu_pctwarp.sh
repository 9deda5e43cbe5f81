#!/bin/bash
# GPU visit: K2w (warp-per-cell percentile kernel) -- parity, then A/B against the CTA-per-cell variant, then one ncu capture.
mkdir -p gpurun_out
TAG=${1:-r2k}
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "percentile or pct or fuzz or adversarial or width or packed" > gpurun_out/pytest_pct_$TAG.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_pct_$TAG.log
for v in 0 1; do
  SLAFB_PCT_CTA=$v timeout 400 python bench.py --workload cfg3p --value-dtype uint16 --no-loader --no-cpu-baseline --no-e2e --steps 40 > gpurun_out/bench_cfg3p_cta${v}_$TAG.json 2> gpurun_out/bench_cfg3p_cta${v}_$TAG.err
  echo "bench cta=$v rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_cfg3p_cta${v}_$TAG.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("  value %.1f M  ms/step %.4f  K2 %.4f ms frac %.3f iso %.3f whole %s gen %s" % (d["value"]/1e6, d["ms_per_step"], r["avg_launch_ms"], r["frac"], r["frac_isolated"], r["whole_step_hbm_frac"], r["general_path"]))
except Exception as ex:
    print("  (no line)", ex)
PY
  tail -2 gpurun_out/bench_cfg3p_cta${v}_$TAG.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tokenize_pct" -s 3 -c 1 -f -o gpurun_out/prof_pctw_$TAG \
    python scripts/prof_driver.py cfg3p 3 32768 uint16 > gpurun_out/ncu_full_pctw_$TAG.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/prof_pctw_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_raw_pctw_$TAG.csv 2>/dev/null
ncu -i gpurun_out/prof_pctw_$TAG.ncu-rep --page source --csv > gpurun_out/ncu_source_pctw_$TAG.csv 2>/dev/null
ncu -i gpurun_out/prof_pctw_$TAG.ncu-rep --page details > gpurun_out/ncu_details_pctw_$TAG.txt 2>/dev/null
rm -f gpurun_out/prof_pctw_$TAG.ncu-rep
du -sh gpurun_out
