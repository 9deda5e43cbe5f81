#!/bin/bash
# GPU visit: K2w (warp-per-cell percentile kernel) -- parity, then A/B of its variants (value stage on/off x stage size) and the
# CTA-per-cell kernel, then ncu captures.   Usage: scripts/gpu_pctwarp.sh <tag> "<vstage:vcap> ..." "<vstage:vcap for ncu> ..."
mkdir -p gpurun_out
TAG=${1:-r2k}
VARIANTS=${2:-"1:4104 1:3200 0:4104 0:3200 0:2568"}
NCU=${3:-"0:3200"}
for vs in 1 0; do
  SLAFB_PCT_VSTAGE=$vs SLAFB_PCT_VCAP=2304 timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "percentile or pct or fuzz or adversarial or width or packed" > gpurun_out/pytest_pct_vs${vs}_$TAG.log 2>&1
  echo "pytest vstage=$vs rc=$?"; tail -3 gpurun_out/pytest_pct_vs${vs}_$TAG.log
done
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
for v in $VARIANTS; do
  VS=${v%%:*}; VC=${v##*:}
  SLAFB_PCT_VSTAGE=$VS SLAFB_PCT_VCAP=$VC timeout 400 python bench.py --workload cfg3p --value-dtype uint16 --no-loader --no-cpu-baseline --no-e2e --steps 40 > gpurun_out/bench_cfg3p_vs${VS}_vc${VC}_$TAG.json 2> gpurun_out/bench_cfg3p_vs${VS}_vc${VC}_$TAG.err
  echo "bench vstage=$VS vcap=$VC rc=$?"
  show gpurun_out/bench_cfg3p_vs${VS}_vc${VC}_$TAG.json
done
for v in $NCU; do
  VS=${v%%:*}; VC=${v##*:}
  N=pctw_vs${VS}_vc${VC}_$TAG
  SLAFB_PCT_VSTAGE=$VS SLAFB_PCT_VCAP=$VC timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tokenize_pct" -s 3 -c 1 -f -o gpurun_out/prof_$N \
      python scripts/prof_driver.py cfg3p 3 32768 uint16 > gpurun_out/ncu_full_$N.log 2>&1
  echo "ncu $v rc=$?"
  ncu -i gpurun_out/prof_$N.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$N.csv 2>/dev/null
  ncu -i gpurun_out/prof_$N.ncu-rep --page source --csv > gpurun_out/ncu_source_$N.csv 2>/dev/null
  ncu -i gpurun_out/prof_$N.ncu-rep --page details > gpurun_out/ncu_details_$N.txt 2>/dev/null
  rm -f gpurun_out/prof_$N.ncu-rep
done
du -sh gpurun_out
