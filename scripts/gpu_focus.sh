#!/bin/bash
# Focused GPU visit: a few bench workloads (no loader, no CPU baseline) and a light ncu capture of the dominant kernel of each,
# exported to CSV on the box (the .ncu-rep files are deleted: gpurun_out/ must stay below 64 MiB).
# Usage: scripts/gpu_focus.sh <tag> "<workload:dtype> ..."     e.g.  scripts/gpu_focus.sh r2d "cfg3p:uint16 cfg2:float32 cfg4:uint16"
mkdir -p gpurun_out
TAG=${1:-r2d}
SPECS=${2:-"cfg3p:uint16 cfg2:float32 cfg4:uint16"}
for spec in $SPECS; do
  W=${spec%%:*}; DT=${spec##*:}
  name=${W}_${DT}
  timeout 400 python bench.py --workload $W --value-dtype $DT --no-loader --no-cpu-baseline --no-e2e --steps 40 > gpurun_out/bench_${name}_$TAG.json 2> gpurun_out/bench_${name}_$TAG.err
  echo "bench $name rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${name}_$TAG.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("  value %.1f M  ms/step %.4f  K2 %.4f ms frac %.3f iso %.3f whole %s gen %s" % (d["value"]/1e6, d["ms_per_step"], r["avg_launch_ms"], r["frac"], r["frac_isolated"], r["whole_step_hbm_frac"], r["general_path"]))
except Exception as ex:
    print("  (no line)", ex)
PY
  tail -2 gpurun_out/bench_${name}_$TAG.err
  if [ -z "$SKIP_NCU" ]; then
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tokenize_cells|tokenize_rank" -s 3 -c 1 -f -o gpurun_out/prof_${name}_$TAG \
        python scripts/prof_driver.py $W 3 32768 $DT > gpurun_out/ncu_full_${name}_$TAG.log 2>&1
    echo "ncu $name rc=$?"
    ncu -i gpurun_out/prof_${name}_$TAG.ncu-rep --page raw --csv > gpurun_out/ncu_raw_${name}_$TAG.csv 2>/dev/null
    ncu -i gpurun_out/prof_${name}_$TAG.ncu-rep --page source --csv > gpurun_out/ncu_source_${name}_$TAG.csv 2>/dev/null
    ncu -i gpurun_out/prof_${name}_$TAG.ncu-rep --page details > gpurun_out/ncu_details_${name}_$TAG.txt 2>/dev/null
    rm -f gpurun_out/prof_${name}_$TAG.ncu-rep
  fi
done
du -sh gpurun_out
