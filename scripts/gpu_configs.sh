#!/bin/bash
# The other BASELINE.json configurations on one GPU with the current build (parity-gated like every bench run).
TAG=${1:-r1}
mkdir -p gpurun_out
for spec in "cfg1 uint16" "cfg3 uint16" "cfg3p uint16" "cfg4 uint16" "cfg2 float32" "cfg3 float32"; do
  set -- $spec; W=$1; VD=$2
  out=gpurun_out/bench_${W}_${VD}_${TAG}.json
  timeout 500 python bench.py --workload $W --value-dtype $VD --steps 60 --warmup 5 --no-loader > $out 2> ${out%.json}.err
  echo "$W $VD rc=$?"
  python - "$out" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().split("\n")[-1]); r=d["roofline"]
print("   value %.1fM ms/step %.4f whole %s | K2 iso %.4f ms frac %.2f | e2e %.3fM (%.1f GB/s) | cpu %.0f" % (d["value"]/1e6, d["ms_per_step"], ("%.2f" % r["whole_step_hbm_frac"]) if r.get("whole_step_hbm_frac") else "-", r["isolated"]["tokenize_cells_kernel"]["avg_launch_ms"], r["isolated"]["tokenize_cells_kernel"]["frac"], d["e2e"]["value"]/1e6, d["e2e"]["h2d_gbs_per_gpu"], d.get("cpu_baseline",{}).get("value",0)))
PY
done
