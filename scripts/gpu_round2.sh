#!/bin/bash
# One GPU-box visit (round 2): smoke, parity tests, bench (both arms, driver-style --steps 20 too), the other workloads,
# ncu launch list of the bench command and ncu --set full of the dominant kernels.  Every step under its own timeout; ncu
# runs only after the same command exited 0 without it.  Outputs land in gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
TAG=${1:-r2b}
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/gpu_$TAG.txt 2>&1
nproc >> gpurun_out/gpu_$TAG.txt; free -g | head -2 >> gpurun_out/gpu_$TAG.txt; lscpu | head -20 >> gpurun_out/gpu_$TAG.txt
if [ -z "$SKIP_TESTS" ]; then
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_$TAG.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu -x --timeout 600 > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_$TAG.log
fi
run_bench() {  # name, args...
  local name=$1; shift
  timeout 700 python bench.py --watchdog 640 "$@" > gpurun_out/bench_${name}_$TAG.json 2> gpurun_out/bench_${name}_$TAG.err; local rc=$?
  echo "bench $name rc=$rc"; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${name}_$TAG.json").read().strip().splitlines()[-1])
    r=d.get("roofline",{}); e=d.get("e2e",{}) or {}
    print("  value %.1f M  ms/step %.4f  frac %s iso %s whole %s | e2e %.3f M (coo %s) | host %s" % (d["value"]/1e6, d["ms_per_step"], r.get("frac"), r.get("frac_isolated"), r.get("whole_step_hbm_frac"), (e.get("value") or 0)/1e6, (e.get("coo_route") or {}).get("value"), d.get("host_us_per_step")))
    if "loader" in e: print("  loader", {k:v for k,v in e["loader"].items() if k.startswith("cells")})
except Exception as ex:
    print("  (no line)", ex)
PY
  tail -3 gpurun_out/bench_${name}_$TAG.err
  return $rc
}
echo "== bench cfg2 (default)"; run_bench cfg2; rc=$?
echo "== bench cfg2 driver-style"; run_bench cfg2_20 --steps 20 --warmup 5 --no-loader --no-cpu-baseline
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"; tail -c 600 gpurun_out/bench_ref_$TAG.json
if [ -z "$SKIP_OTHERS" ]; then
echo "== other workloads"
run_bench cfg3p --workload cfg3p --no-loader --no-cpu-baseline
run_bench cfg2_f32 --value-dtype float32 --no-loader --no-cpu-baseline
run_bench cfg3 --workload cfg3 --no-loader --no-cpu-baseline
run_bench cfg3_f32 --workload cfg3 --value-dtype float32 --no-loader --no-cpu-baseline --no-e2e
run_bench cfg4 --workload cfg4 --no-loader --no-cpu-baseline
fi
if [ $rc -eq 0 ] && [ -z "$SKIP_NCU" ]; then
  echo "== ncu launch list (bench.py, short)"
  timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-loader > gpurun_out/bench_short_$TAG.json 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches_$TAG.csv \
      python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-loader > gpurun_out/ncu_list_$TAG.log 2>&1
  echo "list rc=$?"
  for spec in "cfg2 uint16" "cfg3p uint16" "cfg2 float32" "cfg4 uint16"; do
    set -- $spec
    echo "== ncu full $1 $2"
    timeout 300 python scripts/prof_driver.py $1 6 32768 $2 > gpurun_out/prof_plain_$1_$2_$TAG.log 2>&1 &&
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tokenize_cells|csr_build|tokenize_rank|gene_stats" -s 4 -c 6 -f -o gpurun_out/prof_$1_$2_$TAG \
        python scripts/prof_driver.py $1 4 32768 $2 > gpurun_out/ncu_full_$1_$2_$TAG.log 2>&1
    echo "full rc=$?"; tail -2 gpurun_out/prof_plain_$1_$2_$TAG.log
  done
fi
ls -la gpurun_out/ | tail -40
