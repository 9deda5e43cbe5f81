#!/bin/bash
# The round's closing one-GPU visit: full GPU suite (normal + checked build), every bench workload, the cfg5 sweep, the ncu
# launch list of the bench command and light `ncu --set full` captures of the dominant kernels (exported to CSV / text on the
# box; the .ncu-rep files are deleted so gpurun_out/ stays small), and the sharded loader over a 1 M-cell table.
mkdir -p gpurun_out
TAG=${1:-r2z}
O=gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv > $O/gpu_$TAG.txt 2>&1; nproc >> $O/gpu_$TAG.txt; free -g | head -2 >> $O/gpu_$TAG.txt
echo "== pytest gpu"; timeout 1500 python -m pytest tests -q -m gpu --timeout 600 > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu_$TAG.log
echo "== checked build"; bash scripts/checked_build.sh run $TAG | tail -4
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 $O/smoke_$TAG.log | cut -c1-200
summ() { python - <<PY
import json
try:
    d=json.loads(open("$1").read().strip().splitlines()[-1]); r=d.get("roofline",{}); e=d.get("e2e") or {}
    print("  value %.1f M  ms/step %.4f  K2 %.4f ms frac %.3f iso %.3f whole %s | e2e %.3f M coo %s packed %s" % (d["value"]/1e6, d["ms_per_step"], r.get("avg_launch_ms") or 0, r.get("frac") or 0, r.get("frac_isolated") or 0, r.get("whole_step_hbm_frac"), (e.get("value") or 0)/1e6, (e.get("coo_route") or {}).get("value"), (e.get("packed_route") or {}).get("value")))
    if "loader" in e: print("  loader", {k:round(v/1e6,3) for k,v in e["loader"].items() if k.startswith("cells")})
except Exception as ex: print("  (no line)", ex)
PY
}
echo "== bench cfg2 default"; timeout 700 python bench.py > $O/bench_cfg2_$TAG.json 2> $O/bench_cfg2_$TAG.err; echo rc=$?; summ $O/bench_cfg2_$TAG.json
echo "== bench cfg2 driver-style"; timeout 400 python bench.py --steps 20 --warmup 5 --no-loader --no-cpu-baseline > $O/bench_cfg2_steps20_$TAG.json 2> $O/bench_cfg2_steps20_$TAG.err; echo rc=$?; summ $O/bench_cfg2_steps20_$TAG.json
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_reference_cfg2_$TAG.json 2> $O/bench_reference_$TAG.err; echo rc=$?
for spec in "cfg1:uint16" "cfg3:uint16" "cfg3p:uint16" "cfg2:float32" "cfg3:float32" "cfg4:uint16"; do
  W=${spec%%:*}; DT=${spec##*:}
  echo "== bench $W $DT"; timeout 400 python bench.py --workload $W --value-dtype $DT --no-loader --no-cpu-baseline > $O/bench_${W}_${DT}_$TAG.json 2> $O/bench_${W}_${DT}_$TAG.err; echo rc=$?; summ $O/bench_${W}_${DT}_$TAG.json
done
echo "== sweep"; timeout 900 python scripts/sweep.py --out $O/sweep_cfg5_$TAG.json > $O/sweep_$TAG.log 2>&1; echo rc=$?
echo "== sharded loader, 1 M-cell table"; timeout 900 python bench.py --workload cfg3 --sharded --table-cells 1048576 > $O/bench_sharded_1gpu_$TAG.json 2> $O/bench_sharded_1gpu_$TAG.err; echo rc=$?; python -c "
import json; d=json.loads(open('$O/bench_sharded_1gpu_$TAG.json').read().strip().splitlines()[-1]); print('  ', d['value']/1e6, d['per_rank'], d['every_cell_exactly_once']['ok'], d['config']['source_mode'])"
if [ -z "$SKIP_NCU" ]; then
  echo "== ncu launch list (bench.py, short)"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/launches_bench_cfg2_$TAG.csv \
      python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-loader > $O/ncu_list_$TAG.log 2>&1; echo "list rc=$?"
  for spec in "cfg2:uint16" "cfg3p:uint16" "cfg2:float32"; do
    W=${spec%%:*}; DT=${spec##*:}; name=${W}_${DT}
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tokenize_cells|tokenize_pct_warp|tokenize_f32_warp|csr_build" -s 4 -c 2 -f -o $O/prof_${name}_$TAG \
        python scripts/prof_driver.py $W 3 32768 $DT > $O/ncu_full_${name}_$TAG.log 2>&1; echo "ncu $name rc=$?"
    ncu -i $O/prof_${name}_$TAG.ncu-rep --page raw --csv > $O/ncu_raw_${name}_$TAG.csv 2>/dev/null
    ncu -i $O/prof_${name}_$TAG.ncu-rep --page source --csv > $O/ncu_source_${name}_$TAG.csv 2>/dev/null
    ncu -i $O/prof_${name}_$TAG.ncu-rep --page details > $O/ncu_details_${name}_$TAG.txt 2>/dev/null
    rm -f $O/prof_${name}_$TAG.ncu-rep
  done
fi
du -sh $O
