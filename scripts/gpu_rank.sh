#!/bin/bash
# GPU visit: K6b (bucket-distribution rank-value kernel) -- extension tests on the product and the checked build, then A/B against
# the sorting network (SLAFB_RANK_SORT=1), one ncu capture.
mkdir -p gpurun_out
TAG=${1:-r2w}
timeout 900 python -m pytest tests/test_extension_gpu.py -x -q -m gpu > gpurun_out/pytest_ext_$TAG.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/pytest_ext_$TAG.log
SLAFB_LIB=$PWD/build_checked/libslaf_b200_checked.so SLAFB_CHECK_REPORT=$PWD/gpurun_out/checked_ext_$TAG.txt timeout 900 python -m pytest tests/test_extension_gpu.py -x -q -m gpu > gpurun_out/pytest_ext_checked_$TAG.log 2>&1
echo "pytest checked rc=$?"; tail -2 gpurun_out/pytest_ext_checked_$TAG.log
show() {
  python - <<PY
import json
try:
    d=json.loads(open("$1").read().strip().splitlines()[-1]); r=d["roofline"]
    print("  value %.2f M  ms/step %.4f  K2 %.4f ms frac %s gen %s %s" % (d["value"]/1e6, d["ms_per_step"], r["avg_launch_ms"], r["frac"], r["general_path"], d["config"].get("gene_stats")))
except Exception as ex:
    print("  (no line)", ex)
PY
}
for v in 1 0; do
  SLAFB_RANK_SORT=$v timeout 400 python bench.py --workload cfg4 --no-loader --no-cpu-baseline --no-e2e --steps 20 > gpurun_out/bench_cfg4_sort${v}_$TAG.json 2> gpurun_out/bench_cfg4_sort${v}_$TAG.err
  echo "bench cfg4 sort_only=$v rc=$?"; show gpurun_out/bench_cfg4_sort${v}_$TAG.json; tail -2 gpurun_out/bench_cfg4_sort${v}_$TAG.err
done
N=rank_$TAG
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tokenize_rank" -s 6 -c 3 -f -o gpurun_out/prof_$N \
    python scripts/prof_driver.py cfg4 3 32768 uint16 > gpurun_out/ncu_full_$N.log 2>&1
echo "ncu rc=$?"
ncu -i gpurun_out/prof_$N.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$N.csv 2>/dev/null
ncu -i gpurun_out/prof_$N.ncu-rep --page source --csv > gpurun_out/ncu_source_$N.csv 2>/dev/null
ncu -i gpurun_out/prof_$N.ncu-rep --page details > gpurun_out/ncu_details_$N.txt 2>/dev/null
rm -f gpurun_out/prof_$N.ncu-rep
