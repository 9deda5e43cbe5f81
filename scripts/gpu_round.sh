#!/bin/bash
# One GPU-box visit: smoke, parity tests, bench (both arms), ncu launch list of the bench command,
# ncu --set full of the two dominant kernels.  Every step under its own timeout; ncu runs only after
# the same command exited 0 without it.  Outputs land in gpurun_out/ (copied to profiles/ by hand).
mkdir -p gpurun_out
TAG=${1:-r1}
W=${2:-cfg2}
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
if [ -z "$SKIP_TESTS" ]; then
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1200 python -m pytest tests -q -m gpu -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
fi
echo "== bench"; timeout 600 python bench.py --workload $W --watchdog 540 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; rc=$?; echo "bench rc=$rc"; tail -c 3000 gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --workload $W --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"; tail -c 1500 gpurun_out/bench_ref_$TAG.json
if [ $rc -eq 0 ]; then
  echo "== ncu launch list (bench.py, short)"
  timeout 600 python bench.py --workload $W --steps 4 --warmup 3 --no-cpu-baseline --no-loader > gpurun_out/bench_short_$TAG.json 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
      python bench.py --workload $W --steps 4 --warmup 3 --no-cpu-baseline --no-loader > gpurun_out/ncu_list_$TAG.log 2>&1
  echo "list rc=$?"
  echo "== ncu full (prof_driver: one K1 + K2 sequence per step)"
  timeout 300 python scripts/prof_driver.py $W 6 32768 > gpurun_out/prof_plain_$TAG.log 2>&1 &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tokenize_cells|csr_build" -s 4 -c 4 -f -o gpurun_out/prof_$TAG \
      python scripts/prof_driver.py $W 4 32768 > gpurun_out/ncu_full_$TAG.log 2>&1
  echo "full rc=$?"; tail -3 gpurun_out/prof_plain_$TAG.log
fi
ls -la gpurun_out/
