#!/bin/bash
# compute-sanitizer passes over smoke() and the parity tests that stress the hazardous code paths (K2's read-before-CAS hash set,
# the double-buffered stage / descriptor reuse, K1's spin-gather).  One tool per pass, each under its own timeout; the summaries
# land in gpurun_out/sanitizer_<tool>_<tag>.txt (copied to profiles/ by hand).
mkdir -p gpurun_out
TAG=${1:-r2}
SEL=${2:-"adversarial or float32_distinct or many_tiny or fuzz"}
TOOLS=${3:-"memcheck racecheck synccheck"}
export SLAFB_SANITIZE=1
for tool in $TOOLS; do
  out=gpurun_out/sanitizer_${tool}_$TAG.txt
  echo "== compute-sanitizer --tool $tool : smoke()" > $out
  timeout -s KILL 600 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 9 python __graft_entry__.py smoke >> $out 2>&1
  echo "smoke rc=$?" >> $out
  echo "== compute-sanitizer --tool $tool : pytest tests/test_parity_gpu.py -k \"$SEL\"" >> $out
  timeout -s KILL 1500 compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -q -x -m gpu -k "$SEL" -p no:cacheprovider >> $out 2>&1
  echo "pytest rc=$?" >> $out
  echo "$tool:"; grep -E "rc=|ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" $out | tail -8
done
