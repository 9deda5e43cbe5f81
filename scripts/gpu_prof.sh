#!/bin/bash
# ncu: launch list + full capture of K1/K2 on the fixed profiling workload (1 GPU).
mkdir -p gpurun_out
W=${1:-cfg2}
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.sw_power_cap --format=csv -i 0 > gpurun_out/smi_probe.txt 2>&1
timeout 300 python scripts/prof_driver.py $W 6 > gpurun_out/prof_plain_$W.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_$W.csv python scripts/prof_driver.py $W 6 > gpurun_out/ncu_list_$W.log 2>&1
echo "list rc=$?"; cat gpurun_out/prof_plain_$W.log | tail -3
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tokenize_cells|csr_build" -s 6 -c 4 -f -o gpurun_out/prof_$W python scripts/prof_driver.py $W 5 > gpurun_out/ncu_full_$W.log 2>&1
echo "full rc=$?"; tail -3 gpurun_out/ncu_full_$W.log; ls -la gpurun_out/
