#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python scripts/prof_step.py > gpurun_out/ncu_l.log 2>&1; echo "exit $?"
for k in ${KERNELS:-conv_down32_ts conv_up32_ts conv_wgrad32_tc}; do
  echo "== full $k"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_$k python scripts/prof_step.py > gpurun_out/ncu_$k.log 2>&1; echo "exit $?"
done
ls -la gpurun_out/*.ncu-rep
