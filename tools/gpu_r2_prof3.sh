#!/bin/sh
# ncu --set full of the three kernels: event loop at two full waves of the 8-lane build (23 680 replicas of cfg 3),
# pre-pass and merge at 32 768 replicas
TAG=${1:-s18}
for K in advance:23680 arrivals:32768 merge:32768; do
  N=${K%%:*}; R=${K##*:}
  (timeout 900 ncu --set full --clock-control none --import-source on -k regex:dcsim_$N -s 1 -c 1 -f -o gpurun_out/r2_prof_${TAG}_$N \
     python tools/prof_step.py cfg3_4x64_sinusoid_120s $R 2) > gpurun_out/r2_ncu_${TAG}_$N.log 2>&1; echo full-$N $?; grep -o '"events_per_batch": [0-9]*' gpurun_out/r2_ncu_${TAG}_$N.log
done
