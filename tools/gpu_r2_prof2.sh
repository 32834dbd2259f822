#!/bin/sh
# one ncu --set full capture of the event loop at two full waves of the 8-lane build (23 680 replicas of cfg 3)
TAG=${1:-s14}
REPS=${2:-23680}
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:dcsim_advance -s 1 -c 1 -f -o gpurun_out/r2_prof_${TAG}_advance \
   python tools/prof_step.py cfg3_4x64_sinusoid_120s $REPS 2) > gpurun_out/r2_ncu_${TAG}_advance.log 2>&1; echo full-advance $?; tail -1 gpurun_out/r2_ncu_${TAG}_advance.log
