#!/bin/sh
# profiles refresh on HEAD: launch list (time + DRAM bytes) of two bench-size batches, ncu --set full of the event loop
TAG=${1:-f3}
(timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv \
   --log-file gpurun_out/r2_launches_${TAG}_cfg3_65536.csv python tools/prof_step.py cfg3_4x64_sinusoid_120s 65536 2) > gpurun_out/r2_launches_${TAG}.log 2>&1; echo launch-list $?
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:dcsim_advance -s 1 -c 1 -f -o gpurun_out/r2_prof_${TAG}_advance \
   python tools/prof_step.py cfg3_4x64_sinusoid_120s 23680 2) > gpurun_out/r2_ncu_${TAG}_advance.log 2>&1; echo full-advance $?; grep -o '"events_per_batch": [0-9]*' gpurun_out/r2_ncu_${TAG}_advance.log
grep "dcsim" gpurun_out/r2_launches_${TAG}_cfg3_65536.csv | grep "gpu__time" | cut -d, -f1,5,15- | head -8
