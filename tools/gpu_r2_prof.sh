#!/bin/sh
# round 2 profiling call: launch list of one bench-size batch, ncu --set full of the three kernels (8192 replicas)
TAG=${1:-s2}
python __graft_entry__.py > /dev/null 2>&1
(timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv \
   --log-file gpurun_out/r2_launches_${TAG}_cfg3_65536.csv python tools/prof_step.py cfg3_4x64_sinusoid_120s 65536 2) > gpurun_out/r2_launches_${TAG}.log 2>&1; echo launch-list $?
for K in advance arrivals merge; do
  (timeout 900 ncu --set full --clock-control none --import-source on -k regex:dcsim_$K -s 1 -c 1 -f -o gpurun_out/r2_prof_${TAG}_$K \
     python tools/prof_step.py cfg3_4x64_sinusoid_120s 8192 2) > gpurun_out/r2_ncu_${TAG}_$K.log 2>&1; echo full-$K $?; tail -1 gpurun_out/r2_ncu_${TAG}_$K.log
done
python - <<PY
import csv
rows = list(csv.reader(open("gpurun_out/r2_launches_${TAG}_cfg3_65536.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
cols = rows[hdr]
ki, mi, vi = cols.index("Kernel Name"), cols.index("Metric Name"), cols.index("Metric Value")
for r in rows[hdr + 1:]:
    if len(r) > vi: print(r[0], r[ki][:40], r[mi], r[vi])
PY
