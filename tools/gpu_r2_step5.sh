#!/bin/sh
# round 2, step 5: default lanes-per-replica choice — full parity suite, occupancy variants of the 8-lane build, ncu of it
TAG=s5
(timeout 1800 python -m pytest tests -m gpu -x -q) > gpurun_out/r2_pytest_gpu_$TAG.log 2>&1; tail -4 gpurun_out/r2_pytest_gpu_$TAG.log
timeout 2400 python tools/ab2.py --rounds 2 --cases cfg3,cfg2,sweep_default,cfg5s \
  g32:DCSIM_GROUP=32 g8:DCSIM_GROUP=8 g8c5:lib=variants/libdcsim_g8c5.so:DCSIM_GROUP=8 g8c6:lib=variants/libdcsim_g8c6.so:DCSIM_GROUP=8 auto > gpurun_out/r2_ab_$TAG.jsonl 2> gpurun_out/r2_ab_$TAG.err
python - <<PY
import json
for l in open("gpurun_out/r2_ab_$TAG.jsonl"):
    d = json.loads(l)
    print(d.get("case"), d.get("variant"), d.get("round"), "pre %.1f adv %.1f" % (d.get("prepass_ms", -1), d.get("advance_ms", -1)),
          "total %.3f Gev/s" % (d.get("gev_s", -1)), "warps", d.get("warps_per_sm"), "regs", d.get("regs"), "mode", d.get("mode"), "failed", d.get("failed"), d.get("error", "")[-300:])
PY
sh tools/gpu_r2_prof.sh $TAG 2>&1 | tail -9
