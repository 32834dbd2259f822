#!/bin/sh
# round 2, step 1: GPU parity suite + A/B of the record placement against the round-1 library
(timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r2_pytest_gpu_s1.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu_s1.log
timeout 1500 python tools/ab2.py --rounds 2 --cases cfg3,cfg5s,joint_nf,carbon_cost,cfg2 \
  r1:lib=variants/libdcsim_r1.so cur shared:DCSIM_RECORDS=shared global:DCSIM_RECORDS=global > gpurun_out/r2_ab_s1.jsonl 2> gpurun_out/r2_ab_s1.err
python - <<'PY'
import json
for l in open("gpurun_out/r2_ab_s1.jsonl"):
    d = json.loads(l)
    print(d.get("case"), d.get("variant"), d.get("round"), "pre %.1f adv %.1f" % (d.get("prepass_ms", -1), d.get("advance_ms", -1)),
          "loop %.3f Gev/s" % (d.get("loop_gev_s", -1) / 1e3), "warps", d.get("warps_per_sm"), "mode", d.get("mode"), "failed", d.get("failed"), d.get("error", "")[-200:])
PY
