#!/bin/sh
# round 2, step 2: merged event list + restructured pre-pass — GPU parity suite, A/B against the round-1 library, quick bench
(timeout 1200 python -m pytest tests -m gpu -x -q) > gpurun_out/r2_pytest_gpu_s2.log 2>&1; tail -5 gpurun_out/r2_pytest_gpu_s2.log
timeout 1500 python tools/ab2.py --rounds 2 --cases cfg3,cfg5s,joint_nf,cfg2 \
  r1:lib=variants/libdcsim_r1.so cur global:DCSIM_RECORDS=global > gpurun_out/r2_ab_s2.jsonl 2> gpurun_out/r2_ab_s2.err
python - <<'PY'
import json
for l in open("gpurun_out/r2_ab_s2.jsonl"):
    d = json.loads(l)
    print(d.get("case"), d.get("variant"), d.get("round"), "pre %.1f adv %.1f" % (d.get("prepass_ms", -1), d.get("advance_ms", -1)),
          "total %.3f Gev/s" % (d.get("gev_s", -1)), "warps", d.get("warps_per_sm"), "mode", d.get("mode"), "failed", d.get("failed"), d.get("error", "")[-300:])
PY
