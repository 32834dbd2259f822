#!/bin/sh
# pre-pass variants on the same box
timeout 1200 python tools/ab2.py --rounds 3 --cases cfg3,cfg5s cur noph2:lib=variants/libdcsim_noph2.so nolb:lib=variants/libdcsim_nolb.so > gpurun_out/r2_ab_s8.jsonl 2> gpurun_out/r2_ab_s8.err
python - <<PY
import json
for l in open("gpurun_out/r2_ab_s8.jsonl"):
    d = json.loads(l)
    print(d.get("case"), d.get("variant"), d.get("round"), "pre %.2f adv %.1f" % (d.get("prepass_ms", -1), d.get("advance_ms", -1)), d.get("error", "")[-200:])
PY
