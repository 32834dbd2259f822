#!/bin/sh
# Same-box A/B of prebuilt libraries on the bench workload: tools/ab_bench.sh libA.so libB.so   (interleaved, 2 rounds)
for round in 1 2; do
  for v in "$@"; do
    DCSIM_B200_LIB=$PWD/$v timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e > /tmp/ab_line.json 2>/tmp/ab_err.log
    python - "$v" <<'PY'
import json, sys
try:
    d = json.loads(open("/tmp/ab_line.json").read().strip().splitlines()[-1])
    print(sys.argv[1][-22:], round(d["value"] / 1e9, 4), "advance ms", round(d["kernel_ms"], 2), "prepass ms", round(d["prepass_ms"], 2))
except Exception as e:
    print(sys.argv[1], "FAILED", e, open("/tmp/ab_err.log").read()[-600:])
PY
  done
done
