#!/bin/sh
TAG=s24
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2_smoke_$TAG.log 2>&1 || { echo "SMOKE FAILED"; tail -20 gpurun_out/r2_smoke_$TAG.log; exit 1; }
tail -1 gpurun_out/r2_smoke_$TAG.log
(timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "matches_oracle or resume or random_scenarios or trace_and_logs or full_size or reference_fixture or bench_batch or partially") > gpurun_out/r2_pytest_gpu_$TAG.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu_$TAG.log
timeout 1500 python tools/ab2.py --rounds 2 --cases cfg3,cfg2,joint_nf p23:lib=variants/libdcsim_p23.so cur > gpurun_out/r2_ab_$TAG.jsonl 2> gpurun_out/r2_ab_$TAG.err
python - <<PY
import json
for l in open("gpurun_out/r2_ab_$TAG.jsonl"):
    d = json.loads(l)
    print(d.get("case"), d.get("variant"), d.get("round"), "pre %.1f adv %.1f" % (d.get("prepass_ms", -1), d.get("advance_ms", -1)),
          "total %.3f Gev/s" % (d.get("gev_s", -1)), "warps", d.get("warps_per_sm"), "regs", d.get("regs"), "mode", d.get("mode"), "failed", d.get("failed"), d.get("error", "")[-300:])
PY
tail -5 gpurun_out/r2_ab_$TAG.err
