#!/bin/sh
# round 2, step 3: parity suite (32 lanes per replica, then 8 and 16), A/B against round 1, launch list + ncu, first full bench line
TAG=s3
(timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r2_pytest_gpu_$TAG.log 2>&1; tail -5 gpurun_out/r2_pytest_gpu_$TAG.log
for G in 8 16; do
  (DCSIM_GROUP=$G timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "matches_oracle or resume or random_scenarios or trace_and_logs or cfg5_batch or full_size or reference_fixture") > gpurun_out/r2_pytest_gpu_${TAG}_g$G.log 2>&1; echo "G=$G:"; tail -3 gpurun_out/r2_pytest_gpu_${TAG}_g$G.log
done
timeout 2400 python tools/ab2.py --rounds 2 --cases cfg3,cfg5s,joint_nf,cfg2 \
  r1:lib=variants/libdcsim_r1.so cur g16:DCSIM_GROUP=16 g8:DCSIM_GROUP=8 g8s:DCSIM_GROUP=8:DCSIM_RECORDS=shared > gpurun_out/r2_ab_$TAG.jsonl 2> gpurun_out/r2_ab_$TAG.err
python - <<PY
import json
for l in open("gpurun_out/r2_ab_$TAG.jsonl"):
    d = json.loads(l)
    print(d.get("case"), d.get("variant"), d.get("round"), "pre %.1f adv %.1f" % (d.get("prepass_ms", -1), d.get("advance_ms", -1)),
          "total %.3f Gev/s" % (d.get("gev_s", -1)), "warps", d.get("warps_per_sm"), "regs", d.get("regs"), "mode", d.get("mode"), "failed", d.get("failed"), d.get("error", "")[-300:])
PY
sh tools/gpu_r2_prof.sh $TAG 2>&1 | tail -12
(timeout 1500 python bench.py --steps 2 --warmup 3) > gpurun_out/r2_bench_$TAG.json 2> gpurun_out/r2_bench_$TAG.err; echo bench rc $?; tail -c 600 gpurun_out/r2_bench_$TAG.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench_$TAG.json").read().strip().splitlines()[-1])
print("value %.3e e2e %.3e cold %.0f ms roofline %.3f / step %.3f cpu %.3e" % (d["value"], d["e2e"]["value"], d["e2e"]["cold_first_run_ms"], d["roofline"]["frac"], d["roofline"]["frac_step"], d["cpu_baseline"]["value"]))
for k, v in (d.get("configs") or {}).items():
    print(k, {kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("events_per_s", "ms", "warps_per_sm", "failed_replicas", "replicas_total", "error", "staging_mode")})
PY
