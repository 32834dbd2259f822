#!/bin/sh
# round 2, step 6: candidate final build — full parity suite, A/B (auto vs 32 lanes vs round 1), profiles, both bench arms
TAG=s7
(timeout 1800 python -m pytest tests -m gpu -x -q) > gpurun_out/r2_pytest_gpu_$TAG.log 2>&1; tail -4 gpurun_out/r2_pytest_gpu_$TAG.log
timeout 2400 python tools/ab2.py --rounds 2 --cases cfg3,cfg2,cfg5s,joint_nf \
  r1:lib=variants/libdcsim_r1.so g32:DCSIM_GROUP=32 auto > gpurun_out/r2_ab_$TAG.jsonl 2> gpurun_out/r2_ab_$TAG.err
python - <<PY
import json
for l in open("gpurun_out/r2_ab_$TAG.jsonl"):
    d = json.loads(l)
    print(d.get("case"), d.get("variant"), d.get("round"), "pre %.1f adv %.1f" % (d.get("prepass_ms", -1), d.get("advance_ms", -1)),
          "total %.3f Gev/s" % (d.get("gev_s", -1)), "warps", d.get("warps_per_sm"), "regs", d.get("regs"), "mode", d.get("mode"), "failed", d.get("failed"), d.get("error", "")[-300:])
PY
sh tools/gpu_r2_prof.sh $TAG 2>&1 | tail -9
(timeout 900 python bench.py --impl reference --steps 3 --warmup 3) > gpurun_out/r2_bench_${TAG}_reference.json 2> gpurun_out/r2_bench_${TAG}_reference.err; echo ref-arm rc $?
(timeout 1500 python bench.py --steps 3 --warmup 3) > gpurun_out/r2_bench_$TAG.json 2> gpurun_out/r2_bench_$TAG.err; echo bench rc $?; tail -c 400 gpurun_out/r2_bench_$TAG.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench_$TAG.json").read().strip().splitlines()[-1])
r = json.loads(open("gpurun_out/r2_bench_${TAG}_reference.json").read().strip().splitlines()[-1])
print("value %.3e e2e %.3e (%.1f ms/step, cold %.0f ms) roofline %.3f / step %.3f cpu %.3e ref-arm %.3e on %d threads" % (d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["cold_first_run_ms"], d["roofline"]["frac"], d["roofline"]["frac_step"], d["cpu_baseline"]["value"], r["value"], r["cpu_baseline"]["cores"]))
print("kernel ms", d["roofline"]["kernel_ms"], "prepass ms", d["roofline"]["arrivals_prepass_kernel_ms"], "launch", d["config"]["launch"])
for k, v in (d.get("configs") or {}).items():
    print(k, {kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("events_per_s", "ms", "warps_per_sm", "failed_replicas", "replicas_total", "error", "staging_mode")})
PY
