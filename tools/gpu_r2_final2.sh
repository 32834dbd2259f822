#!/bin/sh
# last validation on HEAD: full GPU suite and both bench arms
TAG=${1:-f2}
(timeout 1800 python -m pytest tests -m gpu -x -q) > gpurun_out/r2_pytest_gpu_$TAG.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu_$TAG.log
(timeout 900 python bench.py --impl reference --steps 3 --warmup 3) > gpurun_out/r2_bench_${TAG}_reference.json 2> gpurun_out/r2_bench_${TAG}_reference.err; echo ref-arm rc $?
(timeout 1500 python bench.py --steps 3 --warmup 3) > gpurun_out/r2_bench_$TAG.json 2> gpurun_out/r2_bench_$TAG.err; echo bench rc $?; tail -c 300 gpurun_out/r2_bench_$TAG.err
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench_$TAG.json").read().strip().splitlines()[-1])
r = json.loads(open("gpurun_out/r2_bench_${TAG}_reference.json").read().strip().splitlines()[-1])
print("value %.4e e2e %.4e (%.1f ms/step, cold %.0f ms) roofline %.3f / step %.3f cpu %.3e ref-arm %.3e on %d threads; stdout lines: %d" % (d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["cold_first_run_ms"], d["roofline"]["frac"], d["roofline"]["frac_step"], d["cpu_baseline"]["value"], r["value"], r["cpu_baseline"]["cores"], len(open("gpurun_out/r2_bench_$TAG.json").read().strip().splitlines())))
print("kernel ms", d["roofline"]["kernel_ms"], "prepass ms", d["roofline"]["arrivals_prepass_kernel_ms"], "clocks", d.get("clocks"), "launches", d.get("gpu_launches"))
for k, v in (d.get("configs") or {}).items():
    print(k, {kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("events_per_s", "ms", "warps_per_sm", "failed_replicas", "replicas_total", "error")})
PY
