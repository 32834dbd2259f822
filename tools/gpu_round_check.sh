#!/bin/sh
# GPU tests, full bench line, 4000-case differential fuzz (one call)
(timeout 800 python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu_v6.log 2>&1; tail -2 gpurun_out/pytest_gpu_v6.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r1_bench_v6b.json 2> gpurun_out/r1_bench_v6b.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r1_bench_v6b.json").read().strip().splitlines()[-1])
print("value", d["value"] / 1e9, "e2e", d["e2e"]["value"] / 1e9, "e2e ms", d["e2e"]["ms_per_step"], "advance ms", d["kernel_ms"], "prepass ms", d["prepass_ms"])
PY
timeout 600 python tools/fuzz_gpu.py --cases 4000 --seed 11 > gpurun_out/fuzz_gpu_4000.json 2> gpurun_out/fuzz_gpu_4000.err; echo fuzz rc $?
python - <<'PY'
import json
d = json.load(open("gpurun_out/fuzz_gpu_4000.json"))
print(d["cases"], "cases", d["failures"], "failures, worst", d["worst_float_rel_err"], "events", d["events_compared"],
      "ill-conditioned:", [(i["case"], i["gpu_rel_err"], i["one_ulp_probe_rel_err"]) for i in d["ill_conditioned"]])
PY
