#!/bin/sh
# round 2, final validation of the committed build: full GPU suite, sanitizer, both bench arms, launch list, ncu captures
TAG=${1:-final}
(timeout 1800 python -m pytest tests -m gpu -x -q) > gpurun_out/r2_pytest_gpu_$TAG.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu_$TAG.log
cat > /tmp/san.py <<PY
import sys; sys.path.insert(0, ".")
from distributed_cluster_gpus_b200 import scenarios as SC
from distributed_cluster_gpus_b200.engine import BatchedEngine
for name, dur in (("cfg3_4x64_sinusoid_120s", 6.0), ("cap_greedy_4x64", 10.0), ("sweep_bandit", 6.0), ("cfg5_8x256_sinusoid_60s", 2.0), ("sweep_joint_nf", 5.0)):
    sc = dict(SC.BY_NAME[name], duration=dur)
    for job_rows in (4096, 0):
        with BatchedEngine(SC.to_spec(sc), 37, 5) as e:
            e.set_logging(1, job_rows, 512); e.set_trace(2, 2048)
            while not e.all_done(): e.advance(700)
            print(name, job_rows, int(e.summary()[:,1].sum()), "events", e.launch_info()["lanes_per_replica"])
PY
(timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/san.py) > gpurun_out/r2_sanitizer_racecheck_$TAG.log 2>&1; echo racecheck exit $?; tail -2 gpurun_out/r2_sanitizer_racecheck_$TAG.log
(timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san.py) > gpurun_out/r2_sanitizer_memcheck_$TAG.log 2>&1; echo memcheck exit $?; tail -2 gpurun_out/r2_sanitizer_memcheck_$TAG.log
(timeout 900 python bench.py --impl reference --steps 3 --warmup 3) > gpurun_out/r2_bench_${TAG}_reference.json 2> gpurun_out/r2_bench_${TAG}_reference.err; echo ref-arm rc $?
(timeout 1500 python bench.py --steps 3 --warmup 3) > gpurun_out/r2_bench_$TAG.json 2> gpurun_out/r2_bench_$TAG.err; echo bench rc $?; tail -c 300 gpurun_out/r2_bench_$TAG.err
(timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv \
   --log-file gpurun_out/r2_launches_${TAG}_cfg3_65536.csv python tools/prof_step.py cfg3_4x64_sinusoid_120s 65536 2) > gpurun_out/r2_launches_${TAG}.log 2>&1; echo launch-list $?
sh tools/gpu_r2_prof3.sh $TAG
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench_$TAG.json").read().strip().splitlines()[-1])
r = json.loads(open("gpurun_out/r2_bench_${TAG}_reference.json").read().strip().splitlines()[-1])
print("value %.3e e2e %.3e (%.1f ms/step, cold %.0f ms) roofline %.3f / step %.3f cpu %.3e ref-arm %.3e on %d threads" % (d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["cold_first_run_ms"], d["roofline"]["frac"], d["roofline"]["frac_step"], d["cpu_baseline"]["value"], r["value"], r["cpu_baseline"]["cores"]))
print("kernel ms", d["roofline"]["kernel_ms"], "prepass ms", d["roofline"]["arrivals_prepass_kernel_ms"], "launch", d["config"]["launch"])
for k, v in (d.get("configs") or {}).items():
    print(k, {kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("events_per_s", "ms", "warps_per_sm", "failed_replicas", "replicas_total", "error", "staging_mode")})
import csv
rows = list(csv.reader(open("gpurun_out/r2_launches_${TAG}_cfg3_65536.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
cols = rows[hdr]
ki, mi, vi = cols.index("Kernel Name"), cols.index("Metric Name"), cols.index("Metric Value")
for r in rows[hdr + 1:]:
    if len(r) > vi: print(r[0], r[ki][:40], r[mi], r[vi])
PY
