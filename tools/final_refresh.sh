#!/bin/sh
# final-build evidence in one call: racecheck, launch list, ncu full of the event loop, bench both arms, full parity, fuzz
cat > /tmp/san.py <<PY
import sys; sys.path.insert(0, ".")
from distributed_cluster_gpus_b200 import scenarios as SC
from distributed_cluster_gpus_b200.engine import BatchedEngine
for name, dur in (("cfg3_4x64_sinusoid_120s", 6.0), ("cap_greedy_4x64", 12.0), ("sweep_bandit", 6.0), ("cfg5_8x256_sinusoid_60s", 2.0), ("full_swing_sinusoid_amp1", 20.0)):
    sc = dict(SC.BY_NAME[name], duration=dur)
    for job_rows in (4096, 0):
        with BatchedEngine(SC.to_spec(sc), 40, 5) as e:
            e.set_logging(1, job_rows, 512); e.set_trace(2, 2048)
            while not e.all_done(): e.advance(700)
            print(name, job_rows, int(e.summary()[:,1].sum()), "events")
PY
(timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/san.py) > gpurun_out/r1_sanitizer_racecheck_v8.log 2>&1; echo racecheck exit $?; tail -1 gpurun_out/r1_sanitizer_racecheck_v8.log
(timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches_v8.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e) > gpurun_out/r1_launches_v8.log 2>&1; echo launch-list $?
(timeout 900 ncu --set full --clock-control none --import-source on -k regex:dcsim_advance -s 1 -c 1 -f -o gpurun_out/r1_prof_v8_advance python bench.py --steps 1 --warmup 1 --replicas 8192 --no-cpu-baseline --no-e2e) > gpurun_out/r1_ncu_full_v8a.log 2>&1; echo full-advance $?
timeout 600 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r1_bench_v8_reference.json 2> gpurun_out/r1_bench_v8_reference.err; echo ref-arm $?
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r1_bench_v8_full.json 2> gpurun_out/r1_bench_v8_full.err; echo bench $?
timeout 900 python tools/full_parity.py > gpurun_out/r1_full_parity_v8.json 2> gpurun_out/r1_full_parity_v8.err; echo full-parity $?; tail -c 330 gpurun_out/r1_full_parity_v8.json
timeout 600 python tools/fuzz_gpu.py --cases 3000 --seed 21 > gpurun_out/fuzz_gpu_v8.json 2> gpurun_out/fuzz_gpu_v8.err; echo fuzz $?
python - <<'PY'
import json
d = json.load(open("gpurun_out/fuzz_gpu_v8.json")); print(d["cases"], "cases", d["failures"], "failures, worst", d["worst_float_rel_err"], "ill-conditioned", len(d["ill_conditioned"]))
d = json.loads(open("gpurun_out/r1_bench_v8_full.json").read().strip().splitlines()[-1])
print("value", d["value"] / 1e9, "e2e", d["e2e"]["value"] / 1e9, "cpu", d["cpu_baseline"]["value"] / 1e7)
PY
