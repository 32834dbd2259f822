#!/bin/sh
# One GPU call: sanitizers on small batches (both kernel modes), the GPU test-suite, a quick bench line.
cat > /tmp/san.py <<PY
import sys; sys.path.insert(0, ".")
from distributed_cluster_gpus_b200 import scenarios as SC
from distributed_cluster_gpus_b200.engine import BatchedEngine
for name, dur in (("cfg3_4x64_sinusoid_120s", 6.0), ("cap_greedy_4x64", 12.0), ("sweep_bandit", 6.0), ("sweep_eco_route", 6.0), ("cfg5_8x256_sinusoid_60s", 2.0), ("full_swing_sinusoid_amp1", 20.0)):
    sc = dict(SC.BY_NAME[name], duration=dur)
    for job_rows in (4096, 0):  # with the job log (full running records) and without (lean records)
        with BatchedEngine(SC.to_spec(sc), 40, 5) as e:
            e.set_logging(1, job_rows, 512); e.set_trace(2, 2048)
            while not e.all_done(): e.advance(700)
            print(name, job_rows, int(e.summary()[:,1].sum()), "events")
PY
(timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/san.py) > gpurun_out/r1_sanitizer_racecheck_v6.log 2>&1; echo racecheck exit $?; tail -1 gpurun_out/r1_sanitizer_racecheck_v6.log
(DCSIM_PREPASS=0 timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/san.py) > gpurun_out/r1_sanitizer_racecheck_v6_legacy.log 2>&1; echo racecheck-legacy exit $?; tail -1 gpurun_out/r1_sanitizer_racecheck_v6_legacy.log
(timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san.py) > gpurun_out/r1_sanitizer_memcheck_v6.log 2>&1; echo memcheck exit $?; tail -1 gpurun_out/r1_sanitizer_memcheck_v6.log
(timeout 600 python -m pytest tests -m gpu -x -q) 2>&1 | tail -2
timeout 200 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('events/s', round(d['value']/1e9,4), 'advance ms', d['kernel_ms'], 'prepass ms', d['prepass_ms'])"
