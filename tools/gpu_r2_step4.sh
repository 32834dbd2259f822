#!/bin/sh
# round 2, step 4: lanes-per-replica variants — parity subsets at 8 and 16 lanes, racecheck at 8, A/B against 32 lanes and round 1
TAG=s4
for G in 8 16; do
  (DCSIM_GROUP=$G timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "matches_oracle or resume or random_scenarios or trace_and_logs or cfg5_batch or full_size or reference_fixture") > gpurun_out/r2_pytest_gpu_${TAG}_g$G.log 2>&1; echo "G=$G:"; tail -3 gpurun_out/r2_pytest_gpu_${TAG}_g$G.log
done
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
for G in 8 32; do
  (DCSIM_GROUP=$G timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python /tmp/san.py) > gpurun_out/r2_sanitizer_racecheck_${TAG}_g$G.log 2>&1; echo racecheck G=$G exit $?; tail -2 gpurun_out/r2_sanitizer_racecheck_${TAG}_g$G.log
done
(DCSIM_GROUP=8 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san.py) > gpurun_out/r2_sanitizer_memcheck_${TAG}_g8.log 2>&1; echo memcheck G=8 exit $?; tail -2 gpurun_out/r2_sanitizer_memcheck_${TAG}_g8.log
timeout 2400 python tools/ab2.py --rounds 2 --cases cfg3,cfg5s,joint_nf,cfg2 \
  r1:lib=variants/libdcsim_r1.so cur g16:DCSIM_GROUP=16 g8:DCSIM_GROUP=8 g8s:DCSIM_GROUP=8:DCSIM_RECORDS=shared > gpurun_out/r2_ab_$TAG.jsonl 2> gpurun_out/r2_ab_$TAG.err
python - <<PY
import json
for l in open("gpurun_out/r2_ab_$TAG.jsonl"):
    d = json.loads(l)
    print(d.get("case"), d.get("variant"), d.get("round"), "pre %.1f adv %.1f" % (d.get("prepass_ms", -1), d.get("advance_ms", -1)),
          "total %.3f Gev/s" % (d.get("gev_s", -1)), "warps", d.get("warps_per_sm"), "regs", d.get("regs"), "mode", d.get("mode"), "failed", d.get("failed"), d.get("error", "")[-300:])
PY
