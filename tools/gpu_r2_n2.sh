#!/bin/sh
# two GPUs: the sharded CLI test over NCCL, then the bench line under torchrun (weak scaling + strong-scaling side record)
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cli_sharded") > gpurun_out/r2_pytest_gpu_n2_cli.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu_n2_cli.log
(timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3) > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo bench rc $?
grep -i "nranks\|NVLS\|Init COMPLETE" gpurun_out/r2_bench_n2.err | head -6
python - <<PY
import json
lines = [l for l in open("gpurun_out/r2_bench_n2.json").read().strip().splitlines() if l.startswith("{")]
print(len(lines), "json line(s) on stdout")
d = json.loads(lines[-1])
print("N=2 value %.3e e2e %.3e step %.1f ms comm %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["comm"]))
print("strong", {k: d["strong_scaling"].get(k) for k in ("events_per_s", "ms", "replicas_per_gpu", "prepass_ms_rank0", "event_loop_ms_rank0", "error")})
for k, v in (d.get("configs") or {}).items():
    print(k, {kk: (round(vv, 3) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("events_per_s", "ms", "replicas_total", "replicas_per_gpu", "failed_replicas", "error", "matches_baseline_size")})
PY
