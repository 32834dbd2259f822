#!/usr/bin/env python
"""Device time of one batch (bench workload, 65 536 replicas) with the drop-in's options switched on one by one."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distributed_cluster_gpus_b200 import scenarios as SC
from distributed_cluster_gpus_b200.engine import BatchedEngine

st = torch.cuda.Stream()
out = {}
for name, hist, log in (("plain", 0, 0), ("histogram", 1, 0), ("job+cluster log of replica 0", 0, 1), ("both (what run() does)", 1, 1), ("plain again", 0, 0)):
    with BatchedEngine(SC.to_spec(SC.CFG3), 65536, 123, 0, 0) as e:
        e.set_stream(st.cuda_stream)
        if hist:
            e.enable_latency_histogram()
        best = None
        for it in range(3):
            e.reset(123 + it, 0)
            e.set_logging(0, 12000 if log else 0, 200 if log else 0)
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            torch.cuda.synchronize(); ev[0].record(st)
            e.prepare(); ev[1].record(st)
            e.advance(0); ev[2].record(st); torch.cuda.synchronize()
            ms = (ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]))
            best = ms if best is None or sum(ms) < sum(best) else best
        info = e.launch_info()
        out[name] = {"prepass_ms": round(best[0], 2), "advance_ms": round(best[1], 2), "smem_per_cta": info["smem_bytes_per_cta"], "warps_per_sm": info["resident_warps_per_sm"]}
        print(name, out[name], flush=True)
