#!/usr/bin/env python
"""Device time of the pre-pass and the event loop per word source (bench workload, 65 536 replicas)."""
import json
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from distributed_cluster_gpus_b200 import scenarios as SC
from distributed_cluster_gpus_b200.engine import BatchedEngine

st = torch.cuda.Stream()
out = {}
with BatchedEngine(SC.to_spec(SC.CFG3), 65536, 123, 0, 0) as e:
    e.set_stream(st.cuda_stream)
    for kind in ("philox", "mt19937", "philox", "mt19937"):
        e.reset(123, 0)
        e.set_rng(kind)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        torch.cuda.synchronize(); ev[0].record(st)
        e.prepare(); ev[1].record(st)
        n = e.advance(0); ev[2].record(st); torch.cuda.synchronize()
        out.setdefault(kind, []).append({"prepass_ms": ev[0].elapsed_time(ev[1]), "advance_ms": ev[1].elapsed_time(ev[2]), "events": int(n)})
print(json.dumps(out))
