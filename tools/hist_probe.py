#!/usr/bin/env python
"""Event-loop time with and without the job-latency histogram, per lanes-per-replica build."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys
sys.path.insert(0, %r)
import torch
from distributed_cluster_gpus_b200 import scenarios as SC
from distributed_cluster_gpus_b200.engine import BatchedEngine
hist = sys.argv[1] == "1"
st = torch.cuda.Stream()
with BatchedEngine(SC.to_spec(SC.CFG3), 65536, 123) as e:
    e.set_stream(st.cuda_stream)
    if hist:
        e.enable_latency_histogram()
    for it in range(3):
        e.reset(123 + it, 0)
        a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        torch.cuda.synchronize(); a.record(st); e.prepare(); b.record(st); e.advance(0, sync=False); c.record(st); torch.cuda.synchronize()
    print("hist" if hist else "plain", "lanes", e.launch_info()["lanes_per_replica"], "pre %%.1f adv %%.1f ms" %% (a.elapsed_time(b), b.elapsed_time(c)))
''' % ROOT
for g in ("8", "32"):
    for h in ("0", "1"):
        r = subprocess.run([sys.executable, "-c", CHILD, h], env=dict(os.environ, DCSIM_GROUP=g), capture_output=True, text=True)
        print(r.stdout.strip() or r.stderr[-500:])
