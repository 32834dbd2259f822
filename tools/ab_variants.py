#!/usr/bin/env python
"""Same-box A/B of prebuilt library variants (DCSIM_B200_LIB): device-timed events/s of a few scenarios.

    python tools/ab_variants.py variants/libdcsim_base.so distributed_cluster_gpus_b200/csrc/libdcsim_b200.so
Each (scenario, library) pair runs in its own process, interleaved A B A B, kernel time from CUDA events.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, json, os
sys.path.insert(0, %r)
import torch
from distributed_cluster_gpus_b200 import scenarios as SC, spec as S
from distributed_cluster_gpus_b200.engine import BatchedEngine
name, reps, dur = sys.argv[1], int(sys.argv[2]), float(sys.argv[3])
sc = dict(SC.BY_NAME[name]) if name in SC.BY_NAME else None
if sc is None:
    algo = name.split(":")[1]
    sc = SC.scenario(name, 4, 64, SC.SIN10, SC.POI(1.0), dur, SC.FREQ3, algo=algo)
sc = dict(sc, duration=dur)
st = torch.cuda.Stream()
with BatchedEngine(SC.to_spec(sc), reps, 123, 0, 0) as e:
    e.set_stream(st.cuda_stream)
    best = None
    for it in range(3):
        e.reset(123 + it, 0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record(st)
        e.prepare(); n = e.advance(0)
        b.record(st); torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        best = ms if best is None or ms < best else best
    info = e.launch_info()
    print(json.dumps({"events": int(n), "ms": best, "gev_s": n / best / 1e6, "warps_per_sm": info["resident_warps_per_sm"],
                      "smem_per_cta": info["smem_bytes_per_cta"], "wpc": info["warps_per_cta"]}))
''' % ROOT

CASES = [("cfg3_4x64_sinusoid_120s", 65536, 120.0), ("sweep:joint_nf", 32768, 120.0), ("sweep:carbon_cost", 32768, 120.0),
         ("sweep:eco_route", 32768, 120.0), ("cfg5_8x256_sinusoid_60s", 16384, 60.0)]

libs = sys.argv[1:]
out = []
for name, reps, dur in CASES:
    for rnd in range(2):
        for lib in libs:
            env = dict(os.environ, DCSIM_B200_LIB=os.path.abspath(lib))
            r = subprocess.run([sys.executable, "-c", CHILD, name, str(reps), str(dur)], env=env, capture_output=True, text=True, timeout=600)
            line = r.stdout.strip().splitlines()[-1] if r.returncode == 0 and r.stdout.strip() else None
            row = {"scenario": name, "lib": os.path.basename(lib), "round": rnd}
            row.update(json.loads(line) if line else {"error": (r.stderr or "")[-400:]})
            out.append(row)
            print(json.dumps(row), flush=True)
