#!/usr/bin/env python
"""Re-runs one recorded fuzz case on the GPU: unchunked vs chunked stepping vs the oracle; prints where they differ."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import oracle_lib as oracle
from distributed_cluster_gpus_b200 import scenarios as SC, spec as S
from distributed_cluster_gpus_b200.engine import BatchedEngine

case = json.load(open(sys.argv[1]))
sc, seed = case["scenario"], case["seed"]
n = 6
sp = SC.to_spec(sc)
want, _ = oracle.run_batch(sp.to_bytes(), n, seed, 0, n_threads=6)
runs = {}
for chunk in (0, 997, 100, 5000):
    with BatchedEngine(sp, n, seed, 0, 0) as eng:
        if chunk == 0:
            eng.advance(0)
        else:
            while not eng.all_done():
                eng.advance(chunk)
        runs[chunk] = eng.summary()
def diff(a, b):
    rel = np.where(a == b, 0.0, np.abs(a - b) / np.maximum(np.abs(b), 1e-300))
    idx = np.argwhere(rel > 1e-12)
    return [(int(r), int(c), float(a[r, c]), float(b[r, c])) for r, c in idx[:12]]
for chunk, got in runs.items():
    print("chunk", chunk, "vs oracle:", diff(got, want))
    print("chunk", chunk, "vs unchunked GPU: identical =", bool(np.array_equal(got, runs[0])), diff(got, runs[0])[:6])

# first divergence of the event trace / job log of one replica
rep = int(sys.argv[2]) if len(sys.argv) > 2 else 2
with BatchedEngine(sp, n, seed, 0, 0) as eng:
    eng.set_logging(rep, 60000, 4000); eng.set_trace(rep, 40000)
    eng.advance(0)
    tr, jl, cl = eng.trace(), eng.job_log(), eng.cluster_log()
sim = oracle.OracleSim(sp.to_bytes(), seed + rep, trace_cap=40000, joblog_cap=60000, clog_cap=4000)
sim.advance(0)
wt, wj, wc = sim.trace(), sim.job_log(), sim.cluster_log()
print("trace lengths", len(tr), len(wt), "job log", len(jl), len(wj))
m = min(len(tr), len(wt))
bad = np.nonzero((tr["seq"][:m] != wt["seq"][:m]) | (tr["kind"][:m] != wt["kind"][:m]) |
                 (np.abs(tr["t"][:m] - wt["t"][:m]) > 1e-9 * np.maximum(1.0, np.abs(wt["t"][:m]))))[0]
print("first trace divergence at", bad[:5])
if len(bad):
    i = int(bad[0])
    for k in range(max(0, i - 4), min(m, i + 6)):
        print(k, "gpu", float(tr["t"][k]).hex(), int(tr["seq"][k]), int(tr["kind"][k]), "| oracle", float(wt["t"][k]).hex(), int(wt["seq"][k]), int(wt["kind"][k]))
mj = min(len(jl), len(wj))
for f in jl.dtype.names:
    a, b = jl[f][:mj], wj[f][:mj]
    d = np.nonzero(a != b)[0] if a.dtype.kind in "iu" else np.nonzero(np.abs(a - b) > 1e-9 * np.maximum(1.0, np.abs(b)))[0]
    if len(d):
        k = int(d[0])
        print("job log field", f, "first differs at row", k, "gpu", jl[k], "oracle", wj[k], "n differing", len(d))
mc = min(len(cl), len(wc))
for f in cl.dtype.names:
    a, b = cl[f][:mc], wc[f][:mc]
    d = np.nonzero(a != b)[0] if a.dtype.kind in "iu" else np.nonzero(np.abs(a - b) > 1e-9 * np.maximum(1.0, np.abs(b)))[0]
    if len(d):
        k = int(d[0])
        print("cluster log field", f, "first differs at row", k, "gpu", cl[k], "oracle", wc[k], "n differing", len(d))
