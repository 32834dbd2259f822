#!/usr/bin/env python
"""Differential fuzz, CPU only: the device core compiled single-lane for the host (tests/hostemu) against the C
oracle on random scenarios — every algo, both policies, ragged GPU counts, 1-8 DCs, random frequency ladders, power
caps, log intervals, arrival modes; summaries, traces, job and cluster logs must be IDENTICAL (bit for bit).
Capacity overflows are not failures (the engine re-runs with raised capacities) but are counted.

    python tools/fuzz_core.py --cases 400 --seed 1          # a few minutes
"""
import argparse
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np  # noqa: E402

import hostemu_lib as hostemu  # noqa: E402
import oracle_lib as oracle  # noqa: E402
from distributed_cluster_gpus_b200 import scenarios as SC, spec as S  # noqa: E402
from distributed_cluster_gpus_b200.engine import CLUSTER_DTYPE, JOB_DTYPE  # noqa: E402

ALGOS = ["default_policy", "joint_nf", "carbon_cost", "eco_route", "debug", "bandit", "cap_uniform", "cap_greedy"]
HIGH_WATER = (S.S_MAX_XFER, S.S_MAX_RUN, S.S_MAX_Q)


def random_scenario(rnd, case):
    n_dc = rnd.choice([1, 1, 2, 2, 3, 4, 4, 5, 8])
    ragged = rnd.random() < 0.4
    gpus_list = [rnd.choice([1, 2, 3, 5, 8, 12, 16, 33, 64, 100]) for _ in range(n_dc)] if ragged else None
    gpus = None if ragged else rnd.choice([1, 2, 4, 8, 16, 64, 128])
    n_lv = rnd.choice([1, 2, 3, 5, 8])
    base = sorted(rnd.sample([0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9], n_lv - 1)) + [1.0] if n_lv > 1 else [1.0]
    def arrival(scale):
        mode = rnd.choice(["poisson", "poisson", "sinusoid", "sinusoid", "off"])
        return dict(mode=mode, rate=scale * 10 ** rnd.uniform(-1.2, 1.8), amp=rnd.choice([-1, 1]) * rnd.uniform(0.0, 1.0),
                    period=10 ** rnd.uniform(0.5, 3.6))
    algo = rnd.choice(ALGOS)
    sc = SC.scenario(f"fuzz{case}", n_dc, gpus, arrival(1.0), arrival(0.1), duration=rnd.uniform(0.5, 80.0),
                     freq_levels=base, algo=algo, policy=rnd.choice(["energy_aware", "perf_first"]),
                     log_interval=rnd.choice([0.5, 1.0, 5.0, 7.3, 1000.0]),
                     power_cap=rnd.choice([0.0, 0.0, 300.0, 2000.0, 20000.0]) if algo in ("cap_greedy", "cap_uniform", "eco_route") else 0.0,
                     num_fixed_gpus=rnd.choice([1, 2, 4, 8]), fixed_freq=rnd.choice([None, None] + base),
                     gpus_list=gpus_list)
    return sc


def same(a, b):
    a, b = a.copy(), b.copy()
    for col in HIGH_WATER:
        a[..., col] = b[..., col] = 0
    return np.array_equal(a, b)


def check(sc, seed, head_only, with_logs):
    """head_only: the host build's "head staged" mode (running-job records used at the block's home, DCSIM_RECORDS)."""
    os.environ["DCSIM_RECORDS"] = "global" if head_only else "shared"
    blob = SC.to_spec(sc).to_bytes()
    want, total = oracle.run_batch(blob, 2, seed, 0)
    kw = dict(rec_replica=1, trace_cap=4000)
    if with_logs:
        kw.update(job_dtype=JOB_DTYPE, jobs_cap=60000, cluster_dtype=CLUSTER_DTYPE, cluster_cap=4000)
    got = hostemu.run_batch(blob, 2, seed, chunk_events=0, **kw)
    st = got["summary"][:, S.S_STATUS]
    if np.any(st != 0):
        return "overflow:%d" % int(st.max())
    if got["events"] != total or not same(got["summary"], want):
        return "SUMMARY MISMATCH cols %s" % np.argwhere(got["summary"] != want)[:8].tolist()
    sim = oracle.OracleSim(blob, seed + 1, trace_cap=4000, joblog_cap=60000, clog_cap=4000)
    sim.advance(0)
    wt = sim.trace()
    for f in ("t", "seq", "kind"):
        if not np.array_equal(got["trace"][f], wt[f][:len(got["trace"])]) or len(got["trace"]) != len(wt):
            return "TRACE MISMATCH " + f
    if with_logs:
        wj, wc = sim.job_log(), sim.cluster_log()
        if len(wj) != len(got["jobs"]) or len(wc) != len(got["cluster"]):
            return "LOG LENGTH MISMATCH"
        for f in JOB_DTYPE.names:
            if not np.array_equal(got["jobs"][f], wj[f]):
                return "JOB LOG MISMATCH " + f
        for f in CLUSTER_DTYPE.names:
            if not np.array_equal(got["cluster"][f], wc[f]):
                return "CLUSTER LOG MISMATCH " + f
    sim.close()
    return "ok"


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", type=int, default=200)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    rnd = random.Random(args.seed)
    tally, bad, t0 = {}, [], time.time()
    for case in range(args.cases):
        sc = random_scenario(rnd, case)
        seed = rnd.randrange(1, 2 ** 40)
        for head_only in (False, True):
            for with_logs in (False, True):
                try:
                    res = check(sc, seed, head_only, with_logs)
                except Exception as e:  # spec rejected etc.
                    res = "EXC " + type(e).__name__ + ": " + str(e)[:120]
                key = res.split(" ")[0] if not res.startswith("overflow") else res
                tally[key] = tally.get(key, 0) + 1
                if res != "ok" and not res.startswith("overflow"):
                    bad.append((case, head_only, with_logs, res, sc, seed))
                    print("FAIL", case, "head" if head_only else "staged", "logs" if with_logs else "nolog", res, sc, seed, flush=True)
    print("events compared per oracle batch: see tally;", end=" ")
    print("cases", args.cases, "tally", tally, "failures", len(bad), "in %.0f s" % (time.time() - t0))
    sys.exit(1 if bad else 0)
