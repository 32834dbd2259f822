#!/usr/bin/env python
"""Where the end-to-end run() time goes: the one-replica companion (logged replica) and the steady-state drop-in call."""
import cProfile
import io
import logging
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from distributed_cluster_gpus_b200 import scenarios as SC, engine as E  # noqa: E402
from distributed_cluster_gpus_b200.configs import paper_config as pc  # noqa: E402
from distributed_cluster_gpus_b200.simcore.simulator_paper_multi import MultiIngressPaperSimulator  # noqa: E402

sc = SC.CFG3
sp = SC.to_spec(sc)
for reps in (1, 1, 64):
    with E.BatchedEngine(sp, reps, 5) as eng:
        eng.set_logging(0, 12000, 200)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.prepare()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        n = eng.advance(0)
        t2 = time.perf_counter()
        print(f"{reps} logged replica(s): pre-pass {1000 * (t1 - t0):.1f} ms, event loop {1000 * (t2 - t1):.1f} ms, {n} events, launch {eng.launch_info()['lanes_per_replica']} lanes")
R = 65536
for rep in range(3):
    t0 = time.perf_counter()
    sim = MultiIngressPaperSimulator(router_policy=pc.build_router_policy(), logger=logging.getLogger("x"), sim_duration=sc["duration"],
                                     log_interval=sc["log_interval"], log_path="/dev/shm/e2e", rng_seed=7 + rep, algo=sc["algo"],
                                     show_progress=False, replicas=R, **SC.build_inputs(sc))
    sim.run()
    print(f"run() #{rep}: {1000 * (time.perf_counter() - t0):.1f} ms")
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
sim = MultiIngressPaperSimulator(router_policy=pc.build_router_policy(), logger=logging.getLogger("x"), sim_duration=sc["duration"],
                                 log_interval=sc["log_interval"], log_path="/dev/shm/e2e", rng_seed=11, algo=sc["algo"],
                                 show_progress=False, replicas=R, **SC.build_inputs(sc))
sim.run()
pr.disable()
print("steady-state ctor+run(): %.1f ms" % (1000 * (time.perf_counter() - t0)))
out = io.StringIO()
pstats.Stats(pr, stream=out).sort_stats("cumulative").print_stats(25)
print(out.getvalue()[:6000])
