import sys, time, logging, os
sys.path.insert(0, ".")
import numpy as np, torch
from distributed_cluster_gpus_b200 import scenarios as SC, spec as S, engine as E
from distributed_cluster_gpus_b200.configs import paper_config as pc
from distributed_cluster_gpus_b200.simcore.simulator_paper_multi import MultiIngressPaperSimulator
sc = SC.CFG3; R = 65536
def T(label, t0): 
    torch.cuda.synchronize(); print(f"{label:34s} {1000*(time.perf_counter()-t0):8.1f} ms"); return time.perf_counter()
for rep in range(2):
    print("--- rep", rep)
    t0 = time.perf_counter()
    kw = SC.build_inputs(sc); t0 = T("build_inputs", t0)
    sim = MultiIngressPaperSimulator(router_policy=pc.build_router_policy(), logger=logging.getLogger("x"), sim_duration=sc["duration"], log_interval=sc["log_interval"], log_path="/dev/shm/e2e", rng_seed=5, algo=sc["algo"], show_progress=False, replicas=R, **kw); t0 = T("ctor (flatten)", t0)
    sp = sim._flatten({}); t0 = T("flatten again", t0)
    eng = E.BatchedEngine(sp, R, 5); t0 = T("engine create (malloc+memset)", t0)
    eng.set_logging(0, 12000, 200); t0 = T("set_logging", t0)
    eng.advance(0); t0 = T("advance (2 kernels, sync)", t0)
    s = eng.summary(); t0 = T("summary D2H pageable 46MB", t0)
    j = eng.job_log(); c = eng.cluster_log(); t0 = T("fetch logs", t0)
    sim._write_csvs(j, c); t0 = T("write csvs", t0)
    eng.close(); t0 = T("engine close (free)", t0)
    t0 = time.perf_counter(); sim.run(); T("sim.run() total", t0)

# steady state of the drop-in call (parked engine reused): where the host time goes
import cProfile, pstats, io
for rep in range(2):
    sim = MultiIngressPaperSimulator(router_policy=pc.build_router_policy(), logger=logging.getLogger("x"), sim_duration=sc["duration"], log_interval=sc["log_interval"], log_path="/dev/shm/e2e", rng_seed=7 + rep, algo=sc["algo"], show_progress=False, replicas=R, **SC.build_inputs(sc))
    sim.run()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
sim = MultiIngressPaperSimulator(router_policy=pc.build_router_policy(), logger=logging.getLogger("x"), sim_duration=sc["duration"], log_interval=sc["log_interval"], log_path="/dev/shm/e2e", rng_seed=11, algo=sc["algo"], show_progress=False, replicas=R, **SC.build_inputs(sc))
sim.run()
pr.disable()
print("steady-state ctor+run(): %.1f ms" % (1000 * (time.perf_counter() - t0)))
out = io.StringIO(); pstats.Stats(pr, stream=out).sort_stats("cumulative").print_stats(22); print(out.getvalue()[:5000])
