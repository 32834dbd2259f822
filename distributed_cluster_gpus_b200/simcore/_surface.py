"""The reference's host-side surface in one place: every record type and leaf function a caller of
``MultiIngressPaperSimulator`` constructs or that ``spec.flatten`` consumes.

The reference spreads these over simcore/models.py, arrivals.py, policy.py, network.py and router.py; the modules of
the same names in this package re-export from here so ``from simcore.models import DataCenter`` style imports keep
working.  Field names, defaults, method semantics and error behaviour follow the reference (file:line cited per
section); the batched engine itself only reads these objects by attribute, so instances created by the reference's
own modules are accepted too.
"""
import heapq
import math
import random
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Tuple


# ====================================================================================================
# Jobs, GPU types, data centres  (reference: simcore/models.py:5-106)
#
# Job / GPUType / DataCenter records (reference: simcore/models.py:5-106).
#
# ``DataCenter`` instances are the *result carrier* of a run, as in the reference: after
# ``MultiIngressPaperSimulator.run()`` the batched engine writes replica 0's final ``energy_joules``,
# ``util_gpu_time``, ``accumulated_job_unit``, ``busy_gpus`` and ``current_freq`` back into them.
# ====================================================================================================

@dataclass
class GPUType:
    name: str
    p_idle: float           # W, clocked but idle
    p_peak: float           # W, dynamic part at f = 1.0
    p_sleep: float          # W, power-gated
    alpha: float = 3.0      # dynamic power ~ f**alpha
    tdp: Optional[float] = None


@dataclass
class Job:
    jid: int
    ingress: str
    jtype: str              # 'inference' | 'training'
    size: float
    arrival_time: float
    deadline: Optional[float] = None
    dc_name: Optional[str] = None
    gpus_assigned: int = 0
    start_time: Optional[float] = None
    finish_time: Optional[float] = None
    net_latency_s: float = 0.0
    f_used: float = 0.0
    units_total: float = 0.0
    units_done: float = 0.0
    last_update: float = 0.0
    ev_gen: int = 0
    preemptible: bool = False
    preempt_count: int = 0
    total_preempt_time: float = 0
    last_checkpoint: float = 0.0


@dataclass
class PreemptedJob:
    job: Job
    preempt_time: float
    reason: str
    preempt_ckpt: dict


@dataclass
class DataCenter:
    name: str
    gpu_type: GPUType
    total_gpus: int
    freq_levels: List[float]
    default_freq: float = 1.0
    power_gating: bool = True

    current_freq: float = field(init=False)
    busy_gpus: int = field(default=0, init=False)
    running_jobs: Dict[int, Tuple[Job, int]] = field(default_factory=dict, init=False)
    q_inf: List[Job] = field(default_factory=list, init=False)
    q_train: List[Job] = field(default_factory=list, init=False)
    energy_joules: float = field(default=0.0, init=False)
    last_energy_time: float = field(default=0.0, init=False)
    util_gpu_time: float = 0.0
    util_last_ts: float = 0.0
    util_begin_ts: float = 0.0
    accumulated_job_unit: float = 0.0
    preempted_jobs: List[PreemptedJob] = field(default_factory=list, init=False)
    preempt_policy: str = "fifo"

    def __post_init__(self):
        # same failure mode as the reference (models.py:75)
        assert self.default_freq in self.freq_levels, "default_freq must be one of freq_levels"
        self.current_freq = self.default_freq

    @property
    def free_gpus(self) -> int:
        return self.total_gpus - self.busy_gpus

    def idle_watts_per_gpu(self) -> float:
        return self.gpu_type.p_sleep if self.power_gating else self.gpu_type.p_idle

    def instantaneous_power_w(self) -> float:
        """DC-level fallback model used for the tail interval only (models.py:82-91, SIM:475)."""
        gt = self.gpu_type
        busy = self.busy_gpus
        dynamic = busy * (gt.p_idle + gt.p_peak * (self.current_freq ** gt.alpha))
        return dynamic + (self.total_gpus - busy) * self.idle_watts_per_gpu()

    def accrue_energy(self, now: float, power_fn: Optional[Callable[["DataCenter"], float]] = None) -> None:
        """E += P * dt with the first-touch sentinel of models.py:100-102."""
        if self.last_energy_time == 0.0:
            self.last_energy_time = now
            return
        dt = max(0.0, now - self.last_energy_time)
        watts = power_fn(self) if power_fn else self.instantaneous_power_w()
        self.energy_joules += watts * dt
        self.last_energy_time = now


# ====================================================================================================
# Arrival processes and job-size laws  (reference: simcore/arrivals.py:5-48)
#
# Arrival processes and job-size laws (reference: simcore/arrivals.py:5-48).
#
# Host-side definitions: the device kernel (csrc/dcsim_kernel.cu: sample_size / next_interarrival)
# implements the same laws on the Philox stream.  These functions draw from the ``random`` module,
# exactly like the reference, so host-only uses keep their meaning.
# ====================================================================================================

PARETO_XM = 1
PARETO_ALPHA = 1.8
LOGNORM_MEDIAN = 50000
LOGNORM_SIGMA = 0.4
LOGNORM_FLOOR = 0.1
UNIFORM_FLOOR = 1e-9

MODES = ("poisson", "sinusoid", "off")


def sample_job_size(jtype: str) -> float:
    if jtype == "inference":
        u = max(UNIFORM_FLOOR, 1 - random.random())
        return PARETO_XM / (u ** (1 / PARETO_ALPHA))
    return max(LOGNORM_FLOOR, random.lognormvariate(math.log(LOGNORM_MEDIAN), LOGNORM_SIGMA))


def expovariate_safe(lmbda: float) -> float:
    if lmbda <= 0:
        return float("inf")
    return random.expovariate(lmbda)


@dataclass
class ArrivalConfig:
    mode: str           # one of MODES
    rate: float         # per ingress, jobs / s
    amp: float = 0.0
    period: float = 3600.0

    def peak_rate(self) -> float:
        return self.rate * (1.0 + abs(self.amp))

    def lambda_t(self, t: float) -> float:
        if self.mode == "poisson":
            return self.rate
        if self.mode == "sinusoid":
            phase = 2 * math.pi * (t % self.period) / self.period
            return max(0.0, self.rate * (1.0 + self.amp * math.sin(phase)))
        if self.mode == "off":
            return 0.0
        raise ValueError("Unknown mode")

    def next_interarrival(self, t: float) -> float:
        if self.mode == "poisson":
            return expovariate_safe(self.rate)
        if self.mode == "sinusoid":
            # The reference's "thinning" keeps only the last candidate gap and never advances the clock
            # on a rejection (arrivals.py:41-45); reproduced as is.
            top = self.peak_rate()
            while True:
                gap = expovariate_safe(top)
                if random.random() <= self.lambda_t(t + gap) / top:
                    return gap
        if self.mode == "off":
            return float("inf")
        raise ValueError("Unknown mode")


# ====================================================================================================
# In-DC default policy  (reference: simcore/policy.py:5-41)
#
# In-DC default policy (reference: simcore/policy.py:5-41).
# ====================================================================================================

POLICY_NAMES = ("energy_aware", "perf_first")


@dataclass
class PolicyConfig:
    name: str
    max_gpus_per_job: int = 8
    inf_priority: bool = True
    dvfs_low: float = 0.6
    dvfs_high: float = 1.0
    train_scale_out_low_freq: bool = True
    reserve_inf_gpus: int = 0


def select_gpus_and_set_freq(dc: DataCenter, job: Job, policy: PolicyConfig) -> int:
    """Grab min(free, max_gpus_per_job) GPUs and *rewrite dc.current_freq* (policy.py:22-38)."""
    if policy.name not in POLICY_NAMES:
        raise ValueError("Unknown policy name")
    free = dc.free_gpus
    grab = max(1, min(free, policy.max_gpus_per_job) if free > 0 else 0)
    if job.jtype == "inference":
        dc.current_freq = policy.dvfs_high
    elif policy.name == "perf_first":
        dc.current_freq = max(dc.current_freq, policy.dvfs_high if len(dc.q_inf) > 0 else dc.default_freq)
    elif policy.train_scale_out_low_freq and free >= 2:
        dc.current_freq = policy.dvfs_low
    else:
        dc.current_freq = max(dc.current_freq, policy.dvfs_low)
    return grab


# ====================================================================================================
# WAN graph  (reference: simcore/network.py:7-62)
#
# WAN graph (reference: simcore/network.py:7-62).
#
# The graph is static, so the batched engine asks it once per (ingress, DC) pair while flattening the
# scenario (spec.py) instead of once per arrival as the reference does (SIM:487).
# ====================================================================================================

@dataclass(frozen=True)
class Ingress:
    name: str
    region: str


@dataclass
class Edge:
    to: str
    latency_ms: float
    capacity_gbps: float = math.inf
    cost_per_GB: float = 0.0


class Graph:
    """Directed graph; nodes are ingress / DC names."""

    def __init__(self):
        self.adj: Dict[str, List[Edge]] = {}

    def add_edge(self, u: str, v: str, latency_ms: float, capacity_gbps: float = math.inf,
                 cost_per_GB: float = 0.0):
        self.adj.setdefault(u, []).append(Edge(v, latency_ms, capacity_gbps, cost_per_GB))

    def shortest_path_latency(self, src: str, dst: str) -> Tuple[float, List[str], float, float]:
        """Dijkstra on latency -> (latency_s, path, bottleneck_Gbps or 0.0 if unbounded, sum cost/GB)."""
        best_ms: Dict[str, float] = {src: 0.0}
        came_from: Dict[str, Tuple[str, Edge]] = {}
        frontier: List[Tuple[float, str]] = [(0.0, src)]
        while frontier:
            d_ms, node = heapq.heappop(frontier)
            if node == dst:
                break
            if d_ms > best_ms.get(node, math.inf):
                continue
            for edge in self.adj.get(node, ()):
                cand = d_ms + edge.latency_ms
                if cand < best_ms.get(edge.to, math.inf):
                    best_ms[edge.to] = cand
                    came_from[edge.to] = (node, edge)
                    heapq.heappush(frontier, (cand, edge.to))
        if dst not in best_ms:
            return math.inf, [], 0.0, math.inf
        hops = [dst]
        narrowest = math.inf
        cost = 0.0
        node = dst
        while node != src:
            prev, edge = came_from[node]
            hops.append(prev)
            narrowest = min(narrowest, edge.capacity_gbps)
            cost += edge.cost_per_GB
            node = prev
        hops.reverse()
        return best_ms[dst] / 1000.0, hops, (0.0 if narrowest is math.inf else narrowest), cost


# ====================================================================================================
# RouterPolicy  (reference: simcore/router.py:3-9)
#
# RouterPolicy — accepted for API compatibility.
#
# The reference constructs it (configs/paper_config.py:289-291), stores it
# (simcore/simulator_paper_multi.py:65) and never reads it; placement is random.choice (:575-577),
# eco_route's arg-min (:544-553) or the RL actor.  Nothing to compute on either side.
# ====================================================================================================

@dataclass
class RouterPolicy:
    w_energy: float = 0.0
    w_latency: float = 1.0
    w_carbon: float = 0.0
    d_choices: int = 0

