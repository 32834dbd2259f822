"""Job / GPUType / DataCenter records (reference: simcore/models.py:5-106).

``DataCenter`` instances are the *result carrier* of a run, as in the reference: after
``MultiIngressPaperSimulator.run()`` the batched engine writes replica 0's final ``energy_joules``,
``util_gpu_time``, ``accumulated_job_unit``, ``busy_gpus`` and ``current_freq`` back into them.
"""
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Tuple


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
