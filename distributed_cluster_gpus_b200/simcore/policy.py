"""In-DC default policy (reference: simcore/policy.py:5-41)."""
from dataclasses import dataclass

from .models import DataCenter, Job

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
