"""Discrete DVFS "atoms" per running task (reference: simcore/freq_load_agg.py:8-80).

Consumed only by the cap_greedy controller at log ticks (SIM:264-273).
"""
from dataclasses import dataclass
from typing import Iterable, List

from .coeffs import TrainLatencyCoeffs, TrainPowerCoeffs


@dataclass
class TaskState:
    job_id: int
    dc_name: str
    n: int
    f: float
    freq_levels: List[float]
    p_coeffs: TrainPowerCoeffs
    t_coeffs: TrainLatencyCoeffs


@dataclass
class Atom:
    rho: float      # dP / dV
    dV: float
    dP: float
    job_id: int
    dc_name: str
    f_from: float
    f_to: float


def _speed(task: TaskState, f: float) -> float:
    seconds = task.t_coeffs.seconds_per_unit(task.n, f)
    return 0.0 if seconds <= 0 else 1.0 / seconds


def _walk(task: TaskState, levels, indices):
    """Atoms along consecutive level indices (pairs k -> k')."""
    out = []
    k0 = indices[0]
    v_cur, p_cur = _speed(task, levels[k0]), task.p_coeffs.task_watts(task.n, levels[k0])
    for k_from, k_to in zip(indices, indices[1:]):
        v_new, p_new = _speed(task, levels[k_to]), task.p_coeffs.task_watts(task.n, levels[k_to])
        d_v, d_p = max(0.0, abs(v_new - v_cur) if (v_new - v_cur) * (k_to - k_from) > 0 else 0.0), \
            max(0.0, abs(p_new - p_cur) if (p_new - p_cur) * (k_to - k_from) > 0 else 0.0)
        if d_v > 0 and d_p >= 0:
            out.append(Atom(rho=d_p / d_v, dV=d_v, dP=d_p, job_id=task.job_id, dc_name=task.dc_name,
                            f_from=levels[k_from], f_to=levels[k_to]))
        v_cur, p_cur = v_new, p_new
    return out


def atoms_for_task(t: TaskState):
    levels = sorted(t.freq_levels)
    i0 = min(range(len(levels)), key=lambda i: abs(levels[i] - t.f))
    up = _walk(t, levels, list(range(i0, len(levels))))
    down = _walk(t, levels, list(range(i0, -1, -1)))
    return up, down


def aggregate_with_atoms(tasks: Iterable[TaskState]):
    ups, downs = [], []
    for task in tasks:
        u, d = atoms_for_task(task)
        ups.extend(u)
        downs.extend(d)
    ups.sort(key=lambda a: a.rho)
    downs.sort(key=lambda a: a.rho)
    return ups, downs
