"""Grid searches over (n, f) (reference: simcore/policy_paper.py:7-77).

All inputs are static per (DC, job type[, hour]) so the host evaluates these once and ships the winners to
the device as tables (see spec.py); strict ``<`` keeps the first minimum, n outer / f inner, as the reference.
"""
from typing import Iterable, Tuple

from .coeffs import TrainLatencyCoeffs, TrainPowerCoeffs

J_PER_KWH = 3.6e6


def energy_tuple(n: int, f: float, p_coeffs: TrainPowerCoeffs, t_coeffs: TrainLatencyCoeffs) -> Tuple[float, float, float]:
    seconds = t_coeffs.seconds_per_unit(n, f)
    watts = p_coeffs.task_watts(n, f)
    return (seconds, watts, watts * seconds)


def best_energy_freq(n: int, freq_levels: Iterable[float], p_coeffs: TrainPowerCoeffs,
                     t_coeffs: TrainLatencyCoeffs) -> float:
    levels = list(freq_levels)
    winner, least = None, float("inf")
    for f in levels:
        joules = energy_tuple(n, f, p_coeffs, t_coeffs)[2]
        if joules < least:
            least, winner = joules, f
    return winner if winner is not None else max(levels)


def keep_perf_when_expand(n0: int, f0: float, n1: int, t_coeffs: TrainLatencyCoeffs,
                          freq_levels: Iterable[float]) -> float:
    target = t_coeffs.seconds_per_unit(n0, max(1e-9, f0))
    denom = target - t_coeffs.alpha_t - t_coeffs.gamma_t * max(1, int(n1))
    if denom <= 1e-12:
        return f0
    want = t_coeffs.beta_t / denom
    return min(list(freq_levels), key=lambda lv: abs(lv - want))


def _objective_score(objective: str, joules: float, carbon_intensity: float, price_kwh: float) -> float:
    if objective == "carbon":
        return joules * carbon_intensity
    if objective == "cost":
        return (joules / J_PER_KWH) * float(price_kwh)
    return joules


def best_nf_grid(n_max: int, freq_levels, p_coeffs: TrainPowerCoeffs, t_coeffs: TrainLatencyCoeffs,
                 objective: str = "energy", carbon_intensity: float = 0.0, price_kwh: float = 0.0,
                 deadline_s=None):
    """Returns (n*, f*, T*, P*, E*)."""
    levels = list(freq_levels)
    winner = None
    for n in range(1, max(1, int(n_max)) + 1):
        for f in levels:
            seconds, watts, joules = energy_tuple(n, f, p_coeffs, t_coeffs)
            if deadline_s is not None and seconds > deadline_s:
                continue
            score = _objective_score(objective, joules, carbon_intensity, price_kwh)
            if winner is None or score < winner[0]:
                winner = (score, n, f, seconds, watts, joules)
    if winner is None:
        f_top = max(levels)
        seconds = t_coeffs.seconds_per_unit(1, f_top)
        watts = p_coeffs.gpu_watts(f_top)
        return 1, f_top, seconds, watts, watts * seconds
    return winner[1:]
