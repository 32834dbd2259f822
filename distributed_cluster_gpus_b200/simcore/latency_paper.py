"""T(n, f) seconds per work unit — function-style entry point of the reference (latency_paper.py:4-9)."""
from .coeffs import TrainLatencyCoeffs


def step_time_s(n_gpus: int, f_ghz: float, coeffs: TrainLatencyCoeffs) -> float:
    return coeffs.seconds_per_unit(n_gpus, f_ghz)
