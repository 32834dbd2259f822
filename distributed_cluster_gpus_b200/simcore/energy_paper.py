"""P_gpu(f) and P_job(n, f) — function-style entry points of the reference (energy_paper.py:4-12)."""
from .coeffs import TrainPowerCoeffs


def gpu_power_w(f_ghz: float, coeffs: TrainPowerCoeffs) -> float:
    return coeffs.gpu_watts(f_ghz)


def task_power_w(n_gpus: int, f: float, coeffs: TrainPowerCoeffs) -> float:
    return coeffs.task_watts(n_gpus, f)
