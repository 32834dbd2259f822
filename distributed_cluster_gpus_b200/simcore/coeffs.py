"""Per-(DC, job type) model coefficients.

Same public names and field order as the reference (simcore/coeffs.py:4-17) so existing
``build_paper_coeffs`` style tables keep working.  The evaluation methods carry the exact
floating-point evaluation order of energy_paper.py:4-12 and latency_paper.py:4-9; they are what the
host uses to fill policy tables, and what csrc/dcsim_kernel.cu re-evaluates per job start on the device.
"""
from dataclasses import dataclass, astuple


@dataclass
class TrainPowerCoeffs:
    alpha_p: float
    beta_p: float
    gamma_p: float

    def gpu_watts(self, f_ghz: float) -> float:
        # ((alpha * f**3) + (beta * f)) + gamma ; f**3 is libm pow (energy_paper.py:6)
        f = max(0.0, float(f_ghz))
        return self.alpha_p * (f ** 3) + self.beta_p * f + self.gamma_p

    def task_watts(self, n_gpus, f_ghz: float) -> float:
        return max(0, int(n_gpus)) * self.gpu_watts(f_ghz)

    def as_tuple(self):
        return astuple(self)


@dataclass
class TrainLatencyCoeffs:
    alpha_t: float
    beta_t: float
    gamma_t: float

    def seconds_per_unit(self, n_gpus, f_ghz: float) -> float:
        n = max(1, int(n_gpus))
        f = max(1e-9, float(f_ghz))
        base = self.alpha_t + self.beta_t / f
        if n == 1:
            return base
        return (base + self.gamma_t * n) / n

    def as_tuple(self):
        return astuple(self)
