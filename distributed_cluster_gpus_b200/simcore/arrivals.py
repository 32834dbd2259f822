"""Arrival processes and job-size laws (reference: simcore/arrivals.py:5-48).

Host-side definitions: the device kernel (csrc/dcsim_kernel.cu: sample_size / next_interarrival)
implements the same laws on the Philox stream.  These functions draw from the ``random`` module,
exactly like the reference, so host-only uses keep their meaning.
"""
import math
import random
from dataclasses import dataclass

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
