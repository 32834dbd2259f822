"""ArrivalConfig, sample_job_size, expovariate_safe — re-exported from simcore/_surface.py (reference: simcore/arrivals.py)."""
from ._surface import (LOGNORM_FLOOR, LOGNORM_MEDIAN, LOGNORM_SIGMA, MODES, PARETO_ALPHA, PARETO_XM,  # noqa: F401
                       UNIFORM_FLOOR, ArrivalConfig, expovariate_safe, sample_job_size)
