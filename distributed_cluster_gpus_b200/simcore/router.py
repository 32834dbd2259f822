"""RouterPolicy — accepted for API compatibility.

The reference constructs it (configs/paper_config.py:289-291), stores it
(simcore/simulator_paper_multi.py:65) and never reads it; placement is random.choice (:575-577),
eco_route's arg-min (:544-553) or the RL actor.  Nothing to compute on either side.
"""
from dataclasses import dataclass


@dataclass
class RouterPolicy:
    w_energy: float = 0.0
    w_latency: float = 1.0
    w_carbon: float = 0.0
    d_choices: int = 0
