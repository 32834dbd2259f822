"""Start-up sanity lint on GPUType power figures (the reference's validate_gpus, simcore/validators.py:5-46)."""
from typing import Iterable, List

from .models import GPUType


def validate_gpus(gpus: Iterable[GPUType], strict: bool = False) -> List[str]:
    notes = []
    for g in gpus:
        if not (0 <= g.p_sleep <= g.p_idle):
            notes.append(f"{g.name}: expected 0 <= p_sleep ({g.p_sleep}) <= p_idle ({g.p_idle})")
        if g.p_peak <= 0:
            notes.append(f"{g.name}: p_peak must be positive")
        if g.alpha <= 0:
            notes.append(f"{g.name}: alpha must be positive")
        if g.tdp is not None and g.p_idle + g.p_peak > 1.25 * g.tdp:
            notes.append(f"{g.name}: p_idle + p_peak = {g.p_idle + g.p_peak} W exceeds TDP {g.tdp} W by >25%")
    if strict and notes:
        raise ValueError("; ".join(notes))
    return notes
