"""Job / PreemptedJob / GPUType / DataCenter — re-exported from simcore/_surface.py (reference: simcore/models.py)."""
from ._surface import DataCenter, GPUType, Job, PreemptedJob  # noqa: F401
