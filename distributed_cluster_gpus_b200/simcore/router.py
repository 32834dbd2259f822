"""RouterPolicy — re-exported from simcore/_surface.py (reference: simcore/router.py)."""
from ._surface import RouterPolicy  # noqa: F401
