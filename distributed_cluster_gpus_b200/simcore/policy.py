"""PolicyConfig, select_gpus_and_set_freq — re-exported from simcore/_surface.py (reference: simcore/policy.py)."""
from ._surface import POLICY_NAMES, PolicyConfig, select_gpus_and_set_freq  # noqa: F401
