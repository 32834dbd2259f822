"""Ingress, Edge, Graph — re-exported from simcore/_surface.py (reference: simcore/network.py)."""
from ._surface import Edge, Graph, Ingress  # noqa: F401
