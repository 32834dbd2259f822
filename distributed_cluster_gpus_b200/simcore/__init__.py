"""Host-side mirror of the reference's ``simcore`` package surface (dataclasses, leaves, simulator)."""
