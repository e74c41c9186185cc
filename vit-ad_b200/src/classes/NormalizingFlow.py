"""src/classes/NormalizingFlow.py of the reference → vitad.nf."""
from vitad.nf import NormalizingFlow, NormalizingFlowReturn  # noqa: F401

__all__ = ["NormalizingFlow", "NormalizingFlowReturn"]
