"""Stand-in for FrEIA==0.2 (requirements.txt:48 of the reference; not vendored, not installable here).

TEST INFRASTRUCTURE ONLY.  Restates SequenceINN and AllInOneBlock as used by
src/classes/NormalizingFlow.py:95,104-114,127 (forward direction, unconditional, hard permutation,
SOFTPLUS global affine).  Second in-tree statement of the same wiring: fastflow_gathierry.py:50-75.
"""
from . import framework, modules  # noqa: F401
