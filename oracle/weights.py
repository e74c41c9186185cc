"""Seeded synthetic weights / inputs shared by the oracle's tests: the generators live in the product package
(vit-ad_b200/vitad/synth_weights.py, no compute) so that bench.py's GPU arm and the tools do not import `oracle/`."""
import os
import sys

_PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "vit-ad_b200")
if _PKG not in sys.path:
    sys.path.append(_PKG)  # after the reference tree when oracle/make_golden.py put it first (both have a `src`)

from vitad.synth_weights import *  # noqa: F401,F403,E402
from vitad.synth_weights import _kaiming_uniform, _tn, _xavier_normal  # noqa: F401,E402
