"""Shared helpers for the tests: golden fixture loading and the gumbel-noise seeding convention used
by oracle/make_golden.py (k-th gumbel_softmax call of a run uses generator seed base+k)."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def gumbel(shape, seed):
    from oracle.vitad_oracle import gumbel_noise

    return gumbel_noise(shape, torch.Generator().manual_seed(seed))


def rel_err(a, b, floor=1e-3):
    """max |a-b| / max(|b|, floor): the parity metric of SURVEY.md §7 (defined at b == 0 through the floor)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))
