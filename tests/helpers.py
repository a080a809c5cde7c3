"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, f'{name}.npz')) as z:
        return {k: z[k] for k in z.files}


def range_rel_err(a, ref):
    """max|a - ref| / max|ref| -- the parity metric used everywhere in this repo (per-pixel
    relative error is ill-conditioned at zero crossings, so errors are normalised by the reference's range)."""
    a = torch.as_tensor(a, dtype=torch.float64)
    ref = torch.as_tensor(ref, dtype=torch.float64)
    return ((a - ref).abs().max() / ref.abs().max().clamp_min(1e-30)).item()
