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


def pixel_rel_err(a, ref, floor=1e-3):
    """north_star's literal metric: max over elements of |a - ref| / max(|ref|, floor * max|ref|) -- per-pixel relative error with an
    absolute floor so that zero crossings do not divide by ~0.  Returned together with the fraction of elements above 1e-3."""
    a = torch.as_tensor(a, dtype=torch.float64)
    ref = torch.as_tensor(ref, dtype=torch.float64)
    den = ref.abs().clamp_min(floor * ref.abs().max().clamp_min(1e-30))
    e = (a - ref).abs() / den
    return e.max().item(), (e > 1e-3).double().mean().item()


def report_parity(name, a, ref):
    """Both parity metrics of one output; appended to gpurun_out/parity_report.jsonl when that directory exists (so a GPU run
    leaves the numbers behind for profiles/).  -> (range_rel, pixel_rel_max, pixel_rel_frac_above_1e-3)."""
    import json
    r = range_rel_err(a, ref)
    pmax, pfrac = pixel_rel_err(a, ref)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'gpurun_out')
    if os.path.isdir(out):
        with open(os.path.join(out, 'parity_report.jsonl'), 'a') as f:
            f.write(json.dumps({'case': name, 'range_rel_err': r, 'pixel_rel_err_max_floor1e-3': pmax, 'pixel_frac_above_1e-3': pfrac}) + '\n')
    return r, pmax, pfrac
