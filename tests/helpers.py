"""Shared helpers for the parity tests (test infrastructure)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def golden_state_dict(fix, prefix="sd/"):
    sd = {}
    for k, v in fix.items():
        if k.startswith(prefix):
            t = torch.from_numpy(np.array(v))
            sd[k[len(prefix):]] = t
    return sd


def rel_err(a, b, floor=0.0):
    """max |a-b| / max(|b|) - the 'relative' tolerance of BASELINE.json's north_star.  `floor` bounds the
    denominator from below for tensors that are analytically zero (e.g. the gradient of a conv bias that
    feeds a training-mode BatchNorm is rounding noise of order 1e-9 in the reference itself)."""
    a, b = a.detach().double().flatten(), b.detach().double().flatten()
    denom = max(b.abs().max().item(), floor)
    return (a - b).abs().max().item() / (denom if denom > 0 else 1.0)


def max_abs(a, b):
    return (a.double() - b.double()).abs().max().item()


def rms_rel_err(a, b):
    """||a-b||_2 / ||b||_2"""
    a, b = a.detach().double().flatten(), b.detach().double().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
