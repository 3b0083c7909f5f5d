"""Shared helpers for the tests (test infrastructure)."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, 'tests', 'golden')


def load_golden(name):
    return dict(np.load(os.path.join(GOLD, name + '.npz')))


def rel_err(a, b):
    """max |a-b| / max(|b|) -- the '1e-4 relative fp32' measure of BASELINE.json (relative to the tensor scale)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def t32(x, device='cpu'):
    return torch.tensor(np.asarray(x), dtype=torch.float32, device=device)
