"""Shared comparison helpers for parity tests."""
import numpy as np
import torch


def nerr(a, b):
    """max|a-b| / max|b|  (SURVEY.md §7 hard part 1: the well-conditioned parity metric)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    d = np.max(np.abs(a - b)) if a.size else 0.0
    s = np.max(np.abs(b)) if b.size else 1.0
    return float(d / max(s, 1e-30))


def assert_close(a, b, tol, name=''):
    if isinstance(a, torch.Tensor):
        a = a.detach().float().cpu().numpy()
    if isinstance(b, torch.Tensor):
        b = b.detach().float().cpu().numpy()
    assert a.shape == b.shape, f'{name}: shape {a.shape} vs {b.shape}'
    assert np.all(np.isfinite(a)), f'{name}: non-finite values'
    e = nerr(a, b)
    assert e <= tol, f'{name}: normalised max error {e:.3e} > {tol:.1e}'
    return e


def strided_sample(flat, n=257):
    idx = torch.linspace(0, flat.numel() - 1, steps=min(n, flat.numel())).long()
    return flat[idx]


def rel_l2(a, b):
    """||a-b||_2 / ||b||_2 -- robust to isolated ReLU-mask flips (bf16 vs fp32 comparisons)."""
    if isinstance(a, torch.Tensor):
        a = a.detach().float().cpu().numpy()
    if isinstance(b, torch.Tensor):
        b = b.detach().float().cpu().numpy()
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
