"""Shared comparison helpers for the parity tests (oracle vs golden, CUDA path vs oracle/golden).

Why not a plain allclose everywhere: the path contains hard branches (LeakyReLU/ReLU at 0) and Adam's
first steps are `lr * g / (|g| + 1e-8) ~= lr * sign(g)`.  A pre-activation that is zero to within fp32
rounding can take the other branch in another (equally valid) fp32 summation order; that moves the
gradient entries it feeds by O(1e-3) relative and can flip the sign of a ~0 gradient, moving a weight by
2*lr.  The reference itself is subject to this, so trajectories are compared
tightly where the arithmetic is smooth (first forward, losses, probabilities) and with flip-tolerant
metrics on gradients and post-step weights.
"""
import numpy as np


def synthetic_real(seed, n, nc, size=224):
    """Uniform[-1,1) images; identical to oracle/make_golden.py::synthetic_real."""
    return (np.random.RandomState(seed).rand(n, nc, size, size).astype(np.float32) * 2 - 1)


def synthetic_noise(seed, n, nz):
    return np.random.RandomState(seed).randn(n, nz, 1, 1).astype(np.float32)


def close(a, b, rtol=1e-4, atol=1e-6, what=''):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    err = np.abs(a - b) - (atol + rtol * np.abs(b))
    assert (err <= 0).all(), f'{what}: max abs diff {np.abs(a - b).max():.3e}, ref max {np.abs(b).max():.3e}'


def grad_close(a, b, what='', bulk=2e-5, l2=2e-3, worst=1e-2):
    """inf-norm-normalised gradient comparison: median entry tight, relative L2 middle, worst entry loose."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    scale = max(np.abs(b).max(), 1e-12)
    err = np.abs(a - b).reshape(-1) / scale
    rel_l2 = np.sqrt(((a - b) ** 2).sum()) / max(np.sqrt((b ** 2).sum()), 1e-12)
    assert np.median(err) < bulk and rel_l2 < l2 and err.max() < worst, \
        f'{what}: median {np.median(err):.2e} relL2 {rel_l2:.2e} max {err.max():.2e} (scale {scale:.2e})'


def weights_close(a, b, what='', lr=2e-4, steps=1, rtol=1e-4, atol=2e-6, frac=0.98):
    """Post-Adam weights: at least `frac` of the entries within (rtol, atol); every entry within the
    2*lr-per-step envelope a gradient sign flip can cause."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    d = np.abs(a - b)
    ok = d <= atol + rtol * np.abs(b)
    assert ok.mean() >= frac, f'{what}: only {ok.mean():.4f} of entries within tolerance (max diff {d.max():.3e})'
    assert d.max() <= 2.05 * lr * steps + atol, f'{what}: max diff {d.max():.3e} exceeds the sign-flip envelope'
