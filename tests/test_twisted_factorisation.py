"""The algebra behind csrc/adi_split.cu, checked in numpy (no GPU, no library).

The half-line kernels solve every tridiagonal system of a sweep,
    (A + eps I) x = d,   A = I - diag(r) L_N   (L_N: 1-D Neumann Laplacian; mnist_test.py:86-93,151-198),
by a TWISTED factorisation: cells 0..H-1 are eliminated top-down (the reference's Thomas pass),
cells N-1..H+1 bottom-up, cell H closes both.  These tests pin the recurrences the kernels use
(`stables_kernel`, `solve`, `reverse_core`) against numpy's dense solver, in fp64, including the
transposed solve of the adjoint and the identity behind the in-place rebuild of a sweep's input.
"""
import numpy as np
import pytest


def _system(N, rng, rmax=2.0, eps=1e-6):
    r = rng.uniform(0.01, rmax, N)
    b = 1.0 + 2.0 * r
    b[0], b[-1] = 1.0 + r[0], 1.0 + r[-1]
    T = np.diag(b + eps)
    for i in range(N):
        if i > 0:
            T[i, i - 1] = -r[i]
        if i < N - 1:
            T[i, i + 1] = -r[i]
    return r, b, T


def _twisted_tables(r, b, eps=1e-6):
    """inv = 1/pivot, e = r/pivot with the pivots of the twisted elimination (stables_kernel)."""
    N, H = len(r), len(r) // 2
    den = np.zeros(N)
    e = np.zeros(N)
    for i in range(H):                       # top-down
        den[i] = b[i] + eps - (r[i] * e[i - 1] if i > 0 else 0.0)
        e[i] = r[i] / den[i]
    for i in range(N - 1, H, -1):            # bottom-up
        den[i] = b[i] + eps - (r[i] * e[i + 1] if i < N - 1 else 0.0)
        e[i] = r[i] / den[i]
    den[H] = b[H] + eps - r[H] * e[H - 1] - r[H] * e[H + 1]    # the closing cell sees both neighbours
    e[H] = r[H] / den[H]
    return 1.0 / den, e


@pytest.mark.parametrize("N", [8, 12, 16, 28, 32])
def test_twisted_solve_matches_dense_solver(N):
    rng = np.random.default_rng(N)
    r, b, T = _system(N, rng)
    inv, e = _twisted_tables(r, b)
    H = N // 2
    d = rng.normal(size=N)
    ds = np.zeros(N)
    for i in range(H):
        ds[i] = inv[i] * d[i] + (e[i] * ds[i - 1] if i > 0 else 0.0)
    for i in range(N - 1, H, -1):
        ds[i] = inv[i] * d[i] + (e[i] * ds[i + 1] if i < N - 1 else 0.0)
    x = np.zeros(N)
    x[H] = inv[H] * d[H] + e[H] * (ds[H - 1] + ds[H + 1])       # one exchange between the two halves
    for i in range(H - 1, -1, -1):
        x[i] = ds[i] + e[i] * x[i + 1]
    for i in range(H + 1, N):
        x[i] = ds[i] + e[i] * x[i - 1]
    np.testing.assert_allclose(x, np.linalg.solve(T, d), rtol=0, atol=1e-13)


@pytest.mark.parametrize("N", [8, 28, 32])
def test_twisted_transposed_solve_and_rebuild(N):
    """The adjoint sweep solves T^T lambda = g through the same factors (with e = r * inv rebuilt
    from r and inv, as the backward kernel does), and x_in = T x_out rebuilds a sweep's input."""
    rng = np.random.default_rng(100 + N)
    r, b, T = _system(N, rng)
    inv, _ = _twisted_tables(r, b)
    e = r * inv
    H = N // 2
    g = rng.normal(size=N)
    w = np.zeros(N)
    for i in range(H):
        w[i] = g[i] + (e[i - 1] * w[i - 1] if i > 0 else 0.0)
    for i in range(N - 1, H, -1):
        w[i] = g[i] + (e[i + 1] * w[i + 1] if i < N - 1 else 0.0)
    w[H] = g[H] + e[H - 1] * w[H - 1] + e[H + 1] * w[H + 1]
    lam = np.zeros(N)
    lam[H] = inv[H] * w[H]
    for i in range(H - 1, -1, -1):
        lam[i] = inv[i] * (w[i] + r[i + 1] * lam[i + 1])
    for i in range(H + 1, N):
        lam[i] = inv[i] * (w[i] + r[i - 1] * lam[i - 1])
    np.testing.assert_allclose(lam, np.linalg.solve(T.T, g), rtol=0, atol=1e-13)
    # rebuild: (1 + eps) x - r (L x) == T x, with the Neumann ends of L
    x = rng.normal(size=N)
    Lx = np.empty(N)
    Lx[0], Lx[-1] = x[1] - x[0], x[-2] - x[-1]
    Lx[1:-1] = x[:-2] - 2.0 * x[1:-1] + x[2:]
    np.testing.assert_allclose((1.0 + 1e-6) * x - r * Lx, T @ x, rtol=0, atol=1e-13)
    # d(loss)/d r_i of one sweep = lambda_i (L x_out)_i  (what the kernels accumulate per pixel)
    d = T @ x
    i0, h = N // 3, 1e-6
    r2 = r.copy()
    r2[i0] += h
    b2 = 1.0 + 2.0 * r2
    b2[0], b2[-1] = 1.0 + r2[0], 1.0 + r2[-1]
    T2 = np.diag(b2 + 1e-6)
    for i in range(N):
        if i > 0:
            T2[i, i - 1] = -r2[i]
        if i < N - 1:
            T2[i, i + 1] = -r2[i]
    fd = (g @ np.linalg.solve(T2, d) - g @ x) / h
    assert abs(fd - lam[i0] * Lx[i0]) <= 1e-5 * max(1.0, abs(fd))
