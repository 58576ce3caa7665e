"""GPU: the persistent tiled Cholesky solve (csrc/pcs_chol.cu) against numpy on random SPD systems.

Sizes cover one tile, ragged last tiles and odd leading dimensions (n = 15 C is rarely a multiple of 32), the bench
size (480 = 32 cameras), grids narrower than a phase's tile count (660 -> 252 tiles on 148 SMs), and systems whose row
blocks no longer fit the shared-memory buffer of the back substitution and stream through it in chunks (n = 700 ... 2100:
1503 = the reduced system of the ccube self-calibration with the poses eliminated, 1920 = 128 cameras; 4800 = more block columns
than CTAs).  Tolerance: relative solution error
<= 1e-9 * cond-scaled bound (the systems are built with cond ~ 1e4), residual <= 1e-11 relative."""
import ctypes as ct

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _solve(n, A, b):
    from pycamset_b200 import _lib
    lib = _lib.load()
    x = np.empty(n)
    info = ct.c_int(0)
    Af = np.asfortranarray(A)
    rc = lib.pcs_spd_solve(0, n, Af.ctypes.data, b.ctypes.data, x.ctypes.data, ct.byref(info))
    return rc, info.value, x


def _spd(n, rng):
    Q, _ = np.linalg.qr(rng.standard_normal((n, n)))
    d = np.logspace(0, 4, n)
    return (Q * d) @ Q.T


@pytest.mark.parametrize("n", [1, 15, 32, 45, 64, 100, 255, 480, 495, 660, 700, 1000, 1503, 1920, 2100])
def test_spd_solve_matches_numpy(n):
    rng = np.random.default_rng(n)
    A = _spd(n, rng)
    b = rng.standard_normal(n)
    L = np.tril(A)                      # only the lower triangle may be read: poison the upper one
    poisoned = L + np.triu(np.full((n, n), np.nan), 1)
    rc, info, x = _solve(n, poisoned, b)
    assert rc == 0 and info == 0
    ref = np.linalg.solve(A, b)
    assert np.linalg.norm(x - ref) <= 1e-9 * np.linalg.norm(ref)
    assert np.linalg.norm(A @ x - b) <= 1e-11 * (np.linalg.norm(A, 2) * np.linalg.norm(x) + np.linalg.norm(b))


def test_spd_solve_flags_indefinite():
    rng = np.random.default_rng(7)
    n = 96
    A = _spd(n, rng)
    A[50, 50] = -1.0
    rc, info, _ = _solve(n, A, rng.standard_normal(n))
    assert rc == -5 and info == 1


def test_spd_solve_large_identity_and_odd_leading_dimension():
    """n = 1921: odd leading dimension (guarded loads instead of 16-byte copies) with chunked row blocks."""
    n = 1921
    rng = np.random.default_rng(3)
    A = np.eye(n) * 2.0
    A[1:, 0] = A[0, 1:] = 1e-3 * rng.standard_normal(n - 1)
    b = rng.standard_normal(n)
    rc, info, x = _solve(n, A, b)
    assert rc == 0 and info == 0
    assert np.linalg.norm(A @ x - b) <= 1e-11 * (np.linalg.norm(A, 2) * np.linalg.norm(x) + np.linalg.norm(b))


def test_spd_solve_more_block_columns_than_ctas():
    """n = 4800: 150 block columns on a grid of at most 148 CTAs -- the panel is wider than the grid (single-tile visits in
    the factorisation) and CTAs own more than one block column in the dataflow back substitution (a CTA walks its blocks
    in descending order; the highest unfinished block can always proceed)."""
    n = 4800
    rng = np.random.default_rng(11)
    B = rng.standard_normal((n, 64)) * 0.05
    A = B @ B.T + np.diag(np.linspace(1.0, 3.0, n))      # SPD, well conditioned, dense
    b = rng.standard_normal(n)
    rc, info, x = _solve(n, A, b)
    assert rc == 0 and info == 0
    ref = np.linalg.solve(A, b)
    assert np.linalg.norm(x - ref) <= 1e-10 * np.linalg.norm(ref)
