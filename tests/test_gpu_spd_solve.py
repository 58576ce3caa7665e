"""GPU: the persistent tiled Cholesky solve (csrc/pcs_chol.cu) against numpy on random SPD systems.

Sizes cover one tile, ragged last tiles and odd leading dimensions (n = 15 C is rarely a multiple of 32), the bench
size (480 = 32 cameras) and grids narrower than a phase's tile count (660 -> 252 tiles on 148 SMs).  Systems too
large for the kernel's shared-memory row-block buffer (n > 672) are refused; pcs_lm_solve uses cuSOLVER there.  Tolerance: relative solution error
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


@pytest.mark.parametrize("n", [1, 15, 32, 45, 64, 100, 255, 480, 495, 660])
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


def test_spd_solve_refuses_oversized():
    n = 1920
    A = np.eye(n)
    rc, _, _ = _solve(n, A, np.ones(n))
    assert rc == -4
