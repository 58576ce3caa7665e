"""GPU: the stream-K FP64 tensor-path SYRK of the reduced camera system (csrc/pcs_schur.cu) against numpy.

Shapes cover ragged tiles (n not a multiple of 96), odd n (8-byte copy path), k not a multiple of the 16-column slab,
fewer work units than SMs, and the bench shape (480 x 12000).  Tolerance: |dS| <= 1e-12 * sum_k |Z_ik||Z_jk| + tiny
(FP64 accumulation in a different order)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,k", [(1, 1), (15, 6), (45, 30), (96, 16), (97, 50), (120, 594), (480, 1200), (495, 333), (480, 12000)])
def test_syrk_matches_numpy(n, k):
    from pycamset_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(n * 131 + k)
    Z = rng.standard_normal((n, k))
    S0 = rng.standard_normal((n, n))
    S0 = S0 + S0.T
    Zf = np.asfortranarray(Z)          # column-major [k][n]: column index slowest
    S = np.asfortranarray(S0.copy())
    rc = lib.pcs_syrk_sub(0, n, k, Zf.ctypes.data, S.ctypes.data)
    assert rc == 0, lib.pcs_last_error()
    ref = S0 - Z @ Z.T
    bound = 1e-12 * (np.abs(Z) @ np.abs(Z).T) + 1e-13
    lo = np.tril_indices(n)
    assert np.all(np.abs(S[lo] - ref[lo]) <= bound[lo])
    up = np.triu_indices(n, 1)
    assert np.array_equal(S[up], S0[up])   # the strict upper triangle is not touched
