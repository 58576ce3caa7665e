"""ctypes front-end of oracle/ba_oracle.c.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
``--impl reference`` legs.  Nothing under pycamset_b200/ may import this module.
"""
from __future__ import annotations

import ctypes as ct
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = None

_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


def build(force: bool = False) -> Path:
    """Compile liboracle.so next to the source (gcc, -O2, OpenMP, no fast-math)."""
    so = _HERE / "liboracle.so"
    src = _HERE / "ba_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-s", "-C", str(_HERE), "-B", "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ct.CDLL(str(build()))
        _LIB.oracle_csr_structure.restype = ct.c_int64
    return _LIB


def _opt(a):
    return None if a is None else a.ctypes.data_as(ct.c_void_p)


class Problem:
    """Plain container for one observation table + chain (mirrors the arguments the reference closes over
    in make_full_loss_fn / make_jacobean, abstract_function_blocks.py:656-667)."""

    def __init__(self, chain, cam, pose, key, uv, C, M, K, template=None):
        self.chain = int(chain)
        self.cam = np.ascontiguousarray(cam, np.int32)
        self.pose = np.ascontiguousarray(pose, np.int32)
        self.key = np.ascontiguousarray(key, np.int32)
        self.uv = np.ascontiguousarray(uv, np.float64).reshape(-1, 2)
        self.N = self.cam.shape[0]
        self.C, self.M, self.K = int(C), int(M), int(K)
        self.template = None if template is None else np.ascontiguousarray(template, np.float64).reshape(-1, 3)
        self.P = 21 if self.chain == 0 else 24
        self.L = 15 * self.C + 6 * self.M + (3 * self.K if self.chain == 1 else 0)

    @classmethod
    def from_dd(cls, chain, dd, template=None, C=None, M=None, K=None):
        """dd: N x 5 float64 [cam, img, key, u, v] (target_detections.py:51-55)."""
        dd = np.asarray(dd, np.float64)
        C = int(dd[:, 0].max()) + 1 if C is None else C
        M = int(dd[:, 1].max()) + 1 if M is None else M
        K = int(dd[:, 2].max()) + 1 if K is None else K
        return cls(chain, dd[:, 0], dd[:, 1], dd[:, 2], dd[:, 3:5], C, M, K, template)

    def _head(self):
        return (self.chain, ct.c_int64(self.N), self.cam, self.pose, self.key)

    def _check(self, params):
        params = np.ascontiguousarray(params, np.float64)
        if params.shape[0] != self.L:
            raise ValueError(f"parameter string has length {params.shape[0]}, expected {self.L}")
        return params

    def residual(self, params):
        params = self._check(params)
        r = np.empty(2 * self.N)
        f = lib().oracle_residual
        f.argtypes = [ct.c_int, ct.c_int64, _i32p, _i32p, _i32p, _f64p, ct.c_int, ct.c_int, ct.c_int, _f64p,
                      ct.c_void_p, _f64p]
        f(*self._head(), self.uv, self.C, self.M, self.K, params, _opt(self.template), r)
        return r

    def jacobian_dense(self, params):
        params = self._check(params)
        J = np.empty((2 * self.N, self.P))
        r = np.empty(2 * self.N)
        f = lib().oracle_jacobian_dense
        f.argtypes = [ct.c_int, ct.c_int64, _i32p, _i32p, _i32p, _f64p, ct.c_int, ct.c_int, ct.c_int, _f64p,
                      ct.c_void_p, _f64p, _f64p]
        f(*self._head(), self.uv, self.C, self.M, self.K, params, _opt(self.template), J, r)
        return J, r

    def csr_structure(self, free_map):
        free_map = np.ascontiguousarray(free_map, np.int32)
        f = lib().oracle_csr_structure
        f.argtypes = [ct.c_int, ct.c_int64, _i32p, _i32p, _i32p, ct.c_int, ct.c_int, ct.c_int, _i32p, ct.c_void_p,
                      ct.c_void_p]
        nnz = f(*self._head(), self.C, self.M, self.K, free_map, None, None)
        col = np.empty(nnz, np.int64)
        rp = np.empty(2 * self.N + 1, np.int64)
        f(*self._head(), self.C, self.M, self.K, free_map, _opt(col), _opt(rp))
        return col, rp

    def csr_values(self, params, free_map, row_ptr):
        params = self._check(params)
        free_map = np.ascontiguousarray(free_map, np.int32)
        vals = np.empty(int(row_ptr[-1]))
        f = lib().oracle_csr_values
        f.argtypes = [ct.c_int, ct.c_int64, _i32p, _i32p, _i32p, _f64p, ct.c_int, ct.c_int, ct.c_int, _f64p,
                      ct.c_void_p, _i32p, _i64p, _f64p]
        f(*self._head(), self.uv, self.C, self.M, self.K, params, _opt(self.template), free_map,
          np.ascontiguousarray(row_ptr, np.int64), vals)
        return vals

    def normal_dense(self, params, free_map):
        params = self._check(params)
        free_map = np.ascontiguousarray(free_map, np.int32)
        n_free = int(free_map.max()) + 1
        JtJ = np.empty((n_free, n_free))
        Jtr = np.empty(n_free)
        cost = ct.c_double()
        f = lib().oracle_normal_dense
        f.argtypes = [ct.c_int, ct.c_int64, _i32p, _i32p, _i32p, _f64p, ct.c_int, ct.c_int, ct.c_int, _f64p,
                      ct.c_void_p, _i32p, ct.c_int64, _f64p, _f64p, ct.POINTER(ct.c_double)]
        f(*self._head(), self.uv, self.C, self.M, self.K, params, _opt(self.template), free_map, n_free, JtJ, Jtr,
          ct.byref(cost))
        return JtJ, Jtr, cost.value

    def normal_blocks(self, params, seg, n_seg):
        """Chain 0 only.  seg[i] = index of the (camera, pose) pair of observation i."""
        if self.chain != 0:
            raise ValueError("block normal equations are defined for the template chain only")
        params = self._check(params)
        seg = np.ascontiguousarray(seg, np.int32)
        U = np.empty((self.C, 15, 15)); gc = np.empty((self.C, 15))
        V = np.empty((self.M, 6, 6)); gp = np.empty((self.M, 6))
        W = np.empty((n_seg, 15, 6))
        cost = ct.c_double()
        f = lib().oracle_normal_blocks
        f.argtypes = [ct.c_int64, _i32p, _i32p, _i32p, _f64p, _i32p, ct.c_int, ct.c_int, ct.c_int, ct.c_int64, _f64p,
                      ct.c_void_p, _f64p, _f64p, _f64p, _f64p, _f64p, ct.POINTER(ct.c_double)]
        f(ct.c_int64(self.N), self.cam, self.pose, self.key, self.uv, seg, self.C, self.M, self.K, n_seg, params,
          _opt(self.template), U, gc, V, gp, W, ct.byref(cost))
        return U, gc, V, gp, W, cost.value


def costfn(dd, im_points, proj, ints, dists):
    """Restatement of numpy_bundle_adjustment_costfn (compiled_helpers.py:517-549): errors (2N,) in dd row order."""
    dd = np.asarray(dd, np.float64)
    cam = np.ascontiguousarray(dd[:, 0], np.int32); pose = np.ascontiguousarray(dd[:, 1], np.int32)
    key = np.ascontiguousarray(dd[:, 2], np.int32); uv = np.ascontiguousarray(dd[:, 3:5], np.float64)
    im_points = np.ascontiguousarray(im_points, np.float64)
    M, K = im_points.shape[0], im_points.shape[1]
    proj = np.ascontiguousarray(proj, np.float64); ints = np.ascontiguousarray(ints, np.float64)
    dists = np.ascontiguousarray(np.asarray(dists, np.float64).reshape(-1, 5))
    out = np.empty(2 * dd.shape[0])
    f = lib().oracle_costfn
    f.argtypes = [ct.c_int64, _i32p, _i32p, _i32p, _f64p, ct.c_int, ct.c_int, ct.c_int, _f64p, _f64p, _f64p, _f64p, _f64p]
    f(ct.c_int64(dd.shape[0]), cam, pose, key, uv, proj.shape[0], M, K, im_points.reshape(-1), proj.reshape(-1),
      ints.reshape(-1), dists.reshape(-1), out)
    return out


def set_threads(n: int) -> None:
    lib().oracle_set_threads(int(n))


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def block_projection(q, X):
    fun = np.empty(2); jac = np.empty(24)
    f = lib().oracle_block_projection
    f.argtypes = [_f64p, _f64p, _f64p, _f64p]
    f(np.ascontiguousarray(q, np.float64), np.ascontiguousarray(X, np.float64), fun, jac)
    return fun, jac.reshape(2, 12)


def block_rigid(p, X):
    fun = np.empty(3); jac = np.empty(27)
    f = lib().oracle_block_rigid
    f.argtypes = [_f64p, _f64p, _f64p, _f64p]
    f(np.ascontiguousarray(p, np.float64), np.ascontiguousarray(X, np.float64), fun, jac)
    return fun, jac.reshape(3, 9)


def block_rodrigues_jac(r):
    out = np.empty(27)
    f = lib().oracle_block_rodrigues_jac
    f.argtypes = [_f64p, _f64p]
    f(np.ascontiguousarray(r, np.float64), out)
    return out.reshape(3, 9)


def free_map_from_mask(unfixed) -> np.ndarray:
    """unfixed boolean mask over the parameter string -> free column index or -1
    (the `conversion` renumbering of abstract_function_blocks.py:482-485)."""
    unfixed = np.asarray(unfixed, bool)
    fm = np.full(unfixed.shape[0], -1, np.int32)
    fm[unfixed] = np.arange(int(unfixed.sum()), dtype=np.int32)
    return fm
