"""CPU: properties of the oracle that do not depend on the reference's outputs -- an independent second pin of the
checker besides the golden vectors (tests/test_oracle_golden.py).

  * the dense Jacobian rows are the derivative of the residual: central differences of `residual` in every column of
    every observation's row pair (both chains), relative error <= 2e-6 of the row scale (FD step 1e-6 relative);
  * the block normal equations equal the blocks of J^T J / J^T r built from the dense Jacobian (<= 1e-10 relative to
    the block diagonals) and r.r equals the cost;
  * additivity over a split of the observation table and invariance under a permutation of its rows (<= 1e-11),
    the two size-independent properties the full-size GPU tests rely on.
"""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.helpers import SYNTH_CASES, available, load_case, oracle_problem

CASES = available(SYNTH_CASES)


def _columns(p, c, m, k):
    """Global parameter-string columns of one observation's dense row (matflow order)."""
    cols = list(range(9 * c, 9 * c + 9)) + list(range(9 * p.C + 6 * c, 9 * p.C + 6 * c + 6)) + \
        list(range(15 * p.C + 6 * m, 15 * p.C + 6 * m + 6))
    if p.chain == 1:
        cols += list(range(15 * p.C + 6 * p.M + 3 * k, 15 * p.C + 6 * p.M + 3 * k + 3))
    return cols


@pytest.mark.parametrize("name", CASES)
def test_jacobian_is_the_derivative_of_the_residual(name):
    g = load_case(name)
    p = oracle_problem(g)
    params = g["param0"].astype(np.float64)
    J, r = p.jacobian_dense(params)
    assert np.array_equal(r, p.residual(params))
    rng = np.random.default_rng(0)
    for i in rng.choice(p.N, size=min(p.N, 12), replace=False):
        cols = _columns(p, int(p.cam[i]), int(p.pose[i]), int(p.key[i]))
        assert len(cols) == p.P
        scale = np.abs(J[2 * i:2 * i + 2]).max()
        for a, col in enumerate(cols):
            h = 1e-6 * max(1.0, abs(params[col]))
            pp, pm = params.copy(), params.copy()
            pp[col] += h
            pm[col] -= h
            fd = (p.residual(pp)[2 * i:2 * i + 2] - p.residual(pm)[2 * i:2 * i + 2]) / (2 * h)
            assert np.all(np.abs(fd - J[2 * i:2 * i + 2, a]) <= 2e-6 * scale + 1e-7), (name, i, a)


def _segments(cam, pose, M):
    keys = cam.astype(np.int64) * M + pose
    uniq, seg = np.unique(keys, return_inverse=True)
    return seg.astype(np.int32), len(uniq), (uniq // M).astype(int), (uniq % M).astype(int)


@pytest.mark.parametrize("name", [c for c in CASES if c.endswith("template")])
def test_blocks_are_the_blocks_of_JtJ(name):
    g = load_case(name)
    p = oracle_problem(g)
    params = g["param0"].astype(np.float64)
    seg, n_seg, seg_c, seg_m = _segments(p.cam, p.pose, p.M)
    U, gc, V, gp, W, cost = p.normal_blocks(params, seg, n_seg)
    J, r = p.jacobian_dense(params)
    assert abs(cost - r @ r) <= 1e-11 * (r @ r)
    Ju = J.reshape(p.N, 2, p.P)
    A, B = Ju[:, :, :15], Ju[:, :, 15:21]          # camera (intrinsic + extrinsic) and pose columns
    rr = r.reshape(p.N, 2)
    for c in range(p.C):
        sel = p.cam == c
        Uc = np.einsum("nra,nrb->ab", A[sel], A[sel])
        d = np.sqrt(np.outer(np.diag(Uc), np.diag(Uc))) + 1e-300
        assert np.max(np.abs(U[c] - Uc) / d) <= 1e-10
        assert np.allclose(gc[c], np.einsum("nra,nr->a", A[sel], rr[sel]), rtol=1e-10, atol=1e-10 * np.abs(gc[c]).max())
    for m in range(p.M):
        sel = p.pose == m
        Vm = np.einsum("nra,nrb->ab", B[sel], B[sel])
        d = np.sqrt(np.outer(np.diag(Vm), np.diag(Vm))) + 1e-300
        assert np.max(np.abs(V[m] - Vm) / d) <= 1e-10
    for s in range(n_seg):
        sel = seg == s
        Ws = np.einsum("nra,nrb->ab", A[sel], B[sel])
        d = np.sqrt(np.outer(np.diag(U[seg_c[s]]), np.diag(V[seg_m[s]]))) + 1e-300
        assert np.max(np.abs(W[s] - Ws) / d) <= 1e-10


@pytest.mark.parametrize("name", [c for c in CASES if c.endswith("template")])
def test_blocks_are_additive_and_permutation_invariant(name):
    g = load_case(name)
    p = oracle_problem(g)
    params = g["param0"].astype(np.float64)
    seg, n_seg, _, _ = _segments(p.cam, p.pose, p.M)
    full = p.normal_blocks(params, seg, n_seg)

    def sub(idx):
        q = orc.Problem(0, p.cam[idx], p.pose[idx], p.key[idx], p.uv[idx], p.C, p.M, p.K, template=p.template)
        return q.normal_blocks(params, seg[idx], n_seg)

    rng = np.random.default_rng(1)
    perm = rng.permutation(p.N)
    half = p.N // 2
    a, b = sub(perm[:half]), sub(perm[half:])
    shuffled = sub(perm)
    for k in range(5):
        scale = np.abs(full[k]).max()
        assert np.max(np.abs(a[k] + b[k] - full[k])) <= 1e-11 * scale
        assert np.max(np.abs(shuffled[k] - full[k])) <= 1e-11 * scale
    assert abs(a[5] + b[5] - full[5]) <= 1e-11 * full[5]


def test_projection_matches_opencv_projectpoints():
    """The reference's own bundle_correctness_test pins its projection model to cv2.projectPoints (< 1e-4 px); the
    oracle is held to the same external ground truth at 1e-8 px, through a non-trivial pose and extrinsic."""
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(11)
    C, M, K = 3, 2, 40
    template = np.column_stack([rng.uniform(-0.03, 0.03, K), rng.uniform(-0.03, 0.03, K), rng.uniform(-0.005, 0.005, K)])
    intr = np.column_stack([rng.uniform(900, 1300, C), 500 + rng.normal(0, 10, C), rng.uniform(900, 1300, C),
                            500 + rng.normal(0, 10, C), rng.normal(0, 0.05, C), rng.normal(0, 0.02, C),
                            rng.normal(0, 1e-3, C), rng.normal(0, 1e-3, C), rng.normal(0, 1e-2, C)])
    extr = np.column_stack([rng.normal(0, 0.2, (C, 3)), rng.normal(0, 0.01, (C, 2)), rng.uniform(0.2, 0.3, C)])
    poses = np.column_stack([rng.normal(0, 0.3, (M, 3)), rng.normal(0, 0.01, (M, 3))])
    cam, pose, key = np.meshgrid(np.arange(C), np.arange(M), np.arange(K), indexing="ij")
    cam, pose, key = cam.ravel(), pose.ravel(), key.ravel()
    p = orc.Problem(0, cam, pose, key, np.zeros((cam.size, 2)), C, M, K, template=template)
    params = np.concatenate([intr.ravel(), extr.ravel(), poses.ravel()])
    uv = p.residual(params).reshape(-1, 2)           # observed pixel = 0, so the residual is the projection
    for c in range(C):
        Kc = np.array([[intr[c, 0], 0, intr[c, 1]], [0, intr[c, 2], intr[c, 3]], [0, 0, 1.0]])
        Rc, _ = cv2.Rodrigues(extr[c, :3])
        for m in range(M):
            Rm, _ = cv2.Rodrigues(poses[m, :3])
            Xw = template @ Rm.T + poses[m, 3:]
            ref, _ = cv2.projectPoints(Xw, extr[c, :3], extr[c, 3:], Kc, intr[c, 4:9])
            sel = (cam == c) & (pose == m)
            assert np.max(np.abs(uv[sel] - ref.reshape(-1, 2))) < 1e-8
            # and the rotation convention: R(rvec) of the chain is OpenCV's Rodrigues
            assert np.allclose((Xw @ Rc.T + extr[c, 3:])[:, 2] > 0, True)
