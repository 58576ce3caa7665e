"""GPU, BASELINE.json config 5 in shape (128-camera dome, n = 15 C = 1920 reduced system) at a size the CPU oracle
finishes in seconds: every kernel against the oracle, and the device LM (the large-system branch of the dense SPD
solve, the camera-window logic of k_lm_segment_Z at C > 32) against the reference's own solver (scipy TRF + LSMR)
driven by the CUDA callbacks.

Tolerances: residual abs <= 1e-9 px; Jacobian entries rel <= 1e-9 (floor 1e-12); CSR structure
bit-exact; block entries <= 1e-9 sqrt(d_a d_b); cost rel <= 1e-11; LM cost within 5e-3 of the scipy solver's."""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu

C, K = 128, 81


def _rig(M, seed=0, detect_prob=0.5, noise_px=0.1):
    from pycamset_b200 import synthetic as syn
    rig = syn.make_rig(C, M, layout="dome", distortion=True, seed=seed, detect_prob=detect_prob, noise_px=noise_px)
    intr, extr, poses = rig.perturbed(np.random.default_rng(seed + 1), 1e-3)
    params = rig.param_string(intr, extr, poses)
    unfixed = np.ones(params.shape[0], bool)
    unfixed[15 * C:15 * C + 6] = False          # pose 0 is the gauge (template_handler.py:134-137)
    return rig, params, unfixed


def _scaled_close(a, b, da, db, tol=1e-9):
    scale = np.sqrt(np.maximum(da, 1e-300))[..., :, None] * np.sqrt(np.maximum(db, 1e-300))[..., None, :]
    return float(np.max(np.abs(a - b) / scale)) < tol


def test_dome128_kernels_against_the_oracle():
    from pycamset_b200.problem import BundleProblem
    M = 200
    rig, params, unfixed = _rig(M)
    cam, pose, key, uv = rig.cam.numpy(), rig.pose.numpy(), rig.key.numpy(), rig.uv.numpy()
    assert rig.n_obs > 300_000 and len(np.unique(cam)) == C
    o = orc.Problem(0, cam, pose, key, uv, C, M, K, rig.template)
    fm = orc.free_map_from_mask(unfixed)
    with BundleProblem(0, cam, pose, key, uv, C, M, K, template=rig.template, unfixed=unfixed) as p:
        p.set_param_string(params)
        r = p.residual()
        col, rp = p.csr_structure()
        vals = p.jacobian_values()
        ne = p.normal_equations()
        sc, sp, sl = p.segments()
    assert np.max(np.abs(r - o.residual(params))) < 1e-9
    col_o, rp_o = o.csr_structure(fm)
    assert np.array_equal(col, col_o) and np.array_equal(rp, rp_o)
    ref = o.csr_values(params, fm, rp_o)
    # Entries many orders below the largest entry of their own row are products of cancellation in the OpenCV dR/dr
    # formula both sides evaluate (worst case here: -5.3e-6 in a row whose largest entry is 2.5e3); their error is bounded
    # relative to the row scale, not to themselves: |d| <= 1e-9 max(|ref|, 1e-8 rowmax).
    rowmax = np.maximum.reduceat(np.abs(ref), rp_o[:-1])
    floor = np.repeat(rowmax, np.diff(rp_o)) * 1e-8
    assert np.max(np.abs(vals - ref) / np.maximum(np.abs(ref), np.maximum(floor, 1e-12))) < 1e-9
    pair = cam.astype(np.int64) * M + pose
    uniq, seg, cnt = np.unique(pair, return_inverse=True, return_counts=True)
    assert np.array_equal(sc.astype(np.int64) * M + sp, uniq) and np.array_equal(sl, cnt)
    U, gc, V, gp, W, cost = o.normal_blocks(params, seg.astype(np.int32), len(uniq))
    dU = np.einsum("cii->ci", U); dV = np.einsum("mii->mi", V)
    assert _scaled_close(ne["U"], U, dU, dU)
    assert _scaled_close(ne["V"], V, dV, dV)
    assert _scaled_close(ne["W"], W, dU[sc], dV[sp])
    assert np.max(np.abs(ne["gc"] - gc) / np.sqrt(np.maximum(dU, 1e-300) * cost)) < 1e-9
    assert np.max(np.abs(ne["gp"] - gp) / np.sqrt(np.maximum(dV, 1e-300) * cost)) < 1e-9
    assert abs(ne["cost"] - cost) <= 1e-11 * cost


def test_dome128_lm_matches_the_reference_solver():
    """n = 1920 reduced camera system: device LM vs scipy TRF + LSMR (x_scale='jac': the reference's call,
    optimisation_handling.py:88-98) on the CUDA callbacks, from the same start."""
    from scipy.optimize import least_squares
    from scipy.sparse import csr_array
    from pycamset_b200.problem import BundleProblem
    M = 40
    rig, params, unfixed = _rig(M, seed=3)
    cam, pose, key, uv = rig.cam.numpy(), rig.pose.numpy(), rig.key.numpy(), rig.uv.numpy()
    with BundleProblem(0, cam, pose, key, uv, C, M, K, template=rig.template, unfixed=unfixed) as p:
        p.set_param_string(params)
        x0 = params[unfixed]
        r0 = p.residual(x0)
        x_lm, st = p.lm_solve(x0, max_iter=60, ftol=1e-12, xtol=1e-12, gtol=1e-10)
        r_lm = p.residual(x_lm)
        col, rp = p.csr_structure()
        jac = lambda x: csr_array((p.jacobian_values(x), col, rp), shape=(2 * p.n_obs, p.n_free))
        res = least_squares(lambda x: p.residual(x), x0, jac=jac, x_scale="jac", ftol=1e-12, xtol=1e-12, gtol=1e-10,
                            max_nfev=40)
    c0, c_lm = 0.5 * float(r0 @ r0), 0.5 * float(r_lm @ r_lm)
    assert st["status"] >= 0, st
    assert abs(c_lm - st["cost_final"]) <= 1e-9 * c_lm
    assert c_lm < 0.5 * c0 and res.cost < 0.5 * c0
    assert c_lm <= res.cost * (1 + 5e-3), (c_lm, res.cost, st)
    px_lm = np.mean(np.linalg.norm(r_lm.reshape(-1, 2), axis=1))
    assert px_lm < 0.14                  # noise is N(0, 0.1 px) per coordinate -> mean norm ~0.125 px
