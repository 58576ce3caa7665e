"""GPU, BASELINE.json config 4 at full size (32-camera ring x 2000 poses, ~2.6 M observations).

PRIMARY check: every output of the fused normal-equation kernel (U, V, W, g_c, g_m, r.r) and of the residual kernel
against the CPU ORACLE (oracle/ba_oracle.c) on the full 2.59 M-observation table.

Secondary, size-independent properties of the fused kernel against independent device paths:

  * r.r from K_ne equals the sum of squares of the residual kernel's output;
  * U / V / W / g_c / g_m equal the corresponding blocks of the DENSE J^T J / J^T r accumulated by a different
    kernel (per-entry atomics, reference dR formula instead of the tangent form + epilogue);
  * shuffling the observation table leaves every block unchanged (only the summation order moves);
  * linearity: the blocks of the first half plus the blocks of the second half of the table equal the blocks of
    the whole table.
Tolerance: 1e-9 relative to sqrt(diag_a diag_b) per entry."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ring32():
    import torch
    from pycamset_b200 import synthetic as syn
    rig = syn.make_rig(32, 2000, layout="ring", distortion=True, seed=0, detect_prob=1.0, device="cuda:0", order="cam")
    rng = np.random.default_rng(1)
    intr, extr, poses = rig.perturbed(rng, 1e-3)
    params = rig.param_string(intr, extr, poses)
    unfixed = np.ones(params.shape[0], bool)
    unfixed[15 * 32:15 * 32 + 6] = False
    return rig, params, unfixed


def _problem(rig, unfixed, sel=None):
    from pycamset_b200.problem import BundleProblem
    cam, pose, key, uv = rig.cam, rig.pose, rig.key, rig.uv
    if sel is not None:
        cam, pose, key, uv = cam[sel], pose[sel], key[sel], uv[sel]
    return BundleProblem(0, cam, pose, key, uv, 32, 2000, 81, template=rig.template, unfixed=unfixed, device=0)


def _scaled_close(a, b, da, db, tol=1e-9):
    scale = np.sqrt(np.maximum(da, 1e-300))[..., :, None] * np.sqrt(np.maximum(db, 1e-300))[..., None, :]
    return float(np.max(np.abs(a - b) / scale)) < tol


def test_full_size_blocks_against_dense_path_and_residual(ring32):
    rig, params, unfixed = ring32
    assert rig.n_obs > 2_000_000
    with _problem(rig, unfixed) as p:
        p.set_param_string(params)
        ne = p.normal_equations()
        r = p.residual()
        sc, sp, sl = p.segments()
        JtJ, Jtr, cost_d = p.normal_dense()
    cost = ne["cost"]
    assert abs(cost - float(r @ r)) <= 1e-11 * cost and abs(cost - cost_d) <= 1e-11 * cost
    assert int(sl.sum()) == rig.n_obs and np.all(np.diff(sc.astype(np.int64) * 2000 + sp) > 0)
    C, M = 32, 2000
    fm = np.full(params.shape[0], -1, np.int64)
    fm[unfixed] = np.arange(int(unfixed.sum()))
    cam_idx = np.stack([np.r_[9 * c:9 * c + 9, 9 * C + 6 * c:9 * C + 6 * c + 6] for c in range(C)])
    pose_idx = 15 * C + 6 * np.arange(M)[:, None] + np.arange(6)[None, :]
    dU = np.einsum("cii->ci", ne["U"]); dV = np.einsum("mii->mi", ne["V"])
    fc = fm[cam_idx]
    U_d = JtJ[fc[:, :, None], fc[:, None, :]]
    assert _scaled_close(ne["U"], U_d, dU, dU)
    assert np.max(np.abs(ne["gc"] - Jtr[fc]) / np.sqrt(np.maximum(dU, 1e-300) * cost)) < 1e-9
    free_pose = np.arange(1, M)                       # pose 0 is fixed: absent from the dense system
    fp = fm[pose_idx[free_pose]]
    V_d = JtJ[fp[:, :, None], fp[:, None, :]]
    assert _scaled_close(ne["V"][free_pose], V_d, dV[free_pose], dV[free_pose])
    assert np.max(np.abs(ne["gp"][free_pose] - Jtr[fp]) / np.sqrt(np.maximum(dV[free_pose], 1e-300) * cost)) < 1e-9
    pick = np.flatnonzero(sp > 0)[:: max(1, len(sp) // 4000)]      # a few thousand W segments spread over the table
    W_d = JtJ[fm[cam_idx[sc[pick]]][:, :, None], fm[pose_idx[sp[pick]]][:, None, :]]
    assert _scaled_close(ne["W"][pick], W_d, dU[sc[pick]], dV[sp[pick]])


def test_full_size_order_invariance_and_linearity(ring32):
    import torch
    rig, params, unfixed = ring32
    with _problem(rig, unfixed) as p:
        p.set_param_string(params)
        ne = p.normal_equations()
    g = torch.Generator(device="cuda:0"); g.manual_seed(7)
    perm = torch.randperm(rig.n_obs, device="cuda:0", generator=g)
    with _problem(rig, unfixed, perm) as p:
        p.set_param_string(params)
        ne_p = p.normal_equations()
    for k in ("U", "V", "W", "gc", "gp"):
        assert np.max(np.abs(ne[k] - ne_p[k])) <= 1e-11 * np.max(np.abs(ne[k])), k
    assert abs(ne["cost"] - ne_p["cost"]) <= 1e-12 * ne["cost"]
    half = rig.n_obs // 2
    parts = []
    for sel in (perm[:half], perm[half:]):
        with _problem(rig, unfixed, sel) as p:
            p.set_param_string(params)
            parts.append((p.normal_equations(with_W=False), ))
    for k in ("U", "V", "gc", "gp"):
        s = parts[0][0][k] + parts[1][0][k]
        assert np.max(np.abs(ne[k] - s)) <= 1e-11 * np.max(np.abs(ne[k])), k
    assert abs(ne["cost"] - parts[0][0]["cost"] - parts[1][0]["cost"]) <= 1e-12 * ne["cost"]


def test_full_size_blocks_and_residual_against_the_oracle(ring32):
    """Config 4 at full size against the CPU oracle (not against another kernel of this library): residual abs <= 1e-9 px,
    block entries <= 1e-9 sqrt(d_a d_b), gradients <= 1e-9 sqrt(d_a cost), cost rel <= 1e-11."""
    from oracle import oracle as orc
    rig, params, unfixed = ring32
    cam, pose, key, uv = (t.cpu().numpy() for t in (rig.cam, rig.pose, rig.key, rig.uv))
    C, M = 32, 2000
    o = orc.Problem(0, cam, pose, key, uv, C, M, 81, rig.template)
    with _problem(rig, unfixed) as p:
        p.set_param_string(params)
        ne = p.normal_equations()
        r = p.residual()
        sc, sp, sl = p.segments()
    r_o = o.residual(params)
    assert r.shape == r_o.shape and np.max(np.abs(r - r_o)) < 1e-9
    pair = cam.astype(np.int64) * M + pose
    uniq, seg = np.unique(pair, return_inverse=True)
    assert np.array_equal(sc.astype(np.int64) * M + sp, uniq)
    U, gc, V, gp, W, cost = o.normal_blocks(params, seg.astype(np.int32), len(uniq))
    assert abs(ne["cost"] - cost) <= 1e-11 * cost and abs(cost - float(r_o @ r_o)) <= 1e-11 * cost
    dU = np.einsum("cii->ci", U); dV = np.einsum("mii->mi", V)
    assert _scaled_close(ne["U"], U, dU, dU)
    assert _scaled_close(ne["V"], V, dV, dV)
    assert _scaled_close(ne["W"], W, dU[sc], dV[sp])
    assert np.max(np.abs(ne["gc"] - gc) / np.sqrt(np.maximum(dU, 1e-300) * cost)) < 1e-9
    assert np.max(np.abs(ne["gp"] - gp) / np.sqrt(np.maximum(dV, 1e-300) * cost)) < 1e-9
