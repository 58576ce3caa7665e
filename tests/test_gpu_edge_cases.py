"""GPU: edge cases of the C-ABI path -- empty and tiny observation tables, ragged segments, a pose / camera nobody
observes, everything fixed, invalid input.  Reference behaviour for bad input is a Python exception
(SURVEY.md 8b "Errors"); here it is a PcsError carrying the library's message."""
import numpy as np
import pytest

from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _rig(C=4, M=6, seed=2, detect_prob=0.7):
    from pycamset_b200 import synthetic as syn
    rig = syn.make_rig(C, M, distortion=True, seed=seed, detect_prob=detect_prob)
    rng = np.random.default_rng(seed)
    intr, extr, poses = rig.perturbed(rng)
    return rig, rig.param_string(intr, extr, poses)


def _check_against_oracle(cam, pose, key, uv, C, M, template, params, unfixed=None):
    from pycamset_b200.problem import BundleProblem
    o = orc.Problem(0, cam, pose, key, uv, C, M, 81, template)
    with BundleProblem(0, cam, pose, key, uv, C, M, 81, template=template, unfixed=unfixed) as p:
        p.set_param_string(params)
        r = p.residual()
        assert r.shape == (2 * len(cam),)
        if len(cam):
            assert np.max(np.abs(r - o.residual(params))) < 1e-9
        ne = p.normal_equations()
        sc, sp, sl = p.segments()
        pair = np.asarray(cam, np.int64) * M + np.asarray(pose, np.int64)
        uniq, seg = np.unique(pair, return_inverse=True)
        assert np.array_equal(sc.astype(np.int64) * M + sp, uniq)
        U, gc, V, gp, W, cost = o.normal_blocks(params, seg.astype(np.int32), len(uniq))
        scale = max(np.max(np.abs(U)), 1e-300)
        for a, b in ((ne["U"], U), (ne["V"], V), (ne["gc"], gc), (ne["gp"], gp)):
            assert np.max(np.abs(a - b)) <= 1e-10 * max(np.max(np.abs(b)), scale * 1e-12)
        if len(uniq):
            assert np.max(np.abs(ne["W"] - W)) <= 1e-10 * np.max(np.abs(W))
        assert abs(ne["cost"] - cost) <= 1e-11 * max(cost, 1e-300)
        return p.n_free, p.nnz, ne


def test_empty_observation_table():
    rig, params = _rig()
    e = np.zeros(0, np.int32)
    n_free, nnz, ne = _check_against_oracle(e, e, e, np.zeros((0, 2)), 4, 6, rig.template, params)
    assert nnz == 0 and ne["cost"] == 0.0 and not ne["U"].any() and ne["W"].shape == (0, 15, 6)


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 65])
def test_tiny_tables(n):
    """Fewer observations than one warp batch, odd counts (masked k-steps), batch boundaries."""
    rig, params = _rig(detect_prob=1.0)
    sel = np.random.default_rng(n).choice(rig.n_obs, n, replace=False)
    sel.sort()
    _check_against_oracle(rig.cam.numpy()[sel], rig.pose.numpy()[sel], rig.key.numpy()[sel], rig.uv.numpy()[sel], 4, 6,
                          rig.template, params)


def test_single_observation_segments_and_unobserved_blocks():
    """Every (camera, pose) pair seen once at most; camera 2 and pose 3 never observed: their blocks stay zero."""
    rig, params = _rig(detect_prob=1.0)
    cam, pose, key, uv = rig.cam.numpy(), rig.pose.numpy(), rig.key.numpy(), rig.uv.numpy()
    pair = cam.astype(np.int64) * 6 + pose
    _, first = np.unique(pair, return_index=True)
    keep = first[(cam[first] != 2) & (pose[first] != 3)]
    n_free, nnz, ne = _check_against_oracle(cam[keep], pose[keep], key[keep], uv[keep], 4, 6, rig.template, params)
    assert not ne["U"][2].any() and not ne["gc"][2].any() and not ne["V"][3].any() and not ne["gp"][3].any()


def test_unsorted_input_order():
    """dd rows in random order (the reference imposes none): outputs keep the caller's row order."""
    rig, params = _rig()
    perm = np.random.default_rng(0).permutation(rig.n_obs)
    _check_against_oracle(rig.cam.numpy()[perm], rig.pose.numpy()[perm], rig.key.numpy()[perm], rig.uv.numpy()[perm], 4, 6,
                          rig.template, params)


def test_everything_fixed_and_partial_masks():
    from pycamset_b200.problem import BundleProblem
    rig, params = _rig()
    cam, pose, key, uv = rig.cam.numpy(), rig.pose.numpy(), rig.key.numpy(), rig.uv.numpy()
    o = orc.Problem(0, cam, pose, key, uv, 4, 6, 81, rig.template)
    none = np.zeros(params.shape[0], bool)
    with BundleProblem(0, cam, pose, key, uv, 4, 6, 81, template=rig.template, unfixed=none) as p:
        p.set_param_string(params)
        assert p.n_free == 0 and p.nnz == 0
        assert np.max(np.abs(p.residual() - o.residual(params))) < 1e-9
        col, rp = p.csr_structure()
        assert col.shape == (0,) and not rp.any()
    # camera 1 intrinsics fixed, camera 2 extrinsics fixed, poses 0 and 4 fixed
    mask = np.ones(params.shape[0], bool)
    mask[9:18] = False; mask[36 + 12:36 + 18] = False; mask[60:66] = False; mask[60 + 24:60 + 30] = False
    fm = orc.free_map_from_mask(mask)
    with BundleProblem(0, cam, pose, key, uv, 4, 6, 81, template=rig.template, unfixed=mask) as p:
        p.set_param_string(params)
        col_o, rp_o = o.csr_structure(fm)
        col, rp = p.csr_structure()
        assert np.array_equal(col, col_o) and np.array_equal(rp, rp_o)
        ref = o.csr_values(params, fm, rp_o)
        assert np.max(np.abs(p.jacobian_values() - ref) / np.maximum(np.abs(ref), 1e-12)) < 1e-9
        x = params[mask]
        assert np.array_equal(p.get_param_string(), params)
        p.residual(x * 1.0)
        assert np.array_equal(p.get_param_string(), params)        # fixed entries untouched by the x scatter


def test_invalid_input_raises():
    from pycamset_b200 import _lib
    from pycamset_b200.problem import BundleProblem
    rig, params = _rig()
    cam, pose, key, uv = rig.cam.numpy().copy(), rig.pose.numpy(), rig.key.numpy(), rig.uv.numpy()
    cam[5] = 4                                                         # camera index out of range
    with pytest.raises(_lib.PcsError, match="outside"):
        BundleProblem(0, cam, pose, key, uv, 4, 6, 81, template=rig.template)
    with pytest.raises(ValueError):
        BundleProblem(0, cam[:-1], pose, key, uv, 4, 6, 81, template=rig.template)
    with pytest.raises(_lib.PcsError):
        BundleProblem(7, rig.cam.numpy(), pose, key, uv, 4, 6, 81, template=rig.template)   # unknown chain id
    with BundleProblem(0, rig.cam.numpy(), pose, key, uv, 4, 6, 81, template=rig.template) as p:
        with pytest.raises(ValueError):
            p.set_param_string(params[:-1])
        with pytest.raises(ValueError):
            p.residual(np.zeros(3))
    with BundleProblem(1, rig.cam.numpy(), pose, key, uv, 4, 6, 81) as p:                  # self-calibration chain
        with pytest.raises(_lib.PcsError, match="template chain"):
            p.set_normal_precision(True)                                                   # the mixed kernel is chain 0 only


def test_lm_alternating_problem_sizes_keep_the_cholesky_shared_memory_optin():
    """A small problem solved between two solves of a live larger one must not lower the dynamic shared-memory opt-in
    of the persistent Cholesky kernel (n = 480 needs ~168 KB; the attribute is per kernel and device, not per problem)."""
    from pycamset_b200 import synthetic as syn
    from pycamset_b200.problem import BundleProblem

    def make(C, M, seed):
        rig = syn.make_rig(C, M, distortion=True, seed=seed, detect_prob=0.9)
        intr, extr, poses = rig.perturbed(np.random.default_rng(seed), 1e-3)
        params = rig.param_string(intr, extr, poses)
        unfixed = np.ones(params.shape[0], bool)
        unfixed[15 * C:15 * C + 6] = False
        p = BundleProblem(0, rig.cam.numpy(), rig.pose.numpy(), rig.key.numpy(), rig.uv.numpy(), C, M, 81,
                          template=rig.template, unfixed=unfixed)
        p.set_param_string(params)
        return p, params[unfixed]

    big, xb = make(32, 6, 1)
    small, xs = make(4, 6, 2)
    try:
        for _ in range(2):
            _, st = big.lm_solve(xb, max_iter=3)
            assert st["cost_final"] < st["cost_initial"]
            _, st = small.lm_solve(xs, max_iter=3)
            assert st["cost_final"] < st["cost_initial"]
    finally:
        big.close(); small.close()
