"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the committed golden vectors.

Tolerances: residual abs <= 1e-9 px; Jacobian entries rel <= 1e-9 (abs floor 1e-12);
J^T J / J^T r rel <= 1e-9 per block (scaled by the block's diagonal); CSR structure bit-exact.
"""
import numpy as np
import pytest

from oracle import oracle as orc
from tests.helpers import CCUBE_CASES, SYNTH_CASES, available, load_case, oracle_problem, rel_err

pytestmark = pytest.mark.gpu
ALL = available(SYNTH_CASES + CCUBE_CASES)


def gpu_problem(g):
    from pycamset_b200.problem import BundleProblem
    dd = g["dd"]
    K = g["template"].shape[0]
    return BundleProblem(g["chain"], dd[:, 0], dd[:, 1], dd[:, 2], dd[:, 3:5], int(g["n_cams"]), int(g["n_poses"]), K,
                         template=g["template"] if g["chain"] == 0 else None, unfixed=g["unfixed"])


@pytest.mark.parametrize("case", ALL)
def test_residual_and_jacobian(case):
    g = load_case(case)
    with gpu_problem(g) as p:
        assert p.n_free == g["x"].shape[0]
        assert p.nnz == g["J_data"].shape[0]
        p.set_param_string(g["param0"])
        r = p.residual(g["x"])
        assert np.max(np.abs(r - g["r"])) < 1e-9            # vs the reference itself
        o = oracle_problem(g)
        assert np.max(np.abs(r - o.residual(g["param0"]))) < 1e-9
        col, rp = p.csr_structure()
        assert np.array_equal(rp, g["J_indptr"])
        assert np.array_equal(col, g["J_indices"].astype(np.int64))
        vals = p.jacobian_values(g["x"])
        assert rel_err(vals, g["J_data"]) < 1e-9
        # parameters scatter: x -> parameter string round trip
        assert np.array_equal(p.get_param_string(), g["param0"])


@pytest.mark.parametrize("case", ALL)
def test_dense_normal_equations(case):
    g = load_case(case)
    if g["x"].shape[0] > 4000:
        pytest.skip("dense path is for small problems")
    with gpu_problem(g) as p:
        p.set_param_string(g["param0"])
        JtJ, Jtr, cost = p.normal_dense(g["x"])
    o = oracle_problem(g)
    JtJ_o, Jtr_o, cost_o = o.normal_dense(g["param0"], orc.free_map_from_mask(g["unfixed"]))
    d = np.sqrt(np.maximum(np.diag(JtJ_o), 1e-300))
    assert np.max(np.abs(JtJ - JtJ_o) / np.outer(d, d)) < 1e-9
    assert np.max(np.abs(Jtr - Jtr_o) / (d * np.sqrt(cost_o))) < 1e-9
    assert abs(cost - cost_o) <= 1e-11 * cost_o
    assert np.max(np.abs(Jtr - g["Jtr"]) / (d * np.sqrt(cost_o))) < 1e-9   # vs scipy on the reference CSR


@pytest.mark.parametrize("case", [c for c in ALL if c.endswith("template")])
def test_block_normal_equations(case):
    g = load_case(case)
    o = oracle_problem(g)
    with gpu_problem(g) as p:
        p.set_param_string(g["param0"])
        ne = p.normal_equations(g["x"])
        sc, sp, sl = p.segments()
    # segment table: sorted unique (camera, pose) pairs with their observation counts
    pair = o.cam.astype(np.int64) * o.M + o.pose
    uniq, seg, cnt = np.unique(pair, return_inverse=True, return_counts=True)
    assert np.array_equal(sc.astype(np.int64) * o.M + sp, uniq)
    assert np.array_equal(sl, cnt)
    U, gc, V, gp, W, cost = o.normal_blocks(g["param0"], seg.astype(np.int32), len(uniq))

    def blk_close(a, b, da, db, tol=1e-9):
        scale = np.sqrt(np.maximum(da, 1e-300))[..., :, None] * np.sqrt(np.maximum(db, 1e-300))[..., None, :]
        return np.max(np.abs(a - b) / scale) < tol

    dU = np.einsum("cii->ci", U); dV = np.einsum("mii->mi", V)
    assert blk_close(ne["U"], U, dU, dU)
    assert blk_close(ne["V"], V, dV, dV)
    assert blk_close(ne["W"], W, dU[sc], dV[sp])
    assert np.max(np.abs(ne["gc"] - gc) / np.sqrt(np.maximum(dU, 1e-300) * cost)) < 1e-9
    assert np.max(np.abs(ne["gp"] - gp) / np.sqrt(np.maximum(dV, 1e-300) * cost)) < 1e-9
    assert abs(ne["cost"] - cost) <= 1e-11 * cost
    assert abs(ne["cost"] - float(g["r"] @ g["r"])) <= 1e-9 * cost


def test_synthetic_medium_vs_oracle():
    """Seeded 8-camera ring x 40 poses, ragged visibility: every output against the oracle on the same inputs."""
    from pycamset_b200 import synthetic as syn
    from pycamset_b200.problem import BundleProblem
    rig = syn.make_rig(8, 40, distortion=True, seed=11, detect_prob=0.7)
    rng = np.random.default_rng(1)
    intr, extr, poses = rig.perturbed(rng)
    params = rig.param_string(intr, extr, poses)
    unfixed = np.ones(params.shape[0], bool)
    unfixed[15 * 8:15 * 8 + 6] = False  # pose 0 fixed
    o = orc.Problem(0, rig.cam.numpy(), rig.pose.numpy(), rig.key.numpy(), rig.uv.numpy(), 8, 40, 81, rig.template)
    with BundleProblem(0, rig.cam.numpy(), rig.pose.numpy(), rig.key.numpy(), rig.uv.numpy(), 8, 40, 81,
                       template=rig.template, unfixed=unfixed) as p:
        p.set_param_string(params)
        r = p.residual()
        assert np.max(np.abs(r - o.residual(params))) < 1e-9
        fm = orc.free_map_from_mask(unfixed)
        col_o, rp_o = o.csr_structure(fm)
        col, rp = p.csr_structure()
        assert np.array_equal(col, col_o) and np.array_equal(rp, rp_o)
        assert rel_err(p.jacobian_values(), o.csr_values(params, fm, rp_o)) < 1e-9
        ne = p.normal_equations()
        sc, sp, sl = p.segments()
    pair = o.cam.astype(np.int64) * o.M + o.pose
    uniq, seg = np.unique(pair, return_inverse=True)
    U, gc, V, gp, W, cost = o.normal_blocks(params, seg.astype(np.int32), len(uniq))
    for a, b in ((ne["U"], U), (ne["V"], V), (ne["W"], W), (ne["gc"], gc), (ne["gp"], gp)):
        assert np.max(np.abs(a - b)) <= 1e-10 * np.max(np.abs(b))
    assert abs(ne["cost"] - cost) <= 1e-11 * cost


def test_lm_converges_template():
    """LM on the device recovers the noise-level optimum from a 1e-3 relative perturbation."""
    from pycamset_b200 import synthetic as syn
    from pycamset_b200.problem import BundleProblem
    rig = syn.make_rig(8, 30, distortion=True, seed=21, detect_prob=0.9)
    rng = np.random.default_rng(2)
    intr, extr, poses = rig.perturbed(rng)
    params = rig.param_string(intr, extr, poses)
    unfixed = np.ones(params.shape[0], bool)
    unfixed[15 * 8:15 * 8 + 6] = False
    with BundleProblem(0, rig.cam.numpy(), rig.pose.numpy(), rig.key.numpy(), rig.uv.numpy(), 8, 30, 81,
                       template=rig.template, unfixed=unfixed) as p:
        p.set_param_string(params)
        x0 = params[unfixed]
        r0 = p.residual(x0)
        x, st = p.lm_solve(x0, max_iter=60, ftol=1e-12, xtol=1e-12, gtol=1e-10)
        r1 = p.residual(x)
    px0 = np.mean(np.linalg.norm(r0.reshape(-1, 2), axis=1))
    px1 = np.mean(np.linalg.norm(r1.reshape(-1, 2), axis=1))
    assert st["status"] >= 0 and st["cost_final"] < st["cost_initial"]
    assert abs(0.5 * float(r1 @ r1) - st["cost_final"]) <= 1e-9 * st["cost_final"]
    assert px1 < 0.14 < px0, (px0, px1)   # noise is N(0, 0.1 px) per coordinate -> mean norm ~0.125 px
