"""CPU: the oracle restatement against golden vectors produced by the reference itself
(tests/golden/make_golden.py).  Tolerances follow SURVEY.md 8d: r abs <= 1e-9 px, J rel <= 1e-9
(abs floor 1e-12), J^T J / J^T r rel <= 1e-9.  In practice agreement is ~1e-12 (reference is fastmath)."""
import numpy as np
import pytest
from scipy.sparse import csr_array

from oracle import oracle as orc
from tests.helpers import CCUBE_CASES, GOLDEN, SYNTH_CASES, available, load_case, oracle_problem, rel_err

ALL = available(SYNTH_CASES + CCUBE_CASES)


def test_goldens_present():
    assert set(SYNTH_CASES) <= set(ALL), "synthetic golden fixtures missing; run tests/golden/make_golden.py"
    assert (GOLDEN / "blocks.npz").exists()


def test_block_known_answers_survey_appendix_b():
    fun, jac = orc.block_projection([1200, 500, 1150, 480, -0.05, 0.02, 0.001, -0.001, 0.003], [0.01, -0.02, 0.25])
    assert np.allclose(fun, [547.959741513728, 388.067962098688], rtol=0, atol=1e-10)
    ref_x = [3.9966451261440002e-02, 1, 0, 0, 3.84e-01, 3.072e-03, -7.68, 13.44, 2.4576000000000003e-05,
             4.7954030754201603e+03, 2.6781519052800005, -1.9160187086438398e+02]
    ref_y = [0, 0, -7.9940902522879997e-02, 1, -7.36e-01, -5.888e-03, 23.92, -7.36, -4.7104000000000002e-05,
             2.5665622425600003, 4.5926647705804799e+03, 3.6731051915673601e+02]
    assert rel_err(jac, np.array([ref_x, ref_y])) < 1e-12
    fun, jac = orc.block_rigid([0.1, -0.2, 0.3, 0.01, 0.02, 0.03], [0.01, 0.02, 0.03])
    assert np.allclose(fun, [0.00788269146389451, 0.03802322471624366, 0.0627212526561976], rtol=0, atol=1e-15)
    assert abs(jac[0, 0] - 2.8720017095126681e-03) < 1e-15 and abs(jac[2, 8] - 0.97529030895304569) < 1e-15
    z = orc.block_rodrigues_jac([0, 0, 0])
    assert np.array_equal(z, [[0, 0, 0, 0, 0, -1, 0, 1, 0], [0, 0, 1, 0, 0, 0, -1, 0, 0], [0, -1, 0, 1, 0, 0, 0, 0, 0]])


def test_blocks_against_reference_blocks():
    g = dict(np.load(GOLDEN / "blocks.npz"))
    for i in range(g["q"].shape[0]):
        fun, jac = orc.block_projection(g["q"][i], g["X"][i])
        assert np.max(np.abs(fun - g["proj_fun"][i])) < 1e-9
        assert rel_err(jac.ravel(), g["proj_jac"][i]) < 1e-10
        fun, jac = orc.block_rigid(g["p6"][i], g["Y"][i])
        assert np.max(np.abs(fun - g["rigid_fun"][i])) < 1e-14
        assert np.max(np.abs(jac.ravel() - g["rigid_jac"][i])) < 1e-13
        assert np.max(np.abs(jac[:, :6].ravel() - g["template_jac"][i])) < 1e-13
        assert np.max(np.abs(orc.block_rodrigues_jac(g["p6"][i, :3]).ravel() - g["rodrigues_jac"][i])) < 1e-12


@pytest.mark.parametrize("case", ALL)
def test_residual_jacobian_normal_eq(case):
    g = load_case(case)
    p = oracle_problem(g)
    assert p.L == g["param0"].shape[0]
    # residual
    r = p.residual(g["param0"])
    assert np.max(np.abs(r - g["r"])) < 1e-9
    # CSR structure: identical integers
    fm = orc.free_map_from_mask(g["unfixed"])
    col, rp = p.csr_structure(fm)
    assert np.array_equal(rp, g["J_indptr"])
    assert np.array_equal(col, g["J_indices"].astype(np.int64))
    # CSR values
    vals = p.csr_values(g["param0"], fm, rp)
    assert rel_err(vals, g["J_data"]) < 1e-9
    # normal equations vs scipy on the reference CSR
    n_free = g["x"].shape[0]
    Jref = csr_array((g["J_data"], g["J_indices"], g["J_indptr"]), shape=(2 * p.N, n_free))
    JtJ_ref = (Jref.T @ Jref).toarray()
    JtJ, Jtr, cost = p.normal_dense(g["param0"], fm)
    scale = np.sqrt(np.outer(np.diag(JtJ_ref), np.diag(JtJ_ref))) + 1e-300
    assert np.max(np.abs(JtJ - JtJ_ref) / scale) < 1e-9
    assert np.max(np.abs(np.diag(JtJ) - g["JtJ_diag"]) / np.maximum(g["JtJ_diag"], 1e-300)) < 1e-9
    gscale = np.sqrt(np.diag(JtJ_ref) * cost) + 1e-300
    assert np.max(np.abs(Jtr - g["Jtr"]) / gscale) < 1e-9
    assert np.max(np.abs(JtJ @ g["probe"] - g["JtJ_probe"]) / (np.abs(g["JtJ_probe"]) + np.linalg.norm(g["JtJ_probe"]) * 1e-3)) < 1e-8
    assert abs(cost - float(g["r"] @ g["r"])) <= 1e-9 * cost


@pytest.mark.parametrize("case", [c for c in ALL if c.endswith("template")])
def test_block_normal_equations_match_dense(case):
    """U/V/W/g blocks (no parameter fixed) == the corresponding blocks of the dense J^T J."""
    g = load_case(case)
    p = oracle_problem(g)
    pair = p.cam.astype(np.int64) * p.M + p.pose
    uniq, seg = np.unique(pair, return_inverse=True)
    U, gc, V, gp, W, cost = p.normal_blocks(g["param0"], seg.astype(np.int32), len(uniq))
    fm = np.arange(p.L, dtype=np.int32)
    JtJ, Jtr, cost2 = p.normal_dense(g["param0"], fm)
    C, M = p.C, p.M
    tol = 1e-11
    for c in range(C):
        idx = np.r_[9 * c:9 * c + 9, 9 * C + 6 * c:9 * C + 6 * c + 6]
        ref = JtJ[np.ix_(idx, idx)]
        assert np.max(np.abs(U[c] - ref)) <= tol * max(1.0, np.abs(ref).max())
        assert np.max(np.abs(gc[c] - Jtr[idx])) <= tol * max(1.0, np.abs(Jtr[idx]).max())
    for m in range(M):
        idx = np.r_[15 * C + 6 * m:15 * C + 6 * m + 6]
        ref = JtJ[np.ix_(idx, idx)]
        assert np.max(np.abs(V[m] - ref)) <= tol * max(1.0, np.abs(ref).max())
        assert np.max(np.abs(gp[m] - Jtr[idx])) <= tol * max(1.0, np.abs(Jtr[idx]).max())
    for s, pr in enumerate(uniq):
        c, m = divmod(int(pr), M)
        ci = np.r_[9 * c:9 * c + 9, 9 * C + 6 * c:9 * C + 6 * c + 6]
        mi = np.r_[15 * C + 6 * m:15 * C + 6 * m + 6]
        ref = JtJ[np.ix_(ci, mi)]
        assert np.max(np.abs(W[s] - ref)) <= tol * max(1.0, np.abs(ref).max())
    assert abs(cost - cost2) <= 1e-12 * cost


def test_costfn_oracle_against_reference():
    """Initialiser cost evaluation: oracle restatement vs the reference's bundle_adjustment_costfn outputs."""
    g = dict(np.load(GOLDEN / "costfn.npz"))
    for b in range(g["tables"].shape[0]):
        e = orc.costfn(g["dd"], g["tables"][b], g["proj"], g["ints"], g["dists"])
        assert np.max(np.abs(e - g["errors"][b])) < 1e-9
        norms = np.sqrt(np.sum(e.reshape(-1, 2) ** 2, axis=1))
        pi = np.bincount(g["dd"][:, 1].astype(int), weights=norms, minlength=int(g["n_poses"]))
        assert np.max(np.abs(pi - g["per_image"][b]) / g["per_image"][b]) < 1e-12


@pytest.mark.parametrize("case,final", [("ccube_template", None), ("ccube_selfcal", "ccube_selfcal_final")])
def test_oracle_reproduces_the_reference_final_iterates(case, final):
    """The reference's own converged iterates (configs 2 / 3, max_nfev = 100): the oracle's residual at x_final gives the
    mean reprojection error the reference reported (2.633 px / 0.218 px; its tests threshold them at 5.10 / 0.50)."""
    from tests.helpers import GOLDEN, load_case, oracle_problem
    if not (GOLDEN / f"{case}.npz").exists():
        pytest.skip("golden missing")
    g = load_case(case)
    f = g if final is None else dict(np.load(GOLDEN / f"{final}.npz"))
    p = g["param0"].copy()
    p[g["unfixed"]] = f["x_final"]
    r = oracle_problem(g).residual(p)
    px = float(np.mean(np.linalg.norm(r.reshape(-1, 2), axis=1)))
    assert abs(px - float(f["final_px"])) < 1e-9
    assert px < (5.10 if case == "ccube_template" else 0.50)
