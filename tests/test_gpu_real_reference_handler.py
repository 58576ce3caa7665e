"""GPU: the REAL reference handler objects (pyCamSet's TemplateBundleHandler / SelfBundleHandler, imported from the
git-ignored install under baseline/_ref that travels with the repository snapshot) driven through the CUDA closures.

The other handler tests use duck-typed stand-ins (tests/fake_reference.py) so that they run without the reference; this
one closes the loop: the same handler instance is handed (i) to the reference's own make_optimisation_function and
(ii) to pycamset_b200.handler.make_optimisation_function, and the two pairs of closures must agree -- residual abs
<= 1e-9 px, CSR structure bit-exact, Jacobian entries rel <= 1e-9 -- followed by the device LM through the drop-in
run_bundle_adjustment, whose `camset` comes from the reference's own get_camset.  Skipped when the reference cannot be
imported on the box (the reason is printed)."""
import numpy as np
import pytest

from tests.helpers import load_case, rel_err

pytestmark = pytest.mark.gpu


def _reference():
    try:
        from baseline import reference_arm as ra
        ra.pin_thread_env()
        ra.import_reference()
        return ra
    except Exception as e:  # pragma: no cover
        pytest.skip(f"reference not importable here: {e!r}")


@pytest.mark.parametrize("case", ["ccube_template", "ccube_selfcal"])
def test_real_reference_handler_through_the_cuda_closures(case):
    ra = _reference()
    from pyCamSet.optimisation.optimisation_handling import make_optimisation_function as ref_make
    from pycamset_b200.handler import GpuBundleHandler, make_optimisation_function, run_bundle_adjustment
    g = load_case(case)
    h = ra.golden_handler(g, max_nfev=30)
    loss_r, jac_r, x0 = ref_make(h, ra.host_threads())
    gpu = GpuBundleHandler(h)
    assert gpu.stock_mapping and gpu.chain == g["chain"]
    loss_g, jac_g, x0_g = make_optimisation_function(gpu, 1)
    assert np.array_equal(x0, x0_g)
    r_r, r_g = loss_r(x0), loss_g(x0)
    assert r_r.shape == r_g.shape and np.max(np.abs(r_r - r_g)) < 1e-9
    J_r, J_g = jac_r(x0), jac_g(x0)
    assert J_r.shape == J_g.shape
    assert np.array_equal(J_r.indptr, J_g.indptr) and np.array_equal(J_r.indices, J_g.indices)
    assert rel_err(J_g.data, J_r.data) < 1e-9
    res, camset = run_bundle_adjustment(gpu, solver="lm")
    px0 = np.mean(np.linalg.norm(r_r.reshape(-1, 2), axis=1))
    px1 = np.mean(np.linalg.norm(res.fun.reshape(-1, 2), axis=1))
    assert px1 < px0 and px1 < (5.10 if case == "ccube_template" else 0.50)
    assert camset is not None and camset.get_n_cams() == int(g["n_cams"])      # built by the reference's own get_camset
    assert np.max(np.abs(loss_r(res.x) - res.fun)) < 1e-9                    # the reference agrees on the final residual
    gpu.close()


def test_initialiser_drop_in_matches_the_reference():
    """estimate_camera_relative_poses (template_handler.py:468-601): the reference's own function and the drop-in
    (pycamset_b200.initialiser, cost evaluation on the GPU) on the same target / detection / camera objects."""
    ra = _reference()
    from pyCamSet import Camera, CameraSet, ChArUco
    from pyCamSet.calibration_targets import TargetDetection
    from pyCamSet.optimisation.template_handler import estimate_camera_relative_poses as ref_fn
    from pyCamSet.utils.general_utils import make_4x4h_tform
    from pycamset_b200 import synthetic as syn
    from pycamset_b200.initialiser import estimate_camera_relative_poses as gpu_fn
    C, M = 5, 9
    rig = syn.make_rig(C, M, layout="dome", distortion=False, seed=3, detect_prob=0.9)
    cams = CameraSet(camera_dict={f"cam_{i}": Camera(extrinsic=make_4x4h_tform(rig.extr[i, :3], rig.extr[i, 3:])) for i in range(C)})
    for name in cams.get_names():
        cams[name].name = name
    target = ChArUco(10, 10, 4)
    det = TargetDetection(cam_names=cams.get_names(), data=rig.dd(), max_ims=M)
    a = ref_fn(target, det, cams)
    b = gpu_fn(target, det, cams)
    assert [x.shape for x in a] == [x.shape for x in b] and a[2].shape == (2 * M,)
    assert np.max(np.abs(a[0] - b[0])) < 1e-12 and np.max(np.abs(a[1] - b[1])) < 1e-12
    assert np.max(np.abs(a[2] - b[2]) / np.maximum(a[2], 1.0)) < 1e-9


def test_gauge_transform_drop_in_matches_the_reference():
    """SelfBundleHandler.apply_gauge_transform (standard_bundle_handler.py:339-410) on the ccube self-calibration fixture at
    the reference's own final iterate: scale, points, poses and extrinsics from pycamset_b200.gauge (pair search on the GPU)
    against the reference's method on copies of the same arrays."""
    ra = _reference()
    from pycamset_b200 import gauge
    from tests.helpers import GOLDEN
    g = load_case("ccube_selfcal")
    h = ra.golden_handler(g)
    x = np.load(GOLDEN / "ccube_selfcal_final.npz")["x_final"]
    model = [np.array(a, np.float64).copy() for a in h.bundlePrimitive.return_bundle_primitives(x)]
    ref = h.apply_gauge_transform(*[a.copy() for a in model])
    got = gauge.apply_gauge_transform_for(h, *[a.copy() for a in model])
    for a, b, name in zip(ref, got, ("proj", "extr", "poses", "points")):
        assert np.all(np.isfinite(a)), name
        assert np.max(np.abs(np.asarray(a) - np.asarray(b))) < 1e-9, name
    # the transform preserves the calibration: residuals before and after agree
    assert abs(np.linalg.norm(ref[3]) - np.linalg.norm(model[3])) > 0      # it did move the points
