"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference (pyCamSet).

Run in the build container only (the reference does not travel to the GPU box):

    PYTHONSAFEPATH=1 PYTHONPATH=/root/repo:/root/repo/baseline/_ref:/root/repo/baseline/_ref/stubs \
        python tests/golden/make_golden.py [--skip-ccube]

`baseline/_ref/pyCamSet` is a writable, git-ignored copy of /root/reference/pyCamSet (the reference's code
generator and numba cache write into the package directory); `baseline/_ref/stubs` holds import stubs for the
plotting / IO packages that are absent in this image (pyvista, matplotlib, coloredlogs, blosc, uniplot) and a
natural-sort `natsort`.

What is dumped per case (all float64 unless noted):
    dd          N x 5 observation table [cam, img, key, u, v] handed to the reference's compiled chain
    template    K x 3 target points
    param0      full parameter string (fixed values included) the reference builds from x
    unfixed     boolean mask over the parameter string
    x           free-parameter vector the callbacks were evaluated at
    r           loss_fun(x)           (2N,)
    J_data/J_indices/J_indptr         jac_fn(x) as CSR
    JtJ_diag, Jtr, JtJ_probe          (J.T @ J).diagonal(), J.T @ r, and (J.T @ J) @ probe for a seeded probe
                                      vector (the full J.T @ J is re-derived from the CSR in the tests)
"""
from __future__ import annotations

import argparse
import os
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent


def dump_case(name, handler, x, extra=None):
    from pyCamSet.optimisation.optimisation_handling import make_optimisation_function

    handler.set_initial_params(np.asarray(x, np.float64).copy())
    loss, jac, x0 = make_optimisation_function(handler, threads=2)
    r = np.asarray(loss(x0)).copy()
    J = jac(x0)
    J.sort_indices() if False else None  # keep the reference's native (unsorted-within-row) order
    inps = handler.get_bundle_adjustment_inputs(x0)
    param0 = handler.op_fun.build_param_list(*inps).copy()
    dd = handler.get_detection_data(flatten=True)
    bp = handler.bundlePrimitive
    masks = [np.repeat(bp.intr_unfixed, 9), np.repeat(bp.extr_unfixed, 6), np.repeat(bp.poses_unfixed, 6)]
    selfcal = hasattr(bp, "bdpt_unfixed")
    if selfcal:
        masks.append(np.asarray(bp.bdpt_unfixed, bool))
    unfixed = np.concatenate(masks)
    JtJ = (J.T @ J).tocsr()
    rng = np.random.default_rng(12345)
    probe = rng.normal(size=J.shape[1])
    out = dict(
        dd=dd, template=handler.target.point_data.reshape(-1, 3).astype(np.float64), param0=param0,
        unfixed=unfixed, x=x0, r=r, J_data=J.data, J_indices=J.indices.astype(np.int32),
        J_indptr=J.indptr.astype(np.int64), JtJ_diag=JtJ.diagonal(), Jtr=J.T @ r, probe=probe,
        JtJ_probe=JtJ @ probe, chain=np.int32(1 if selfcal else 0),
        n_cams=np.int32(bp.intr.shape[0]), n_poses=np.int32(bp.poses.shape[0]),
    )
    if extra:
        out.update(extra)
    np.savez_compressed(HERE / f"{name}.npz", **out)
    print(f"[golden] {name}: N={dd.shape[0]} n_free={x0.shape[0]} nnz={J.nnz} "
          f"mean px={np.mean(np.linalg.norm(r.reshape(-1, 2), axis=1)):.4f}")


def synthetic_handlers(seed, n_cams, n_poses, detect_prob, fixed_cam_ext=False):
    """Reference handler objects around a synthetic ring (examples/make_camera_ring.py recipe)."""
    from pyCamSet import Camera, CameraSet, ChArUco
    from pyCamSet.calibration_targets import TargetDetection
    from pyCamSet.optimisation.template_handler import TemplateBundleHandler
    from pyCamSet.optimisation.standard_bundle_handler import SelfBundleHandler
    from pyCamSet.utils.general_utils import make_4x4h_tform
    from pycamset_b200 import synthetic as syn

    rig = syn.make_rig(n_cams, n_poses, layout="ring", distortion=True, seed=seed, detect_prob=detect_prob)
    tforms = [make_4x4h_tform((0, b / n_cams * 2 * np.pi, 0), (0, 0, 0.2)) for b in range(n_cams)]
    cams = CameraSet(camera_dict={f"cam_{i}": Camera(extrinsic=t) for i, t in enumerate(tforms)})
    target = ChArUco(10, 10, 4)
    assert np.abs(target.point_data - rig.template).max() < 1e-8
    rig.template = target.point_data.reshape(-1, 3).astype(np.float64)  # float32-rounded corners, as the reference sees them
    det = TargetDetection(cam_names=cams.get_names(), data=rig.dd(), max_ims=n_poses)
    rng = np.random.default_rng(seed + 100)
    intr, extr, poses = rig.perturbed(rng, 1e-3)
    fixed = None
    if fixed_cam_ext:
        # camera 1: extrinsic and intrinsic both held fixed (template_handler.py:112-132, :204-213)
        fixed = {"cam_1": {"ext": extr[1].copy(), "int": intr[1].copy()}}
    opts = {"outliers": "n", "verbosity": 0}
    th = TemplateBundleHandler(cams, target, det, fixed_params=fixed, options=opts)
    th.missing_poses = np.zeros(n_poses, bool)
    bp = th.bundlePrimitive
    x_t = np.concatenate([intr[bp.intr_unfixed].ravel(), extr[bp.extr_unfixed].ravel(),
                          poses[bp.poses_unfixed].ravel()])
    sh = SelfBundleHandler(cams, target, det, fixed_params=fixed, options=opts)
    sh.missing_poses = np.zeros(n_poses, bool)
    pts = target.point_data.reshape(-1).astype(np.float64)
    pts = pts + 1e-5 * rng.normal(size=pts.shape)
    x_s = np.concatenate([x_t, pts[sh.feat_unfixed]])
    return rig, th, x_t, sh, x_s


def block_goldens():
    """Block-level answers straight from the reference's numba blocks (function_block_implementations.py)."""
    from pyCamSet.optimisation import function_block_implementations as fb
    from pyCamSet.optimisation.compiled_helpers import numba_rodrigues_jac

    rng = np.random.default_rng(7)
    n = 16
    q = np.stack([rng.uniform(900, 1300, n), 500 + rng.normal(0, 10, n), rng.uniform(900, 1300, n),
                  500 + rng.normal(0, 10, n), rng.normal(0, 0.05, n), rng.normal(0, 0.02, n),
                  rng.normal(0, 1e-3, n), rng.normal(0, 1e-3, n), rng.normal(0, 1e-2, n)], 1)
    X = np.stack([rng.normal(0, 0.03, n), rng.normal(0, 0.03, n), rng.uniform(0.15, 0.4, n)], 1)
    p6 = np.concatenate([rng.normal(0, 0.5, (n, 3)), rng.normal(0, 0.05, (n, 3))], 1)
    p6[0, :3] = 0.0          # small-angle branch (theta < 1e-10), hit by every observation of fixed pose 0
    p6[1, :3] = [3e-11, 0, 0]
    Y = rng.normal(0, 0.05, (n, 3))
    pf = np.empty((n, 2)); pj = np.empty((n, 24)); rf = np.empty((n, 3)); rj = np.empty((n, 27))
    tj = np.empty((n, 18)); dr = np.empty((n, 27))
    for i in range(n):
        fb.projection.compute_fun(q[i], X[i], pf[i], np.empty(1))
        fb.projection.compute_jac(q[i], X[i], pj[i], np.empty(1))
        out = np.empty(27)
        fb.rigidTform3d.compute_fun(p6[i], Y[i], out, np.empty(27)); rf[i] = out[:3]
        fb.rigidTform3d.compute_jac(p6[i], Y[i], rj[i], np.empty(27))
        fb.template_points.compute_jac(p6[i], Y[i], tj[i], np.empty(27))
        numba_rodrigues_jac(p6[i, :3].copy(), dr[i])
    np.savez_compressed(HERE / "blocks.npz", q=q, X=X, p6=p6, Y=Y, proj_fun=pf, proj_jac=pj, rigid_fun=rf,
                        rigid_jac=rj, template_jac=tj, rodrigues_jac=dr)
    print("[golden] blocks: 16 random points per block")


def costfn_golden():
    """Initialiser cost evaluation: estimate_camera_relative_poses evaluates
    ch.bundle_adjustment_costfn once per candidate pose table (template_handler.py:510-593,
    compiled_helpers.py:517-549).  Dumps its inputs and outputs for a seeded ring with B candidate tables, plus the
    per-image sums of the per-observation error norms the initialiser reduces them to (:550-560)."""
    from pyCamSet.optimisation import compiled_helpers as ch
    from pyCamSet.utils.general_utils import h_tform, make_4x4h_tform
    from pycamset_b200 import synthetic as syn

    C, M, B = 5, 7, 4
    rig = syn.make_rig(C, M, layout="ring", distortion=True, seed=17, detect_prob=0.8)
    rng = np.random.default_rng(99)
    dd = rig.dd()
    ints = np.zeros((C, 3, 3)); dists = np.zeros((C, 5)); proj = np.zeros((C, 3, 4))
    for c in range(C):
        fx, px, fy, py, k1, k2, p1, p2, k3 = rig.intr[c]
        ints[c] = [[fx, 0, px], [0, fy, py], [0, 0, 1]]
        dists[c] = [k1, k2, p1, p2, k3]
        proj[c] = ints[c] @ make_4x4h_tform(rig.extr[c, :3], rig.extr[c, 3:])[:3, :]
    tables = np.empty((B, M, rig.template.shape[0], 3))
    for b in range(B):   # candidate pose tables: the truth and perturbed copies (one per reference camera in the initialiser)
        for m in range(M):
            pose = rig.poses[m] + (0 if b == 0 else 1e-2 * rng.normal(size=6))
            tables[b, m] = h_tform(rig.template, make_4x4h_tform(pose[:3], pose[3:]))
    errors = np.stack([ch.bundle_adjustment_costfn(dd, tables[b], proj, ints, dists) for b in range(B)])
    norms = np.sqrt(np.sum(errors.reshape(B, -1, 2) ** 2, axis=2))
    per_image = np.stack([[np.sum(norms[b][dd[:, 1] == m]) for m in range(M)] for b in range(B)])
    np.savez_compressed(HERE / "costfn.npz", dd=dd, tables=tables, proj=proj, ints=ints, dists=dists, errors=errors,
                        per_image=per_image, n_cams=np.int32(C), n_poses=np.int32(M))
    print(f"[golden] costfn: N={dd.shape[0]} B={B} mean |e| = {norms.mean():.3f} px")


def ccube_goldens():
    """Configs 2 and 3: the reference's own test paths (tests/calibrate_ccube_test.py:6-19,
    tests/self_calibrate_ccube_test.py:10-37), dumped at the initial and at the final iterate."""
    from pyCamSet import calibrate_cameras, Ccube
    from pyCamSet.optimisation.standard_bundle_handler import SelfBundleHandler
    from pyCamSet.optimisation.optimisation_handling import run_bundle_adjustment
    import cv2

    loc = Path("/root/reference/tests/test_data/calibration_ccube")
    target = Ccube(n_points=10, length=40, aruco_dict=cv2.aruco.DICT_6X6_1000, border_fraction=0.2)
    cams = calibrate_cameras(f_loc=loc, calibration_target=target, draw=False, save=False,
                             problem_options={"outliers": "n", "verbosity": 0})
    h = cams.calibration_handler
    x_final = np.asarray(cams.calibration_params).copy()
    x_init = np.asarray(h.get_initial_params()).copy()
    final_px = float(np.mean(np.linalg.norm(np.reshape(cams.calibration_result_fun
                                                       if hasattr(cams, "calibration_result_fun") else
                                                       h.make_loss_fun(2)(x_final), (-1, 2)), axis=1)))
    dump_case("ccube_template", h, x_init, extra=dict(x_final=x_final, final_px=np.float64(final_px)))

    sh = SelfBundleHandler(detection=h.detection, target=target, camset=cams,
                           options={"outliers": "n", "verbosity": 0, "max_nfev": 100})
    sh.set_from_templated_camset(cams)
    x_s = np.asarray(sh.get_initial_params()).copy()
    dump_case("ccube_selfcal", sh, x_s, extra=dict(fixed_inds=np.asarray(sh.fixed_inds, np.int32)))


def ccube_selfcal_final():
    """Config 3 at the reference's FINAL iterate: tests/self_calibrate_ccube_test.py:23-37 verbatim (SelfBundleHandler
    from the template calibration, max_nfev 100, run_bundle_adjustment), dumping the solver's x, cost, nfev and the mean
    reprojection error the reference's own test thresholds (< 0.50 px).  Written to ccube_selfcal_final.npz; the
    start point is asserted to be the `x` of ccube_selfcal.npz."""
    from pyCamSet import calibrate_cameras, Ccube
    from pyCamSet.optimisation.standard_bundle_handler import SelfBundleHandler
    from pyCamSet.optimisation.optimisation_handling import run_bundle_adjustment
    import cv2
    import time

    loc = Path("/root/reference/tests/test_data/calibration_ccube")
    target = Ccube(n_points=10, length=40, aruco_dict=cv2.aruco.DICT_6X6_1000, border_fraction=0.2)
    cams = calibrate_cameras(f_loc=loc, calibration_target=target, draw=False, save=False,
                             problem_options={"outliers": "n", "verbosity": 0})
    sh = SelfBundleHandler(detection=cams.calibration_handler.detection, target=target, camset=cams,
                           options={"outliers": "n", "verbosity": 0, "max_nfev": 100})
    sh.set_from_templated_camset(cams)
    x0 = np.asarray(sh.get_initial_params()).copy()
    g = np.load(HERE / "ccube_selfcal.npz")
    assert np.max(np.abs(x0 - g["x"])) < 1e-9, "start point differs from the committed golden"
    t0 = time.time()
    res, _ = run_bundle_adjustment(param_handler=sh, threads=os.cpu_count())
    dt = time.time() - t0
    px = float(np.mean(np.linalg.norm(np.reshape(res.fun, (-1, 2)), axis=1)))
    np.savez_compressed(HERE / "ccube_selfcal_final.npz", x_final=res.x, final_px=np.float64(px), cost_final=np.float64(res.cost),
                        nfev=np.int32(res.nfev), status=np.int32(res.status), seconds=np.float64(dt))
    print(f"[golden] ccube_selfcal_final: nfev={res.nfev} status={res.status} cost={res.cost:.6f} px={px:.6f} ({dt:.1f} s)")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-ccube", action="store_true")
    ap.add_argument("--only-costfn", action="store_true")
    ap.add_argument("--selfcal-final", action="store_true", help="only (re)generate ccube_selfcal_final.npz")
    args = ap.parse_args()
    import pyCamSet
    assert "baseline/_ref" in pyCamSet.__file__, pyCamSet.__file__
    if args.selfcal_final:
        ccube_selfcal_final()
        return 0
    costfn_golden()
    if args.only_costfn:
        return 0
    block_goldens()
    rig, th, x_t, sh, x_s = synthetic_handlers(seed=3, n_cams=4, n_poses=6, detect_prob=0.8)
    dump_case("ring4_template", th, x_t)
    dump_case("ring4_selfcal", sh, x_s)
    rig, th, x_t, sh, x_s = synthetic_handlers(seed=5, n_cams=5, n_poses=7, detect_prob=0.6, fixed_cam_ext=True)
    dump_case("ring5_fixedcam_template", th, x_t)
    dump_case("ring5_fixedcam_selfcal", sh, x_s)
    if not args.skip_ccube:
        ccube_goldens()


if __name__ == "__main__":
    sys.exit(main())
