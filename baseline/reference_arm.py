"""The UNMODIFIED reference (pyCamSet, installed as a writable copy under baseline/_ref) as bench.py's CPU arm.

Nothing here belongs to the product path: it imports the reference package, builds the reference's OWN handler objects
(TemplateBundleHandler / SelfBundleHandler, template_handler.py:80-193, standard_bundle_handler.py:100-226) around the
same synthetic rigs / golden fixtures the GPU arm uses, and times the reference's OWN closures and solver:

    loss_fun(x), jac_fn(x)       numba prange over `threads` chunks (abstract_function_blocks.py:351-387, :552-652)
    J.T @ J, J.T @ r             scipy.sparse, what a normal-equation solver on top of the reference would pay
    run_bundle_adjustment        scipy TRF + LSMR (optimisation_handling.py:52-117)

baseline/_ref is git-ignored but travels with the repository snapshot to the GPU box; the plotting / IO packages the
reference imports at package import time and that are absent from the image are stubbed in baseline/_ref/stubs
(SURVEY.md App. D).  If the import fails, `import_reference` raises and the caller reports the exception text.
"""
from __future__ import annotations

import os
import sys
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF = HERE / "_ref"


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def pin_thread_env(n: int | None = None) -> int:
    """Set the thread-count variables BEFORE numba / OpenMP initialise, so that a launcher (torchrun sets
    OMP_NUM_THREADS=1 for its children) cannot shrink the CPU arm."""
    n = n or host_threads()
    for k in ("OMP_NUM_THREADS", "NUMBA_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = str(n)
    return n


def import_reference():
    """Import pyCamSet from baseline/_ref.  Returns the package; raises ImportError with the reason otherwise."""
    if not (REF / "pyCamSet").is_dir():
        raise ImportError(f"{REF / 'pyCamSet'} is missing (the reference install is git-ignored; it is created in the build "
                          "container by the recipe in DESIGN.md and travels with the gpurun snapshot)")
    for q in (str(REF / "stubs"), str(REF)):
        if q not in sys.path:
            sys.path.insert(0, q)
    import pyCamSet
    if "_ref" not in str(Path(pyCamSet.__file__).resolve()):
        raise ImportError(f"pyCamSet resolved to {pyCamSet.__file__}, not to baseline/_ref")
    return pyCamSet


def ring_handlers(rig, layout: str, x_template=None, selfcal: bool = False):
    """Reference handler around a synthetic rig (pycamset_b200.synthetic.SyntheticRig), the way
    tests/golden/make_golden.py builds its goldens: CameraSet of default Cameras at the rig's extrinsics, ChArUco(10,10,4),
    TargetDetection over rig.dd(), initial parameters injected with set_initial_params (the OpenCV initialiser is
    by-passed: it is not on the measured path)."""
    import_reference()
    from pyCamSet import Camera, CameraSet, ChArUco
    from pyCamSet.calibration_targets import TargetDetection
    from pyCamSet.optimisation.standard_bundle_handler import SelfBundleHandler
    from pyCamSet.optimisation.template_handler import TemplateBundleHandler
    from pyCamSet.utils.general_utils import make_4x4h_tform

    C, M = rig.n_cams, rig.n_poses
    tforms = [make_4x4h_tform(rig.extr[b, :3], rig.extr[b, 3:]) for b in range(C)]
    cams = CameraSet(camera_dict={f"cam_{i:04d}": Camera(extrinsic=t) for i, t in enumerate(tforms)})
    target = ChArUco(10, 10, 4)
    det = TargetDetection(cam_names=cams.get_names(), data=rig.dd(), max_ims=M)
    opts = {"outliers": "n", "verbosity": 0}
    cls = SelfBundleHandler if selfcal else TemplateBundleHandler
    h = cls(cams, target, det, options=opts)
    h.missing_poses = np.zeros(M, bool)
    if x_template is not None:
        h.set_initial_params(np.asarray(x_template, np.float64).copy())
    return h


def golden_handler(g: dict, max_nfev: int = 100):
    """Reference handler around a committed golden case (tests/golden/ccube_*.npz: configs 2 / 3 -- the reference's own
    Ccube fixture, whose images are not shipped to the GPU box): the real Ccube target object, a TargetDetection rebuilt
    from the flattened observation table, initial parameters = the golden's x."""
    import_reference()
    import cv2
    from pyCamSet import Camera, CameraSet, Ccube
    from pyCamSet.calibration_targets import TargetDetection
    from pyCamSet.optimisation.standard_bundle_handler import SelfBundleHandler
    from pyCamSet.optimisation.template_handler import TemplateBundleHandler

    target = Ccube(n_points=10, length=40, aruco_dict=cv2.aruco.DICT_6X6_1000, border_fraction=0.2)
    tshape = target.point_data.shape[:-1]
    tmpl = np.asarray(target.point_data, np.float64).reshape(-1, 3)
    if tmpl.shape != g["template"].shape or np.max(np.abs(tmpl - g["template"])) > 1e-9:
        raise RuntimeError("Ccube target differs from the golden's template")
    dd = np.asarray(g["dd"], np.float64)
    keys = np.stack(np.unravel_index(dd[:, 2].astype(np.int64), tshape), axis=1).astype(np.float64)
    data = np.concatenate([dd[:, :2], keys, dd[:, 3:5]], axis=1)
    C, M = int(g["n_cams"]), int(g["n_poses"])
    cams = CameraSet(camera_dict={f"cam_{i:04d}": Camera() for i in range(C)})
    det = TargetDetection(cam_names=cams.get_names(), data=data, max_ims=M)
    opts = {"outliers": "n", "verbosity": 0, "max_nfev": int(max_nfev)}
    cls = SelfBundleHandler if int(g["chain"]) == 1 else TemplateBundleHandler
    h = cls(cams, target, det, options=opts)
    h.missing_poses = np.zeros(M, bool)
    h.set_initial_params(np.asarray(g["x"], np.float64).copy())
    return h


def time_callbacks(handler, x, threads: int, repeats: int = 5):
    """Best-of-`repeats` wall time of the reference's closures at x after a JIT warm-up call:
    dict(loss_s, jac_s, jtj_s, jtr_s, total_s, n_obs, nnz, build_s)."""
    import_reference()
    from pyCamSet.optimisation.optimisation_handling import make_optimisation_function
    t0 = time.perf_counter()
    loss, jac, x0 = make_optimisation_function(handler, threads)
    x = np.asarray(x0 if x is None else x, np.float64)
    r = loss(x)            # JIT warm-up (numba cache=True: compiled once per install)
    J = jac(x)
    build_s = time.perf_counter() - t0
    best = dict(loss_s=np.inf, jac_s=np.inf, jtj_s=np.inf, jtr_s=np.inf)
    for _ in range(repeats):
        t = time.perf_counter(); r = loss(x); best["loss_s"] = min(best["loss_s"], time.perf_counter() - t)
        t = time.perf_counter(); J = jac(x); best["jac_s"] = min(best["jac_s"], time.perf_counter() - t)
        t = time.perf_counter(); JtJ = J.T @ J; best["jtj_s"] = min(best["jtj_s"], time.perf_counter() - t)
        t = time.perf_counter(); Jtr = J.T @ r; best["jtr_s"] = min(best["jtr_s"], time.perf_counter() - t)
    best["total_s"] = best["loss_s"] + best["jac_s"] + best["jtj_s"] + best["jtr_s"]
    best.update(n_obs=int(r.shape[0] // 2), nnz=int(J.nnz), build_s=build_s)
    return best, (loss, jac)


def one_pass(loss, jac, x):
    """One evaluation of the metric's unit of work with the reference's own closures."""
    r = loss(x)
    J = jac(x)
    return J.T @ J, J.T @ r


def run_ba(handler, threads: int):
    """The reference's own run_bundle_adjustment; returns (result, seconds)."""
    import_reference()
    from pyCamSet.optimisation.optimisation_handling import run_bundle_adjustment
    t = time.perf_counter()
    res, _ = run_bundle_adjustment(param_handler=handler, threads=threads)
    return res, time.perf_counter() - t
