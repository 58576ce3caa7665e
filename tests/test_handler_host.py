"""CPU: host-side mirror of the reference interface (pycamset_b200.handler) -- export logic only, no device calls.

When the writable reference copy is present in this container (baseline/_ref, see tests/golden/make_golden.py) the
export is also checked on a REAL pyCamSet handler; that part is skipped elsewhere (the reference does not travel)."""
import numpy as np
import pytest

from tests import fake_reference as fr
from tests.helpers import available, load_case, SYNTH_CASES, CCUBE_CASES

ALL = available(SYNTH_CASES + CCUBE_CASES)


@pytest.mark.parametrize("case", ALL)
def test_export_matches_golden(case):
    from pycamset_b200 import handler as H
    g = load_case(case)
    h = (fr.SelfBundleHandler if g["chain"] == 1 else fr.TemplateBundleHandler)(g)
    e = H.export_problem(h)
    assert e["blocks"] == (("projection", "extrinsic3D", "rigidTform3d", "free_point") if g["chain"] == 1 else
                           ("projection", "extrinsic3D", "template_points"))
    assert np.array_equal(e["unfixed"], g["unfixed"])
    assert np.array_equal(e["dd"], g["dd"])
    assert e["stock_mapping"]
    assert (e["n_cams"], e["n_poses"]) == (int(g["n_cams"]), int(g["n_poses"]))
    # x -> parameter string through the handler reproduces the reference's string
    ps = h.op_fun.build_param_list(*h.get_bundle_adjustment_inputs(g["x"]))
    assert np.array_equal(ps, g["param0"])


def test_subclass_mapping_is_detected():
    from pycamset_b200 import handler as H
    g = load_case(ALL[0])
    assert not H.export_problem(fr.FocalInKiloPixels(g))["stock_mapping"]
    assert H.export_problem(fr.UnknownChainHandler(g))["blocks"][-1] == "my_custom_block"


def test_optimize_result_access():
    from pycamset_b200.handler import OptimizeResult
    r = OptimizeResult(x=np.zeros(3), nfev=4)
    assert r.nfev == 4 and r["x"].shape == (3,)
    with pytest.raises(AttributeError):
        r.missing


def test_export_on_real_reference_handler():
    """Build a real TemplateBundleHandler around the synthetic ring and export it."""
    import os, sys
    from pathlib import Path
    ref = Path(__file__).resolve().parent.parent / "baseline" / "_ref"
    if not (ref / "pyCamSet").exists() or os.environ.get("PCS_SKIP_REFERENCE"):
        pytest.skip("reference copy not present (GPU box / fresh checkout)")
    sys.path[:0] = [str(ref), str(ref / "stubs")]
    try:
        sys.path.insert(0, str(Path(__file__).resolve().parent / "golden"))
        import make_golden as mg
        from pycamset_b200 import handler as H
        handlers = mg.synthetic_handlers(seed=3, n_cams=4, n_poses=6, detect_prob=0.8)
    except Exception as e:  # numba / cv2 missing in some other environment
        pytest.skip(f"reference not importable here: {e}")
    rig, th, x_t, sh, x_s = handlers
    e = H.export_problem(th)
    assert e["blocks"] == ("projection", "extrinsic3D", "template_points")
    assert e["stock_mapping"]
    assert e["dd"].shape[1] == 5 and e["unfixed"].shape[0] == 15 * e["n_cams"] + 6 * e["n_poses"]
    assert np.array_equal(e["dd"], rig.dd())
    es = H.export_problem(sh)
    assert es["blocks"] == ("projection", "extrinsic3D", "rigidTform3d", "free_point") and es["stock_mapping"]
    assert es["unfixed"].shape[0] == 15 * es["n_cams"] + 6 * es["n_poses"] + 3 * es["n_keys"]
    assert int(es["unfixed"].sum()) == x_s.shape[0]
