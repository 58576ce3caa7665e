"""Diagnostic: worst CSR Jacobian entries of the dome128 rig, CUDA vs oracle (which columns, how large, which rvec)."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import oracle as orc
from pycamset_b200 import synthetic as syn
from pycamset_b200.problem import BundleProblem

C, M, K = 128, 200, 81
rig = syn.make_rig(C, M, layout="dome", distortion=True, seed=0, detect_prob=0.5)
intr, extr, poses = rig.perturbed(np.random.default_rng(1), 1e-3)
params = rig.param_string(intr, extr, poses)
unfixed = np.ones(params.shape[0], bool); unfixed[15 * C:15 * C + 6] = False
cam, pose, key, uv = rig.cam.numpy(), rig.pose.numpy(), rig.key.numpy(), rig.uv.numpy()
o = orc.Problem(0, cam, pose, key, uv, C, M, K, rig.template)
fm = orc.free_map_from_mask(unfixed)
with BundleProblem(0, cam, pose, key, uv, C, M, K, template=rig.template, unfixed=unfixed) as p:
    p.set_param_string(params)
    col, rp = p.csr_structure()
    vals = p.jacobian_values()
ref = o.csr_values(params, fm, rp)
rel = np.abs(vals - ref) / np.maximum(np.abs(ref), 1e-12)
worst = np.argsort(rel)[-12:][::-1]
rows = np.searchsorted(rp, worst, side="right") - 1
for w, r in zip(worst, rows):
    i = r // 2
    c, m = cam[i], pose[i]
    rowmax = np.max(np.abs(ref[rp[r]:rp[r + 1]]))
    print(f"rel {rel[w]:.3e} val {vals[w]:.6e} ref {ref[w]:.6e} col {col[w]} (pos in row {w - rp[r]}) rowmax {rowmax:.3e} "
          f"cam {c} |r_c| {np.linalg.norm(extr[c, :3]):.6f} pose {m} |r_m| {np.linalg.norm(poses[m, :3]):.6f}")
