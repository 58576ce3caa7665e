"""LM solve for ncu launch lists: python tools/lm_profile.py [iters] [workload] [mixed]

workload: ring32 (config 4, default) | dome128:<poses> (config-5 shape, C = 128, n = 1920) | ccube_selfcal (config 3, the
self-calibration chain on the block path) | ccube_template (config 2)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pycamset_b200 import synthetic as syn
from pycamset_b200.problem import BundleProblem

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 6
workload = sys.argv[2] if len(sys.argv) > 2 else "ring32"
mixed = len(sys.argv) > 3 and sys.argv[3] == "mixed"

if workload.startswith("ccube"):
    from tests.helpers import load_case
    g = load_case(workload)
    dd = g["dd"]
    prob = BundleProblem(int(g["chain"]), dd[:, 0], dd[:, 1], dd[:, 2], dd[:, 3:5], int(g["n_cams"]), int(g["n_poses"]),
                         g["template"].shape[0], template=g["template"] if int(g["chain"]) == 0 else None, unfixed=g["unfixed"])
    params, x0 = g["param0"], g["x"]
else:
    if workload.startswith("dome128"):
        C, M, layout, prob_d = 128, int(workload.split(":")[1]) if ":" in workload else 2500, "dome", 0.5
    else:
        C, M, layout, prob_d = 32, 2000, "ring", 1.0
    rig = syn.make_rig(C, M, layout=layout, distortion=True, seed=0, detect_prob=prob_d, device="cuda:0")
    intr, extr, poses = rig.perturbed(np.random.default_rng(1), 1e-3)
    params = rig.param_string(intr, extr, poses)
    unfixed = np.ones(params.shape[0], bool); unfixed[15 * C:15 * C + 6] = False
    prob = BundleProblem(0, rig.cam, rig.pose, rig.key, rig.uv, C, M, 81, template=rig.template, unfixed=unfixed)
    x0 = params[unfixed]
if mixed:
    prob.set_normal_precision(True)
prob.set_param_string(params)
prob.lm_solve(x0, max_iter=2, ftol=0, xtol=0, gtol=0)
prob.set_param_string(params)
torch.cuda.synchronize()
x, st = prob.lm_solve(x0, max_iter=iters, ftol=0, xtol=0, gtol=0)
print(workload, "mixed" if mixed else "fp64", st, "iter/s", st["iterations"] / st["seconds"])
