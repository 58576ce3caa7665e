"""LM solve on config 4 (for ncu launch lists): 6 iterations."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pycamset_b200 import synthetic as syn
from pycamset_b200.problem import BundleProblem
rig = syn.make_rig(32, 2000, distortion=True, seed=0, device="cuda:0")
rng = np.random.default_rng(1)
intr, extr, poses = rig.perturbed(rng, 1e-3)
params = rig.param_string(intr, extr, poses)
unfixed = np.ones(params.shape[0], bool); unfixed[15 * 32:15 * 32 + 6] = False
prob = BundleProblem(0, rig.cam, rig.pose, rig.key, rig.uv, 32, 2000, 81, template=rig.template, unfixed=unfixed)
prob.set_param_string(params)
x0 = params[unfixed]
prob.lm_solve(x0, max_iter=2, ftol=0, xtol=0, gtol=0)
prob.set_param_string(params)
torch.cuda.synchronize()
x, st = prob.lm_solve(x0, max_iter=int(sys.argv[1]) if len(sys.argv) > 1 else 6, ftol=0, xtol=0, gtol=0)
print(st)
