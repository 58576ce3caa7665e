"""A/B of the residual kernel variants on config 4 (device-resident): per-call time with the L2 flushed, and equality of
the outputs.  Usage: [KRES_RIG=C,M,layout,detect_prob] python tools/kres_ab.py [out.npy]"""
import sys, os, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pycamset_b200 import synthetic as syn
from pycamset_b200.problem import BundleProblem

spec = os.environ.get("KRES_RIG", "32,2000,ring,1.0").split(",")
C, M, layout, dp = int(spec[0]), int(spec[1]), spec[2], float(spec[3])
rig = syn.make_rig(C, M, distortion=True, seed=0, device="cuda:0", layout=layout, detect_prob=dp)
rng = np.random.default_rng(1)
intr, extr, poses = rig.perturbed(rng, 1e-3)
params = rig.param_string(intr, extr, poses)
unfixed = np.ones(params.shape[0], bool); unfixed[15 * C:15 * C + 6] = False
stream = torch.cuda.Stream()
prob = BundleProblem(0, rig.cam, rig.pose, rig.key, rig.uv, C, M, 81, template=rig.template, unfixed=unfixed, stream=stream.cuda_stream)
prob.set_param_string(params)
N = prob.n_obs
r = torch.empty(2 * N, dtype=torch.float64, device="cuda:0")
x = torch.from_numpy(params[unfixed]).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
with torch.cuda.stream(stream):
    fn = lambda: prob.residual_device(r.data_ptr(), x.data_ptr())
    for _ in range(5): fn()
    ts = []
    for _ in range(40):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(); e1.record(stream); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
ms = float(np.median(ts))
print(json.dumps({"rig": spec, "n_obs": N, "ms_per_call": ms,
                  "GBps_40B": 40.0 * N / ms / 1e6, "frac_hbm_40B": 40.0 * N / ms / 1e6 / 6544.7, "r_checksum": float(r.double().abs().sum())}))
if len(sys.argv) > 1:
    np.save(sys.argv[1], r.cpu().numpy())
