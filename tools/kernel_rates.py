"""Throughput of the residual (K_res), explicit-Jacobian (K_jac) and cost kernels on config 4 (device-resident)."""
import sys, os, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pycamset_b200 import synthetic as syn
from pycamset_b200.problem import BundleProblem

dev = 0
rig = syn.make_rig(32, 2000, distortion=True, seed=0, device="cuda:0")
rng = np.random.default_rng(1)
intr, extr, poses = rig.perturbed(rng, 1e-3)
params = rig.param_string(intr, extr, poses)
unfixed = np.ones(params.shape[0], bool); unfixed[15 * 32:15 * 32 + 6] = False
stream = torch.cuda.Stream()
prob = BundleProblem(0, rig.cam, rig.pose, rig.key, rig.uv, 32, 2000, 81, template=rig.template, unfixed=unfixed, stream=stream.cuda_stream)
prob.set_param_string(params)
N, nnz = prob.n_obs, prob.nnz
r = torch.empty(2 * N, dtype=torch.float64, device="cuda:0")
vals = torch.empty(nnz, dtype=torch.float64, device="cuda:0")
x = torch.from_numpy(params[unfixed]).cuda()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
out = {"n_obs": N, "nnz": nnz}
peak = 6544.7
with torch.cuda.stream(stream):
    for name, fn, nbytes in (("K_res", lambda: prob.residual_device(r.data_ptr(), x.data_ptr()), 44.0 * N),
                             ("K_jac", lambda: prob.jacobian_values_device(vals.data_ptr(), x.data_ptr()), 28.0 * N + 8.0 * nnz)):
        for _ in range(5): fn()
        ts = []
        for _ in range(30):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); fn(); e1.record(stream); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        out[name] = {"ms_per_call_incl_prepare": ms, "Gobs_per_s": N / ms / 1e6, "algorithmic_GBps": nbytes / ms / 1e6,
                     "frac_of_measured_hbm_peak": nbytes / ms / 1e6 / peak}
print(json.dumps(out))
