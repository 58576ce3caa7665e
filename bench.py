#!/usr/bin/env python
"""bench.py -- throughput of the bundle-adjustment hot path on B200 (contract: see DESIGN.md "Measurement").

One "step" = one evaluation of residual + analytic Jacobian + J^T J / J^T r (block normal equations) at a
fixed parameter vector over this rank's observation shard, plus (N > 1) the NCCL all-reduce of the camera
blocks.  Metric: Mobs/s (whole job, all ranks).

    python bench.py                              # N = 1, config 4: 32-camera ring x 2000 poses
    torchrun ... bench.py --gpus N               # weak scaling: 32-camera ring x 2000 poses PER GPU, sharded by pose
    python bench.py --workload dome128           # config 5 (strong scaling across --gpus): 128-camera dome x 20000 poses
    python bench.py --impl reference             # CPU arm: the UNMODIFIED reference (baseline/_ref) on the host cores

Besides the headline the JSON line carries `config5` (the 128-camera dome x 20 000 poses of BASELINE.json configs[4],
STRONG scaling over --gpus: value, ms per evaluation, LM iterations / s) and `lm_e2e` (run_bundle_adjustment from a host
handler to the solution x on configs 1-3, device LM vs the reference's own solver in the reference arm).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "Mobs/s residual+Jacobian+JtJ eval"
ALGO_BYTES_PER_OBS = 28.0  # (u, v) 16 B + cam, pose, key 12 B; K_ne writes O(params), not O(N)
# FP64 work of K_ne per observation: ~115 evaluation FMA slots + 1.5 DMMA m8n8k4 (384 FMA slots issued,
# 272 of them are the upper triangle of the 16-column Gram update); 2 flop per FMA.
FP64_ISSUED_FLOP_PER_OBS = 2.0 * (115 + 384 + 24)   # evaluation + 1.5 DMMA m8n8k4 + segment flush (6 DMMA / ~64 obs)
FP64_USEFUL_FLOP_PER_OBS = 2.0 * (115 + 272)        # upper triangle of the 16 x 16 Gram update, 2 rows
FP64_PEAK_TFLOPS = 36.9    # measured on this pool's B200 by tools/fp64_peak.cu (profiles/r1_fp64_peak_b200.json)

WORKLOADS = {
    # name: (layout, n_cams, poses, detect_prob, scaling)
    "ring32": ("ring", 32, 2000, 1.0, "weak"),      # config 4; poses are PER GPU
    "dome128": ("dome", 128, 20000, 0.5, "strong"),  # config 5; poses are the job total
    "ring8": ("ring", 8, 100, 1.0, "weak"),         # config 1 (parity-sized; for quick runs)
}


def peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md "clocks line").

    NVML (the library nvidia-smi itself reads) is polled every few ms from a thread, because the timed region of
    this bench is tens of ms and `nvidia-smi -lms` cannot sample that fast; if NVML is unavailable the sampler
    falls back to an `nvidia-smi --query-gpu ... -lms 100` subprocess."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device, period_s=0.004):
        self.device, self.period = device, period_s
        self.sm, self.power, self.reasons = [], [], set()
        self.sm_max = None
        self.proc = self.thread = self.nvml = None
        self.stop_flag = threading.Event()
        self.source = None

    def _physical_index(self):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[self.device])
            except Exception:
                pass
        return self.device

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index())
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.source = "nvml"
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self._physical_index())], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._read_smi, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll_nvml(self):
        n = self.nvml
        names = {"hw_slowdown": "nvmlClocksEventReasonHwSlowdown", "hw_thermal_slowdown": "nvmlClocksEventReasonHwThermalSlowdown",
                 "sw_thermal_slowdown": "nvmlClocksEventReasonSwThermalSlowdown", "sw_power_cap": "nvmlClocksEventReasonSwPowerCap"}
        bits = {}
        for k, v in names.items():
            b = getattr(n, v, None) or getattr(n, v.replace("ClocksEventReason", "ClocksThrottleReason"), None)
            if b is not None:
                bits[k] = b
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
                r = n.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                for k, b in bits.items():
                    if r & b:
                        self.reasons.add(k)
                self.power.append(n.nvmlDeviceGetPowerUsage(self.handle) / 1000.0)
            except Exception:
                pass
            time.sleep(self.period)

    def _read_smi(self):
        for line in self.proc.stdout:
            r = [c.strip() for c in line.split(",")]
            try:
                self.sm.append(float(r[1])); self.sm_max = float(r[2]); self.power.append(float(r[3]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    self.reasons.add(name)

    def stop(self):
        self.stop_flag.set()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        if self.thread:
            self.thread.join(timeout=2)
        if self.source is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock sampling unavailable"], "samples": 0}
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.sm_max,
                "sm_min_mhz": float(np.min(self.sm)) if self.sm else None,
                "power_w_max": float(np.max(self.power)) if self.power else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": self.source}


def build_shard(args, rank, world, device, workload=None, sample_poses=0):
    """Synthetic observations of this rank's pose shard, generated directly in HBM.
    sample_poses > 0 (CPU arms): only the first `sample_poses` poses of the job's rig are generated."""
    import torch
    from pycamset_b200 import synthetic as syn
    from pycamset_b200.distributed import even_pose_ranges

    layout, n_cams, poses, detect_prob, scaling = WORKLOADS[workload or args.workload]
    if args.poses and workload in (None, args.workload):
        poses = args.poses
    total_poses = poses * world if scaling == "weak" else poses
    p0, p1 = even_pose_ranges(total_poses, world)[rank]
    if sample_poses:
        p0, p1 = 0, min(sample_poses, total_poses)
    rig = syn.make_rig(n_cams, total_poses, layout=layout, distortion=True, seed=args.seed, detect_prob=detect_prob,
                       pose_start=p0, pose_stop=p1, device=device, order="cam")
    rng = np.random.default_rng(args.seed + 1)
    intr, extr, posesp = rig.perturbed(rng, 1e-3)
    m_local = p1 - p0
    params = np.concatenate([intr.ravel(), extr.ravel(), posesp[p0:p1].ravel()])
    unfixed = np.ones(params.shape[0], bool)
    if p0 == 0:
        unfixed[15 * n_cams:15 * n_cams + 6] = False  # pose 0 is the gauge (template_handler.py:134-137)
    pose_local = (rig.pose - p0).to(torch.int32)
    return dict(rig=rig, cam=rig.cam, pose=pose_local, key=rig.key, uv=rig.uv, n_cams=n_cams, n_poses=m_local,
                params=params, unfixed=unfixed, total_poses=total_poses, scaling=scaling, layout=layout,
                detect_prob=detect_prob)


def oracle_pass(sh, max_obs, threads=None):
    """The CPU restatement (oracle port of the reference path) on a bounded sample of the workload.  Returns
    (one_pass, n_obs_sample, cores): one_pass() evaluates residual + CSR Jacobian values + block J^T J / J^T r once.
    The static structure (CSR pattern, segment ids) is built once outside, like the reference does."""
    from oracle import oracle as orc
    if threads:
        orc.set_threads(threads)
    cores = orc.num_threads()
    n = min(int(sh["cam"].shape[0]), max_obs)
    cam = sh["cam"][:n].cpu().numpy(); pose = sh["pose"][:n].cpu().numpy(); key = sh["key"][:n].cpu().numpy()
    uv = sh["uv"][:n].cpu().numpy()
    C, M = sh["n_cams"], sh["n_poses"]
    o = orc.Problem(0, cam, pose, key, uv, C, M, 81, sh["rig"].template)
    fm = orc.free_map_from_mask(sh["unfixed"])
    col, rp = o.csr_structure(fm)
    pair = cam.astype(np.int64) * M + pose
    _, seg = np.unique(pair, return_inverse=True)
    seg = seg.astype(np.int32); n_seg = int(seg.max()) + 1 if n else 0
    params = sh["params"]

    def one_pass():
        o.residual(params)
        o.csr_values(params, fm, rp)
        o.normal_blocks(params, seg, n_seg)

    return one_pass, n, cores


def oracle_eval_mobs(sh, max_obs, min_seconds, threads=None):
    """Best-of-passes throughput of the oracle port over ~min_seconds: (Mobs/s, n_obs_sample, cores, passes)."""
    one_pass, n, cores = oracle_pass(sh, max_obs, threads)
    one_pass()  # warm-up
    best, passes, t_all = float("inf"), 0, time.perf_counter()
    while passes < 3 or (time.perf_counter() - t_all) < min_seconds:
        t0 = time.perf_counter(); one_pass(); dt = time.perf_counter() - t0
        best = min(best, dt); passes += 1
        if passes >= 200:
            break
    return n / best / 1e6, n, cores, passes


def workload_config(args, world, workload=None):
    """The `config` object: identical in both arms (it names the job, not the arm)."""
    workload = workload or args.workload
    layout, n_cams, poses, detect_prob, scaling = WORKLOADS[workload]
    if args.poses and workload == args.workload:
        poses = args.poses
    total = poses * world if scaling == "weak" else poses
    return {
        "workload": f"{n_cams}-camera {layout} x {total} poses, ChArUco(10,10,4) 81 pts, radial+tangential "
                    f"distortion, template chain (P=21)",
        "n_cams": n_cams, "n_poses_total": total, "poses_per_gpu": -(-total // world),
        "detect_prob": detect_prob, "sharding": f"by pose, {world} rank(s)", "seed": args.seed,
        "l2": "flushed between timed steps (256 MiB write)" + ("; ranks re-aligned after the flush by an untimed NCCL all-reduce" if world > 1 else ""),
    }


# ---------------------------------------------------------------------------------------------------------------
# run_bundle_adjustment end to end (host handler in -> x out) on BASELINE.json configs 1-3
# ---------------------------------------------------------------------------------------------------------------
LM_E2E_CASES = ("config1_ring8x100", "config2_ccube_template", "config3_ccube_selfcal")


def _px(r):
    return float(np.mean(np.linalg.norm(np.reshape(r, (-1, 2)), axis=1)))


def _config1_rig(seed):
    from pycamset_b200 import synthetic as syn
    rig = syn.make_rig(8, 100, layout="ring", distortion=True, seed=seed, detect_prob=1.0)
    intr, extr, poses = rig.perturbed(np.random.default_rng(seed + 1), 1e-3)
    x0 = np.concatenate([intr.ravel(), extr.ravel(), poses[1:].ravel()])      # pose 0 fixed (template_handler.py:134-137)
    return rig, x0


def lm_e2e_handlers(seed, real_reference, max_nfev):
    """Yields (name, handler) one at a time: reference handler objects when the reference is importable, else the
    duck-typed stand-ins of tests/fake_reference.py filled from the same data.  Lazily, and with max_nfev set right before
    the hand-over: the reference's handlers share ONE options dict (DEFAULT_OPTIONS is aliased and mutated,
    template_handler.py:108-110), so building a second handler changes the first one's max_nfev."""
    from tests.helpers import load_case
    for name in LM_E2E_CASES:
        if name == "config1_ring8x100":
            rig, x0 = _config1_rig(seed)
            if real_reference:
                from baseline import reference_arm as ra
                h = ra.ring_handlers(rig, "ring", x_template=x0)
            else:
                from tests import fake_reference as fr
                g = dict(dd=rig.dd(), template=rig.template, chain=0, n_cams=8, n_poses=100, x=x0,
                         param0=np.concatenate([x0[:120], np.zeros(6), x0[120:]]),
                         unfixed=np.concatenate([np.ones(120, bool), np.zeros(6, bool), np.ones(594, bool)]))
                h = fr.TemplateBundleHandler(g)
        else:
            g = load_case("ccube_template" if name == "config2_ccube_template" else "ccube_selfcal")
            if real_reference:
                from baseline import reference_arm as ra
                h = ra.golden_handler(g, max_nfev=max_nfev[name])
            else:
                from tests import fake_reference as fr
                h = (fr.SelfBundleHandler if g["chain"] == 1 else fr.TemplateBundleHandler)(g)
        h.problem_opts["max_nfev"] = max_nfev[name]
        yield name, h


def lm_e2e_device(seed, device):
    """Device arm: pycamset_b200.handler.run_bundle_adjustment(handler) -- problem export + upload, device LM, result
    read-back, residual / Jacobian at the solution -- timed by wall clock around the whole call (best of three calls after
    a warm-up call that pays CUDA module loading and workspace allocation)."""
    from pycamset_b200.handler import run_bundle_adjustment
    out, real = {}, True
    try:
        from baseline import reference_arm as ra
        ra.import_reference()
    except Exception as e:
        real = False
        out["handlers"] = f"stand-ins (reference not importable: {str(e)[:120]})"
    else:
        out["handlers"] = "reference handler objects (baseline/_ref)"
    budget = {k: 100 for k in LM_E2E_CASES}           # the reference's max_nfev (template_handler.py:24-31)
    for name, h in lm_e2e_handlers(seed, real, budget):
        try:
            run_bundle_adjustment(h, device=device)    # warm-up
            dt = float("inf")
            for _ in range(3):                         # best of three: the call is mostly host-side Python, the box is shared
                t0 = time.perf_counter()
                res, _ = run_bundle_adjustment(h, device=device)
                dt = min(dt, time.perf_counter() - t0)
            st = res["lm"]
            out[name] = {"seconds": dt, "iterations": st["iterations"], "iter_per_s": st["iterations"] / dt,
                         "solve_seconds_device": st["seconds"], "status": st["status"], "cost_final": float(res.cost),
                         "final_px": _px(res.fun), "n_free": int(len(res.x)), "n_obs": int(len(res.fun) // 2)}
        except Exception as e:  # report, never hide
            out[name] = {"error": str(e)[:200]}
    return out


def lm_e2e_reference(seed, threads):
    """Reference arm: the reference's own run_bundle_adjustment (scipy TRF + LSMR on its numba closures).  Configs 2 / 3
    run the reference's full budget (max_nfev = 100, as its tests do); config 1 is cut to 10 evaluations to keep the arm
    within minutes -- the rate (evaluations / s) is what is compared."""
    from baseline import reference_arm as ra
    budget = {"config1_ring8x100": 10, "config2_ccube_template": 100, "config3_ccube_selfcal": 100}
    out = {"handlers": "reference handler objects (baseline/_ref)"}
    for name, h in lm_e2e_handlers(seed, True, budget):
        try:
            res, dt = ra.run_ba(h, threads)
            out[name] = {"seconds": dt, "iterations": int(res.nfev), "iter_per_s": res.nfev / dt, "status": int(res.status),
                         "cost_final": float(res.cost), "final_px": _px(res.fun), "n_free": int(len(res.x)),
                         "n_obs": int(len(res.fun) // 2), "max_nfev": budget[name]}
        except Exception as e:
            out[name] = {"error": str(e)[:200]}
    return out


# ---------------------------------------------------------------------------------------------------------------
# --impl reference
# ---------------------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """--impl reference: the UNMODIFIED reference (baseline/_ref/pyCamSet) on all host cores.  One step = one pass
    loss_fun(x) + jac_fn(x) + J.T @ J + J.T @ r (the reference's closures + scipy.sparse) over a bounded sample of the
    arm's workload: the first --ref-poses poses of the same rig.  W warm-up passes (the first one JIT-compiles), then
    exactly K timed passes.  The oracle port (C + OpenMP) is timed beside it on the same sample as `port_value`.  If the
    reference cannot be imported the port is the arm (kind "port") and the exception text is reported.  Rank 0 only."""
    if rank != 0:
        return
    from baseline import reference_arm as ra
    cores = ra.pin_thread_env()      # before numba / OpenMP start: torchrun exports OMP_NUM_THREADS=1 to its children
    t_all = time.perf_counter()
    sh = build_shard(args, 0, world, "cpu", sample_poses=args.ref_poses)
    rig = sh["rig"]
    n = int(sh["cam"].shape[0])
    kind, err, build_s = "reference", None, None
    port_pass, _, port_cores = oracle_pass(sh, n, cores)
    try:
        import dataclasses
        m = sh["n_poses"]
        trimmed = dataclasses.replace(rig, poses=rig.poses[:m])
        x0 = sh["params"][sh["unfixed"]]
        t0 = time.perf_counter()
        h = ra.ring_handlers(trimmed, sh["layout"], x_template=x0)
        from pyCamSet.optimisation.optimisation_handling import make_optimisation_function
        loss, jac, x_init = make_optimisation_function(h, cores)
        assert x_init.shape == x0.shape
        ra.one_pass(loss, jac, x0)                      # JIT compile + first call
        build_s = time.perf_counter() - t0
        one = lambda: ra.one_pass(loss, jac, x0)
    except Exception as e:
        kind, err = "port", f"{type(e).__name__}: {str(e)[:300]}"
        one = port_pass
    for _ in range(max(args.warmup, 1)):
        one()
    t1 = time.perf_counter()
    for _ in range(args.steps):
        one()
    dt = time.perf_counter() - t1
    v = n * args.steps / dt / 1e6
    port_pass()
    tp = time.perf_counter(); port_pass(); port_v = n / (time.perf_counter() - tp) / 1e6
    sample = (f"first {sh['n_poses']} poses ({n} observations) of the arm's rig, camera-major; one pass = loss_fun + jac_fn "
              f"(numba, threads = {cores}) + J.T@J + J.T@r (scipy.sparse)" if kind == "reference" else
              f"first {sh['n_poses']} poses ({n} observations) of the arm's rig; oracle port: residual + CSR Jacobian + block JtJ/Jtr")
    lm = None
    if not args.no_lm and kind == "reference":
        try:
            lm = lm_e2e_reference(args.seed, cores)
        except Exception as e:
            lm = {"error": str(e)[:200]}
    out = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": "Mobs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": WORKLOADS[args.workload][4], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, world),
        "cpu_baseline": {"value": v, "unit": "Mobs/s", "cores": cores, "kind": kind, "sample": sample,
                         "port_value": port_v, "port_cores": port_cores, "reference_error": err,
                         "reference_build_s": build_s},
        "e2e": {"value": v, "unit": "Mobs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "lm_e2e": lm, "gpu_launches": 0, "wall_s": time.perf_counter() - t_all,
    }
    print(json.dumps(out), flush=True)


def cpu_baseline_for_our_arm(args, sh):
    """`cpu_baseline` of the GPU arm's line (rank 0, N = 1): the reference itself, run as a child process (`bench.py
    --impl reference` on a small sample: its numba / OpenMP thread pools stay out of this process), with the oracle port
    on ALL observations of the workload beside it."""
    mobs, n_s, cores, passes = oracle_eval_mobs(sh, int(sh["cam"].shape[0]), 3.0)
    port = {"value": mobs, "cores": cores, "sample": f"all {n_s} observations; residual + CSR Jacobian + block JtJ/Jtr, best of {passes} passes"}
    try:
        cmd = [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "1", "--no-lm",
               "--workload", args.workload, "--seed", str(args.seed), "--ref-poses", str(args.cpu_ref_poses)]
        env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
        pr = subprocess.run(cmd, capture_output=True, text=True, timeout=args.cpu_timeout, env=env)
        line = [l for l in pr.stdout.splitlines() if l.startswith("{")][-1]
        ref = json.loads(line)["cpu_baseline"]
        if ref["kind"] != "reference":
            raise RuntimeError(ref.get("reference_error") or "reference arm fell back to the port")
        return {"value": ref["value"], "unit": "Mobs/s", "cores": ref["cores"], "kind": "reference", "sample": ref["sample"],
                "port_value": port["value"], "port_cores": port["cores"], "port_sample": port["sample"]}
    except Exception as e:
        return {"value": port["value"], "unit": "Mobs/s", "cores": port["cores"], "kind": "port", "sample": port["sample"],
                "reference_error": f"{type(e).__name__}: {str(e)[:300]}"}


# ---------------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------------
class Exchange:
    """N > 1: the one exchange step of a pose-sharded evaluation is the sum of the camera blocks [U | gc | cost].
    Default: the library's one-shot all-reduce over NVLink peer memory (csrc/pcs_p2p.cu), checked against an NCCL
    all-reduce once before timing; --exchange nccl (or a failed set-up) uses torch.distributed."""

    def __init__(self, args, prob, x_dev, stream, dev, world):
        import torch
        import torch.distributed as dist
        from pycamset_b200 import distributed as pdist
        self.prob, self.world, self.kind, self.check, self.p2p = prob, world, None, None, None
        self.pdist = pdist
        if world == 1:
            return
        self.kind = "nccl"
        if args.exchange != "p2p":
            return
        try:
            p2p = pdist.P2PCameraAllReduce(prob)
            n_head = prob.n_cams * 240 + 1
            head = pdist.tensor_from_ptr(prob.device_buffers().U, n_head, dev)
            with torch.cuda.stream(stream):
                prob.normal_equations_device(x_dev.data_ptr())
                ref = head.clone()
                dist.all_reduce(ref)
                p2p()
                err = float(((head - ref).abs().max() / ref.abs().max()).item())
                p2p()   # second call exercises the other data slot (result: world * sum; only the protocol matters)
            torch.cuda.synchronize(dev)
            ok = torch.tensor([1.0 if err < 1e-12 else 0.0], device=f"cuda:{dev}")
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if ok.item() < 1.0:
                raise RuntimeError(f"peer-memory all-reduce disagrees with NCCL (rel err {err:.3e})")
            self.p2p, self.kind, self.check = p2p, "p2p", f"matches NCCL all-reduce, rel err {err:.1e}"
        except Exception as e:  # symmetric memory unavailable on this box: say so and use NCCL
            self.kind, self.check = "nccl", f"p2p unavailable: {str(e)[:160]}"

    def __call__(self):
        if self.world == 1:
            return
        if self.kind == "p2p":
            self.p2p()
        else:
            self.pdist.allreduce_camera_blocks(self.prob)


def measure_workload(args, workload, rank, world, dev, stream, flush_buf, steps, warmup, with_extras):
    """Build this rank's shard of `workload`, time `steps` evaluations (device-resident inputs), and optionally the
    end-to-end / callback / LM figures.  Returns a dict (identical keys on every rank; aggregates are whole-job)."""
    import torch
    import torch.distributed as dist
    from pycamset_b200.problem import BundleProblem
    from pycamset_b200 import distributed as pdist

    t_setup = time.perf_counter()
    sh = build_shard(args, rank, world, f"cuda:{dev}", workload=workload)
    torch.cuda.synchronize(dev)
    prob = BundleProblem(0, sh["cam"], sh["pose"], sh["key"], sh["uv"], sh["n_cams"], sh["n_poses"], 81,
                         template=sh["rig"].template, unfixed=sh["unfixed"], device=dev, stream=stream.cuda_stream)
    prob.set_param_string(sh["params"])
    x_host = torch.from_numpy(sh["params"][sh["unfixed"]].copy()).pin_memory()
    x_dev = x_host.to(f"cuda:{dev}")
    n_local = prob.n_obs
    setup_s = time.perf_counter() - t_setup
    exch = Exchange(args, prob, x_dev, stream, dev, world)

    def step():
        prob.normal_equations_device(x_dev.data_ptr())
        exch()

    # N > 1: the untimed L2 flush before every step takes a slightly different time on every rank; without re-alignment
    # that jitter would be charged to the step (the exchange makes the early ranks wait for the late ones).  A tiny NCCL
    # all-reduce enqueued on the stream after the flush and BEFORE the start event lines the ranks up again; it is not
    # inside any timed interval.
    align_buf = torch.zeros(1, device=f"cuda:{dev}") if world > 1 else None

    def flush_and_align():
        flush_buf.zero_()
        if world > 1:
            dist.all_reduce(align_buf)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    res = {}
    with torch.cuda.stream(stream):
        for _ in range(warmup):
            flush_buf.zero_()
            step()
        barrier()
        sampler = ClockSampler(dev)
        if rank == 0:
            sampler.start()
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        barrier()
        launches0 = prob.launch_count()
        t_wall = time.perf_counter()
        for k in range(steps):
            flush_and_align()                 # L2 flush (+ rank alignment), outside the per-step event pair
            starts[k].record(stream)
            step()
            ends[k].record(stream)
        barrier()
        wall_s = time.perf_counter() - t_wall
        gpu_launches = prob.launch_count() - launches0
        clocks = sampler.stop() if rank == 0 else None
        # K_ne's own duration (roofline): a second loop with the library's per-launch event pairs.  They are kept out of the
        # timed loop above because an event record between two launches switches off the programmatic dependent launch
        # that lets the kernel become resident behind the table set-up.
        prob.timing_enable(True)
        for k in range(steps):
            flush_buf.zero_()
            step()
        barrier()
        kern_ms = prob.timing_all_ms()[-steps:]
        prob.timing_enable(False)
    step_ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
    total_ms = float(np.sum(step_ms))
    n_total = n_local
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=f"cuda:{dev}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        c = torch.tensor([n_local], dtype=torch.int64, device=f"cuda:{dev}")
        dist.all_reduce(c)
        n_total = int(c.item())
    ms_per_step = total_ms / steps
    if exch.kind == "p2p" and prob.p2p_timed_out():   # a poll gave up waiting for a peer: the sums (and the number) are invalid
        exch.check = (exch.check or "") + "; TIMED OUT waiting for a peer"
    res.update(value=n_total / (ms_per_step * 1e-3) / 1e6, ms_per_step=ms_per_step, n_obs_total=n_total, n_obs_per_gpu=n_local,
               n_segments_per_gpu=prob.n_segments, n_free_per_gpu=prob.n_free, kernel_ms=float(np.mean(kern_ms)),
               gpu_launches=gpu_launches, clocks=clocks, setup_s=setup_s, wall_s_timed_region=wall_s,
               exchange=exch.kind, exchange_check=exch.check, scaling=sh["scaling"], sh=sh)
    xh = x_host.numpy()

    def timed_steps(n_steps):
        """(ms per step max over ranks, mean kernel ms) of n_steps evaluations with the current settings"""
        with torch.cuda.stream(stream):
            for _ in range(3):
                flush_buf.zero_(); step()
            barrier()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(n_steps)]
            for e0, e1 in ev:
                flush_and_align(); e0.record(stream); step(); e1.record(stream)
            barrier()
            prob.timing_enable(True)
            for _ in range(n_steps):
                flush_buf.zero_(); step()
            barrier()
            km = prob.timing_all_ms()[-n_steps:]
            prob.timing_enable(False)
        tot = torch.tensor([float(sum(a.elapsed_time(b) for a, b in ev))], dtype=torch.float64, device=f"cuda:{dev}")
        if world > 1:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        return float(tot.item()) / n_steps, float(np.mean(km))

    def lm_rate():
        prob.set_param_string(sh["params"])
        with torch.cuda.stream(stream):
            prob.lm_solve(xh, max_iter=2, ftol=0, xtol=0, gtol=0)        # warm-up (workspace allocation)
            prob.set_param_string(sh["params"])
            barrier()
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            _, st = prob.lm_solve(xh, max_iter=args.lm_iters, ftol=0, xtol=0, gtol=0)   # host x0 in, host x out (synchronises)
            wall = time.perf_counter() - t0
            barrier()
        secs = torch.tensor([st["seconds"], wall], dtype=torch.float64, device=f"cuda:{dev}")
        if world > 1:
            dist.all_reduce(secs, op=dist.ReduceOp.MAX)
        # end to end through the solver call a user makes: every LM iteration evaluates residual + Jacobian + JtJ over all
        # observations of all ranks; wall clock around the host-facing call, max over ranks
        return {"iter_per_s": st["iterations"] / float(secs[0].item()), "iterations": st["iterations"],
                "cost_initial": st["cost_initial"], "cost_final": st["cost_final"], "status": st["status"],
                "e2e_wall": {"value": n_total * st["iterations"] / float(secs[1].item()) / 1e6, "unit": "Mobs/s",
                             "seconds": float(secs[1].item()), "h2d_bytes_per_solve": int(xh.nbytes), "d2h_bytes_per_solve": int(xh.nbytes),
                             "call": "BundleProblem.lm_solve(x0_host) -> x_host (one evaluation per iteration, reduced system "
                                     "all-reduced over NCCL for N > 1)"}}

    if with_extras:
        # ---- end to end through the host-facing C-ABI call: host x in, all blocks out to (pinned) host memory ------
        C, M, S = prob.n_cams, prob.n_poses, prob.n_segments
        def pinned(*shape):
            return torch.empty(shape, dtype=torch.float64).pin_memory().numpy()
        outb = dict(U=pinned(C, 15, 15), gc=pinned(C, 15), V=pinned(M, 6, 6), gp=pinned(M, 6), W=pinned(S, 15, 6),
                    cost_buf=pinned(1))
        h2d = int(xh.nbytes)
        d2h = int(sum(v.nbytes for v in outb.values()))
        with torch.cuda.stream(stream):
            for _ in range(3):
                prob.normal_equations(xh, out=outb)
            e2e_steps = max(3, min(steps, 10))
            barrier()
            e2e_s = 0.0
            for _ in range(e2e_steps):
                flush_buf.zero_()
                torch.cuda.synchronize(dev)
                t0 = time.perf_counter()
                prob.normal_equations(xh, out=outb)   # H2D x, kernels, D2H blocks, stream sync
                e2e_s += time.perf_counter() - t0
            barrier()
        e2e_t = torch.tensor([e2e_s], dtype=torch.float64, device=f"cuda:{dev}")
        if world > 1:
            dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        res["e2e"] = {"value": n_total / (float(e2e_t.item()) / e2e_steps) / 1e6, "unit": "Mobs/s", "h2d_bytes_per_step": h2d,
                      "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                      "call": "BundleProblem.normal_equations(x_host) -> U, gc, V, gp, W, cost on host"}
        del outb

        # ---- the reference's own callbacks (loss_fun / jac_fn drop-ins): HBM-bound kernels ------------------------------
        callbacks = None
        if world == 1 and not args.no_callbacks:
            try:
                peak_hbm, _ = peaks()
                nnz = prob.nnz
                r_dev = torch.empty(2 * n_local, dtype=torch.float64, device=f"cuda:{dev}")
                v_dev = torch.empty(max(nnz, 1), dtype=torch.float64, device=f"cuda:{dev}")
                callbacks = {}
                with torch.cuda.stream(stream):
                    for name, fn, nbytes in (
                            ("K_res", lambda: prob.residual_device(r_dev.data_ptr(), x_dev.data_ptr()), 44.0 * n_local),
                            ("K_jac", lambda: prob.jacobian_values_device(v_dev.data_ptr(), x_dev.data_ptr()), 28.0 * n_local + 8.0 * nnz)):
                        for _ in range(3):
                            fn()
                        ts = []
                        for _ in range(20):
                            flush_buf.zero_()
                            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                            e0.record(stream); fn(); e1.record(stream)
                            torch.cuda.synchronize(dev)
                            ts.append(e0.elapsed_time(e1))
                        ms = float(np.median(ts))
                        callbacks[name] = {"ms_per_call": ms, "Mobs_per_s": n_local / ms / 1e3, "algorithmic_GBps": nbytes / ms / 1e6,
                                           "frac_of_hbm_peak": nbytes / ms / 1e6 / peak_hbm,
                                           "includes": "x scatter + table set-up + kernel (whole call)"}
                del r_dev, v_dev
            except Exception as e:  # report, never hide
                callbacks = {"error": str(e)[:200]}
        res["callbacks"] = callbacks

    # ---- LM iterations / s (device-resident solve; all-reduce of the reduced camera system for N > 1) --------------
    lm = None
    if not args.no_lm:
        try:
            if world > 1:
                pdist.install_nccl_allreduce(prob)
            lm = lm_rate()
        except Exception as e:  # report, never hide
            lm = {"error": str(e)[:200]}

    # ---- the same evaluation with the mixed-precision kernel (FP64 residual / cost / gradients, BF16-split J^T J) ------
    try:
        prob.set_normal_precision(True)
        ms_m, kern_m = timed_steps(steps)
        mixed = {"value": n_total / (ms_m * 1e-3) / 1e6, "unit": "Mobs/s", "ms_per_step": ms_m, "kernel_ms": kern_m,
                 "contract": "cost, g_c, g_m FP64 (1e-9 vs oracle); U, V, W within 1e-4 sqrt(d_a d_b); tests/test_gpu_mixed_precision.py"}
        if not args.no_lm:
            mixed["lm"] = lm_rate()
        prob.set_normal_precision(False)
    except Exception as e:
        mixed = {"error": str(e)[:200]}
    res["mixed"] = mixed
    res["lm"] = lm
    prob.close()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="ring32", choices=sorted(WORKLOADS))
    ap.add_argument("--poses", type=int, default=0, help="override the workload's pose count")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--ref-poses", type=int, default=100, help="--impl reference: poses of the rig in the CPU sample")
    ap.add_argument("--cpu-ref-poses", type=int, default=40, help="GPU arm: poses in the cpu_baseline sample of the reference")
    ap.add_argument("--cpu-timeout", type=int, default=240)
    ap.add_argument("--no-lm", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-callbacks", action="store_true", help="skip the K_res / K_jac (loss_fun / jac_fn drop-in) rates")
    ap.add_argument("--no-config5", action="store_true", help="skip the dome128 (BASELINE config 5) strong-scaling measurement")
    ap.add_argument("--no-lm-e2e", action="store_true", help="skip run_bundle_adjustment end to end on configs 1-3")
    ap.add_argument("--lm-iters", type=int, default=10)
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"], help="N > 1: camera-block all-reduce path")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = local_rank
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{dev}"))
    stream = torch.cuda.Stream(device=dev)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=f"cuda:{dev}")

    main_res = measure_workload(args, args.workload, rank, world, dev, stream, flush_buf, args.steps, args.warmup, True)
    sh = main_res.pop("sh")

    # BASELINE.json configs[4]: 128-camera dome x 20 000 poses (~10^8 observations), STRONG scaling over the ranks
    config5 = None
    if not args.no_config5 and args.workload != "dome128":
        try:
            r5 = measure_workload(args, "dome128", rank, world, dev, stream, flush_buf, max(3, min(args.steps, 10)), 3, False)
            r5.pop("sh")
            config5 = {"value": r5["value"], "unit": "Mobs/s", "ms_per_step": r5["ms_per_step"], "scaling": "strong",
                       "n_obs_total": r5["n_obs_total"], "n_obs_per_gpu": r5["n_obs_per_gpu"], "kernel_ms": r5["kernel_ms"],
                       "lm": r5["lm"], "mixed": r5.get("mixed"), "exchange": r5["exchange"], "setup_s": r5["setup_s"],
                       "config": workload_config(args, world, "dome128")}
        except Exception as e:  # report, never hide
            config5 = {"error": str(e)[:200]}

    if rank == 0:
        peak, peak_src = peaks()
        n_local, kern_avg_ms = main_res["n_obs_per_gpu"], main_res["kernel_ms"]
        achieved = ALGO_BYTES_PER_OBS * n_local / (kern_avg_ms * 1e-3) / 1e9
        traffic = None
        ncu_pipes = None   # measured pipe utilisation of the same kernel under ncu (next to the op-count estimate below)
        tf = ROOT / "profiles" / "traffic.json"
        if tf.exists():
            try:
                tj = json.loads(tf.read_text())
                traffic = tj.get(args.workload)
                ncu_pipes = tj.get(args.workload + "_ncu_pipes")
            except Exception:
                traffic = None
        cpu = None
        if world == 1 and not args.no_cpu:
            cpu = cpu_baseline_for_our_arm(args, sh)
        lm_e2e = None
        if world == 1 and not args.no_lm_e2e and not args.no_lm:
            try:
                lm_e2e = lm_e2e_device(args.seed, dev)
            except Exception as e:
                lm_e2e = {"error": str(e)[:200]}
        mixed_out = main_res.get("mixed")
        if mixed_out and "kernel_ms" in mixed_out:
            a_m = ALGO_BYTES_PER_OBS * n_local / (mixed_out["kernel_ms"] * 1e-3) / 1e9
            mixed_out["roofline"] = {"bound": "hbm", "achieved": a_m, "peak": peak, "unit": "GB/s", "frac": a_m / peak,
                                     "kernel": "k_normal_mixed", "kernel_ms": mixed_out["kernel_ms"], "traffic": None}
        out = {
            "metric": METRIC, "value": main_res["value"], "unit": "Mobs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main_res["ms_per_step"], "higher_is_better": True, "scaling": main_res["scaling"],
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, world),
            "config_detail": {k: main_res[k] for k in ("exchange", "exchange_check", "n_obs_total", "n_obs_per_gpu",
                                                        "n_segments_per_gpu", "n_free_per_gpu")},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "normal-equation kernel (K_ne)",
                         "kernel_ms": kern_avg_ms, "algorithmic_bytes_per_obs": ALGO_BYTES_PER_OBS,
                         "note": "K_ne is bound by the dispatch cycles of its FP64 / tensor-FP64 instructions, not by HBM (DESIGN.md 4, "
                                 "profiles/r2_kne_knockout.txt); see the fp64 object",
                         "fp64": {"issued_tflops": FP64_ISSUED_FLOP_PER_OBS * n_local / (kern_avg_ms * 1e-3) / 1e12,
                                  "useful_tflops": FP64_USEFUL_FLOP_PER_OBS * n_local / (kern_avg_ms * 1e-3) / 1e12,
                                  "peak_tflops": FP64_PEAK_TFLOPS, "peak_source": "measured (tools/fp64_peak.cu)", "ncu": ncu_pipes,
                                  "frac_issued": FP64_ISSUED_FLOP_PER_OBS * n_local / (kern_avg_ms * 1e-3) / 1e12 / FP64_PEAK_TFLOPS}},
            "cpu_baseline": cpu,
            "e2e": main_res["e2e"],
            "gpu_launches": main_res["gpu_launches"], "clocks": main_res["clocks"], "lm": main_res["lm"],
            "callbacks": main_res["callbacks"], "mixed": mixed_out, "config5": config5, "lm_e2e": lm_e2e,
            "setup_s": main_res["setup_s"], "wall_s_timed_region": main_res["wall_s_timed_region"],
        }
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
