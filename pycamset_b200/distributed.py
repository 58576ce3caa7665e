"""Multi-GPU plumbing: shard observations by target pose, combine camera blocks with NCCL.

One process per GPU.  Pose blocks V_m, pose-camera blocks W_{c,m} and g_m live on exactly one rank; the camera
blocks (and, inside the LM solver, the Schur-reduced camera system) are partial sums that are all-reduced over
NVLink once per evaluation.  x is replicated; no observation ever moves.
"""
from __future__ import annotations

import numpy as np


def even_pose_ranges(n_poses: int, world: int):
    """Contiguous pose ranges of (almost) equal size: [(start, stop)] * world."""
    base, rem = divmod(n_poses, world)
    out, s = [], 0
    for r in range(world):
        e = s + base + (1 if r < rem else 0)
        out.append((s, e))
        s = e
    return out


def balanced_pose_ranges(obs_per_pose, world: int):
    """Contiguous pose ranges balanced by observation count (greedy split of the prefix sum)."""
    obs_per_pose = np.asarray(obs_per_pose, np.int64)
    M = obs_per_pose.shape[0]
    csum = np.concatenate([[0], np.cumsum(obs_per_pose)])
    total = csum[-1]
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        c = int(np.searchsorted(csum, target, side="left"))
        c = min(max(c, cuts[-1]), M)
        cuts.append(c)
    cuts.append(M)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def shard_observations(cam, pose, key, uv, pose_range):
    """Rows whose pose lies in [start, stop), with pose indices made local to the shard."""
    s, e = pose_range
    sel = (pose >= s) & (pose < e)
    return cam[sel], pose[sel] - s, key[sel], uv[sel]


def shard_param_string(params, n_cams, n_poses, pose_range, n_keys=0):
    """Parameter string of a shard: all cameras, the shard's poses (and all points)."""
    s, e = pose_range
    params = np.asarray(params)
    C15 = 15 * n_cams
    parts = [params[:C15], params[C15 + 6 * s:C15 + 6 * e]]
    if n_keys:
        parts.append(params[C15 + 6 * n_poses:])
    return np.concatenate(parts)


class _DevicePtr:
    """Zero-copy view of a raw device pointer through the CUDA array interface."""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False), "version": 2}


def tensor_from_ptr(ptr: int, n: int, device: int):
    import torch
    return torch.as_tensor(_DevicePtr(ptr, n), device=f"cuda:{device}")


def install_nccl_allreduce(problem, group=None):
    """Install a torch.distributed all-reduce as the LM solver's combine hook (op 0 = sum, 1 = max)."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dev = problem.device

    def fn(ptr, n, op, stream):
        t = tensor_from_ptr(ptr, n, dev)
        ext = torch.cuda.ExternalStream(stream, device=dev) if stream else torch.cuda.current_stream(dev)
        with torch.cuda.stream(ext):
            dist.all_reduce(t, op=dist.ReduceOp.MAX if op == 1 else dist.ReduceOp.SUM, group=group)

    problem.set_allreduce(fn, rank, world)
    return rank, world


def allreduce_camera_blocks(problem, group=None):
    """Sum [U | gc | cost] (contiguous head of the normal-equation buffer) across ranks, on the problem's stream."""
    import torch
    import torch.distributed as dist

    b = problem.device_buffers()
    n = problem.n_cams * 240 + 1
    t = tensor_from_ptr(b.U, n, problem.device)
    ext = torch.cuda.ExternalStream(b.stream, device=problem.device)
    with torch.cuda.stream(ext):
        dist.all_reduce(t, group=group)
    return t


class P2PCameraAllReduce:
    """Low-latency replacement for `allreduce_camera_blocks`: the camera blocks are exchanged by one single-CTA kernel
    over NVLink peer memory (csrc/pcs_p2p.cu).  torch's symmetric-memory allocator only provides the plumbing: a
    buffer per rank that every peer has mapped."""

    def __init__(self, problem, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        group = group or dist.group.WORLD
        self.problem = problem
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        nbytes = problem.p2p_buffer_bytes(self.world)
        self.buf = symm_mem.empty(nbytes // 8, dtype=torch.float64, device=f"cuda:{problem.device}")
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, group.group_name)
        torch.cuda.synchronize(problem.device)
        dist.barrier(group)                       # every rank's flags are zero before anyone starts signalling
        problem.p2p_setup(self.rank, self.world, list(self.handle.buffer_ptrs), nbytes)
        torch.cuda.synchronize(problem.device)
        dist.barrier(group)                       # ... and every rank's slots hold the "empty" sentinel (filled by the setup call)

    def __call__(self):
        self.problem.p2p_allreduce_camera_blocks()


def lm_solve_sharded(cam, pose, key, uv, n_cams, n_poses, n_keys, template, params, unfixed, *, device=None, group=None,
                     max_iter=100, ftol=1e-8, xtol=1e-8, gtol=1e-8, lambda0=1e-3, verbose=0):
    """Pose-sharded Levenberg-Marquardt over the ranks of `group` (template chain), one process per GPU.

    Every rank passes the SAME full problem (observation table in dd order, full parameter string, boolean `unfixed`
    mask over it); the rank keeps only the observations of its contiguous pose range (balanced by observation count),
    eliminates its own poses and all-reduces the Schur-reduced camera system over NCCL.  Returns
    (params_out, stats): the FULL parameter string at the solution -- camera blocks are replicated, pose blocks are
    all-gathered from their owners -- identical on every rank."""
    import torch
    import torch.distributed as dist
    from .problem import BundleProblem

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if device is None:
        device = torch.cuda.current_device()
    cam = np.asarray(cam); pose = np.asarray(pose); key = np.asarray(key); uv = np.asarray(uv).reshape(-1, 2)
    params = np.asarray(params, np.float64)
    unfixed = np.asarray(unfixed, bool)
    ranges = balanced_pose_ranges(np.bincount(pose.astype(np.int64), minlength=n_poses), world)
    s, e = ranges[rank]
    c_s, p_s, k_s, uv_s = shard_observations(cam, pose, key, uv, (s, e))
    par = shard_param_string(params, n_cams, n_poses, (s, e))
    unf = shard_param_string(unfixed, n_cams, n_poses, (s, e)).astype(bool)
    with BundleProblem(0, c_s, p_s, k_s, uv_s, n_cams, e - s, n_keys, template=template, unfixed=unf, device=device) as prob:
        prob.set_param_string(par)
        if world > 1:
            install_nccl_allreduce(prob, group)
        _, stats = prob.lm_solve(par[unf], max_iter=max_iter, ftol=ftol, xtol=xtol, gtol=gtol, lambda0=lambda0, verbose=verbose)
        local = prob.get_param_string()
    out = params.copy()
    C15 = 15 * n_cams
    out[:C15] = local[:C15]
    if world > 1:
        # pose rows travel as one padded all-gather (ranges differ by at most a few poses)
        width = max(r[1] - r[0] for r in ranges) * 6
        mine = torch.zeros(width, dtype=torch.float64, device=f"cuda:{device}")
        mine[:6 * (e - s)] = torch.from_numpy(local[C15:]).to(mine.device)
        parts = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine, group=group)
        for r, (rs, re) in enumerate(ranges):
            out[C15 + 6 * rs:C15 + 6 * re] = parts[r][:6 * (re - rs)].cpu().numpy()
    else:
        out[C15:C15 + 6 * n_poses] = local[C15:]
    return out, stats
