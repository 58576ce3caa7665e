"""CPU, world_size 2 (gloo): the multi-GPU host logic -- shard by pose, sum the camera blocks.

Each rank evaluates the block normal equations of ITS pose shard (here with the CPU oracle standing in for the
device kernel; the GPU path is covered by the -m gpu tests) and the camera blocks are all-reduced exactly as
bench.py / the LM solver do.  The result must equal the unsharded evaluation: camera blocks after the all-reduce,
pose blocks and W segments owned by exactly one rank."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import oracle as orc
from pycamset_b200 import distributed as pdist
from pycamset_b200 import synthetic as syn


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _rig():
    rig = syn.make_rig(6, 17, distortion=True, seed=5, detect_prob=0.75)
    rng = np.random.default_rng(3)
    intr, extr, poses = rig.perturbed(rng)
    return rig, rig.param_string(intr, extr, poses)


def _blocks(cam, pose, key, uv, C, M, template, params):
    o = orc.Problem(0, cam, pose, key, uv, C, M, 81, template)
    pair = cam.astype(np.int64) * M + pose
    uniq, seg = np.unique(pair, return_inverse=True)
    U, gc, V, gp, W, cost = o.normal_blocks(params, seg.astype(np.int32), len(uniq))
    return U, gc, V, gp, W, cost, uniq


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rig, params = _rig()
    cam, pose, key, uv = rig.cam.numpy(), rig.pose.numpy(), rig.key.numpy(), rig.uv.numpy()
    counts = np.bincount(pose, minlength=17)
    rng_ = pdist.balanced_pose_ranges(counts, world)[rank]
    c_s, p_s, k_s, uv_s = pdist.shard_observations(cam, pose, key, uv, rng_)
    par_s = pdist.shard_param_string(params, 6, 17, rng_)
    U, gc, V, gp, W, cost, uniq = _blocks(c_s, p_s, k_s, uv_s, 6, rng_[1] - rng_[0], rig.template, par_s)
    head = torch.from_numpy(np.concatenate([U.ravel(), gc.ravel(), [cost]]))
    dist.all_reduce(head)                                  # what allreduce_camera_blocks does on the device
    n_obs = torch.tensor([c_s.shape[0]]); dist.all_reduce(n_obs)
    out[rank] = dict(head=head.numpy(), V=V, gp=gp, W=W, range=rng_, n_obs=int(n_obs.item()), local=int(c_s.shape[0]))
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_pose_sharded_blocks_sum_to_the_unsharded_evaluation():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    rig, params = _rig()
    cam, pose, key, uv = rig.cam.numpy(), rig.pose.numpy(), rig.key.numpy(), rig.uv.numpy()
    U, gc, V, gp, W, cost, uniq = _blocks(cam, pose, key, uv, 6, 17, rig.template, params)
    ref_head = np.concatenate([U.ravel(), gc.ravel(), [cost]])
    assert out[0]["n_obs"] == cam.shape[0] and out[0]["local"] + out[1]["local"] == cam.shape[0]
    assert abs(out[0]["local"] - out[1]["local"]) <= 0.2 * cam.shape[0]      # balanced by observation count
    for r in range(world):
        assert np.max(np.abs(out[r]["head"] - ref_head)) <= 1e-10 * np.max(np.abs(ref_head))
        s, e = out[r]["range"]
        assert np.max(np.abs(out[r]["V"] - V[s:e])) <= 1e-12 * np.max(np.abs(V))
        assert np.max(np.abs(out[r]["gp"] - gp[s:e])) <= 1e-12 * max(np.max(np.abs(gp)), 1e-300)
    assert out[0]["range"][1] == out[1]["range"][0] and out[0]["range"][0] == 0 and out[1]["range"][1] == 17
    # W segments: every (camera, pose) pair is owned by exactly one rank
    cam_of, pose_of = uniq // 17, uniq % 17
    for r in range(world):
        s, e = out[r]["range"]
        mine = (pose_of >= s) & (pose_of < e)
        assert out[r]["W"].shape[0] == int(mine.sum())
        order = np.argsort(cam_of[mine] * (e - s) + (pose_of[mine] - s), kind="stable")
        assert np.max(np.abs(out[r]["W"] - W[mine][order])) <= 1e-12 * np.max(np.abs(W))


def test_pose_ranges():
    assert pdist.even_pose_ranges(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    r = pdist.balanced_pose_ranges([10, 0, 0, 10, 10, 10], 2)
    assert r[0][0] == 0 and r[-1][1] == 6 and r[0][1] == r[1][0]
    p = np.arange(15 * 2 + 6 * 5, dtype=float)
    s = pdist.shard_param_string(p, 2, 5, (1, 3))
    assert s.shape[0] == 30 + 12 and s[30] == 30 + 6
