"""2+ GPUs: latency of the peer-memory camera-block all-reduce vs NCCL, back to back on one stream (torchrun)."""
import os, sys, time
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pycamset_b200 import synthetic as syn, distributed as pdist
from pycamset_b200.problem import BundleProblem

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{lr}"))
rig = syn.make_rig(32, 8, distortion=True, seed=0, device=f"cuda:{lr}")
stream = torch.cuda.Stream(device=lr)
prob = BundleProblem(0, rig.cam, rig.pose, rig.key, rig.uv, 32, 8, 81, template=rig.template, device=lr, stream=stream.cuda_stream)
prob.set_param_string(rig.param_string())
p2p = pdist.P2PCameraAllReduce(prob)
head = pdist.tensor_from_ptr(prob.device_buffers().U, 32 * 240 + 1, lr)
res = {}
with torch.cuda.stream(stream):
    prob.normal_equations_device()
    for name, fn in (("p2p", p2p), ("nccl", lambda: dist.all_reduce(head))):
        for _ in range(20): fn()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(200):
            head.mul_(0.5)      # keep the values bounded; also separates consecutive collectives by one tiny kernel
            fn()
        e1.record(stream); torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / 200 * 1e3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(200): head.mul_(0.5)
    e1.record(stream); torch.cuda.synchronize()
    res["mul_only"] = e0.elapsed_time(e1) / 200 * 1e3
if rank == 0:
    print({k: round(v, 2) for k, v in res.items()}, "us per call")
dist.destroy_process_group()
