"""torchrun worker of tests/test_gpu_lm_multirank.py::test_two_gpus_nccl_match_single_rank (one rank per GPU, NCCL)."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    out, iters = sys.argv[1], int(sys.argv[2])
    from pycamset_b200 import distributed as pdist
    from tests.test_gpu_lm_multirank import C, K, M, _rig
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    rig, params = _rig()
    unfixed = np.ones(params.shape[0], bool)
    unfixed[15 * C:15 * C + 6] = False
    full, st = pdist.lm_solve_sharded(rig.cam.numpy(), rig.pose.numpy(), rig.key.numpy(), rig.uv.numpy(), C, M, K,
                                      rig.template, params, unfixed, device=local, max_iter=iters, ftol=0.0, xtol=0.0, gtol=0.0)
    if dist.get_rank() == 0:
        np.savez(out, x=full[unfixed], cost_final=st["cost_final"], iterations=st["iterations"])
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
