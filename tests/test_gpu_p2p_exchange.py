"""2 GPUs (skipped below 2 devices; run with `gpurun --gpus 2`): the one-shot peer-memory all-reduce of the camera blocks
(csrc/pcs_p2p.cu) against an NCCL all-reduce of the same data -- both protocols ("the data is the signal" with
sentinel-initialised slots, and fence + flag), nine exchanges each so that both slot sets are reused several times.
Bitwise identical on every rank, equal to NCCL to rounding (different summation order for world > 2 only)."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("sentinel", ["1", "0"])
def test_peer_memory_exchange_matches_nccl(sentinel):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    out = ROOT / "gpurun_out" / f"mp_p2p_out_{sentinel}.json"
    out.parent.mkdir(exist_ok=True)
    env = dict(os.environ, PYTHONPATH=str(ROOT), PCS_P2P_SENTINEL=sentinel)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "2974" + sentinel, str(ROOT / "tests" / "mp_p2p_worker.py"), str(out)]
    subprocess.run(cmd, check=True, env=env, timeout=600)
    r = json.loads(out.read_text())
    assert r["world"] == 2 and not r["timed_out"]
    assert r["rel_err"] <= 1e-14, r
