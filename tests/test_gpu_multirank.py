"""GPU tier, needs >= 2 devices: the NCCL path of the candidate search (search.DeviceEvaluator, top_k, cem_search)
with one process per GPU, against the single-GPU evaluation of the same batch.  Skipped on a 1-GPU box."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.timeout(600)
def test_two_rank_nccl_search(hsl, tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "dist_nccl_worker.py"), str(tmp_path)]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=550)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert (tmp_path / "rank0.ok").exists() and (tmp_path / "rank1.ok").exists()
