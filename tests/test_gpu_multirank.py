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


@pytest.mark.timeout(600)
def test_cpp_mirror_sharded_sweep(hsl, orc, tmp_path):
    """hsl::modelplayer::measure_cot_sweep with a shard (one process per GPU, hsl_nccl_* + hsl_allgather_costs_host): both
    ranks end up with the whole sweep, equal to the oracle's."""
    import re
    import numpy as np
    import torch
    from conftest import MODELS, PRESETS, model_xml
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (run with gpurun --gpus 2)")
    from hslabs_b200 import build
    lib = build.build()
    exe = str(tmp_path / "host_mirror_demo")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "host_mirror_demo.cpp"),
                           "-L", os.path.dirname(lib), "-lhsl_b200", "-Wl,-rpath," + os.path.dirname(lib), "-o", exe])
    idf = str(tmp_path / "nccl_id.bin")
    procs = []
    for r in range(2):
        env = dict(os.environ, HSL_RANK=str(r), HSL_WORLD="2", HSL_NCCL_ID_FILE=idf, HSL_GATHER_FILES="1", HSL_DEVICE=str(r))   # both GPUs visible (peer mapping)
        procs.append(subprocess.Popen([exe, PRESETS, MODELS, "8"], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=500) for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    params, name = orc.load_preset(PRESETS, 8)
    _, cots = orc.Model(model_xml(name)).measure_cot_sweep(params, 20, "period", 3, 18, 15)
    for r, (so, _) in enumerate(outs):
        got = np.array([float(v) for v in re.search(r"sharded sweep rank %d:(.*)" % r, so).group(1).split()])
        assert got.shape == (16,) and np.abs(got - cots).max() <= 1e-9 * np.abs(cots).max()
        peer = np.array([float(v) for v in re.search(r"peer-gathered sweep rank %d:(.*)" % r, so).group(1).split()])
        assert np.array_equal(peer, got)   # NVLink peer-memory gather: the same costs as through NCCL
