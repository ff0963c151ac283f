"""GPU tier: the C++ host mirror classes (include/hsl_host.hpp: modelplayer / pgssweeper / periodic with the
reference's signatures) driven like the reference's main.cpp, compared with the oracle."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import MODELS, PRESETS, ROOT, model_xml

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def demo(tmp_path_factory):
    from hslabs_b200 import build
    lib = build.build()
    exe = str(tmp_path_factory.mktemp("cpp") / "host_mirror_demo")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([cxx, "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "host_mirror_demo.cpp"),
                           "-L", os.path.dirname(lib), "-lhsl_b200", "-Wl,-rpath," + os.path.dirname(lib), "-o", exe])
    return exe


@pytest.mark.parametrize("pid", [8, 9])
def test_main_like_sequence(demo, orc, pid, tmp_path):
    traj_file = str(tmp_path / "traj.txt")
    out = subprocess.run([demo, PRESETS, MODELS, str(pid), traj_file], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    params, name = orc.load_preset(PRESETS, pid)
    m = orc.Model(model_xml(name))
    vals, cots = m.measure_cot_sweep(params, 20, "period", 3, 18, 15)
    got = re.findall(r"val = (\S+) COT = (\S+)", out.stdout)[:16]
    assert len(got) == 16
    assert np.allclose([float(v) for v, _ in got], vals, rtol=1e-5)
    assert np.allclose([float(c) for _, c in got], cots, rtol=1e-5)  # printed with 6 significant digits, like the reference
    ref = m.measure_cot(params, 20, detail=True)
    cot = float(re.search(r"^COT = (\S+)$", out.stdout, re.M).group(1))
    work = float(re.search(r"work = (\S+)", out.stdout).group(1))
    assert abs(cot - ref["cot"]) <= 1e-9 * abs(ref["cot"]) and abs(work - ref["work"]) <= 1e-9 * abs(ref["work"])
    cfz = np.array([float(v) for v in re.search(r"cfz:(.*)", out.stdout).group(1).split()])
    tq = np.array([float(v) for v in re.search(r"torques:(.*)", out.stdout).group(1).split()])
    assert np.abs(cfz - ref["z"][0].reshape(-1, 3)[:, 2]).max() <= 1e-9 * np.abs(ref["z"]).max()
    assert np.abs(tq - ref["tau"][0]).max() <= 1e-9 * np.abs(ref["tau"]).max()
    # second half of test_dynamics: contact forces recovered from the torques (exact with >= 3 contacts, i.e. hexapod)
    cf1 = np.array([float(v) for v in re.search(r"cf1:(.*)", out.stdout).group(1).split()])
    ref_cf1 = m.solve_forces_frames(params, 20, ref["tau"])[0]
    assert np.abs(cf1 - ref_cf1).max() <= 1e-9 * np.abs(ref["z"]).max()
    if m.nf == 6:
        assert float(re.search(r"s = (\S+)", out.stdout).group(1)) < 1e-9
    # pergensetup::set_rec / kinematicmodel::set_jvalues_with_lik on their own
    rec = np.array([float(v) for v in re.search(r"^rec:(.*)$", out.stdout, re.M).group(1).split()])
    jv = np.array([float(v) for v in re.search(r"^jvalues:(.*)$", out.stdout, re.M).group(1).split()])
    ref_rec = m.gait_rec(params, 0.7)
    assert np.abs(rec - ref_rec).max() < 1e-12
    rc, ref_q = m.ik(ref_rec)
    assert rc == 0
    dq = np.abs(jv - ref_q); dq[6:] = np.minimum(dq[6:], np.abs(dq[6:] - 2 * np.pi))
    assert dq.max() < 1e-12
    # set_rec_rotation (main.cpp:38, pergen.cpp:309-313) on the candidate and on the sweep built from it
    rec = ((0, 0, 0), (0, 0, -1.571))
    rot = float(re.search(r"rotated COT = (\S+)", out.stdout).group(1))
    ref_rot = m.measure_cot(params, 20, rec_transform=rec)["cot"]
    assert abs(rot - ref_rot) <= 1e-9 * abs(ref_rot)
    sw = [float(v) for v in re.search(r"rotated sweep:(.*)", out.stdout).group(1).split()]
    for sl, c in zip((0.3, 0.4, 0.5), sw):
        p2 = np.array(params, np.float64); p2[8] = sl
        r2 = m.measure_cot(p2, 20, rec_transform=rec)["cot"]
        assert abs(c - r2) <= 1e-9 * abs(r2)
    # traj.txt wire format (periodic.cpp:408-426, core.cpp:47-61): n_t rows of [q, qdot, tau], n_t = T/play_dt
    rows = np.loadtxt(traj_file)
    n_t = int(params[7] / 0.02 + .5)
    assert rows.shape == (n_t, 2 * m.config_dim + m.nmj)
    ref2 = m.measure_cot(params, n_t, detail=True)
    cd, dt = m.config_dim, params[7] / n_t
    for tsi in range(n_t):
        t2 = tsi + n_t if tsi < 2 else tsi
        assert np.allclose(rows[tsi, :cd], ref2["traj"][t2], rtol=2e-5, atol=2e-6)
        # q-dot columns: periodic::compute_vel_traj (periodic.cpp:261-282), central difference with the +-pi wrap
        d = ref2["traj"][t2 + 1] - ref2["traj"][t2 - 1]
        d = np.where(d > np.pi, d - 2 * np.pi, np.where(d < -np.pi, d + 2 * np.pi, d))
        assert np.allclose(rows[tsi, cd:2 * cd], d / (2 * dt), rtol=2e-5, atol=2e-6), tsi
        assert np.allclose(rows[tsi, 2 * cd:], ref2["tau"][t2 - 2], rtol=2e-5, atol=2e-6)
