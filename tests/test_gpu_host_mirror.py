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
    # liksolver / orient_torso / dynpart seams (lik.h:46-56, model.h:122-130, dynrec.h:29-72)
    cons = m.constants()
    q0 = np.zeros(m.config_dim); q0[:6] = params[:6]
    A0, _ = m.fk(q0)
    hips = np.array([float(v) for v in re.search(r"^hips:(.*)$", out.stdout, re.M).group(1).split()]).reshape(-1, 3)
    for l, top in enumerate(cons["limb_top"]):
        assert np.abs(hips[l] - A0[top][12:15]).max() < 1e-12
    assert abs(float(re.search(r"rcap = (\S+)", out.stdout).group(1)) - m.rcap) < 1e-15
    rec4 = m.gait_rec(params, 0.4)
    rc, q4 = m.ik(rec4)
    pl = np.array([float(v) for v in re.search(r"^place_limbs:(.*)$", out.stdout, re.M).group(1).split()])
    d = np.abs(pl - q4); d[6:] = np.minimum(d[6:], np.abs(d[6:] - 2 * np.pi))
    assert rc == 0 and d.max() < 1e-12
    rec5 = rec4.copy(); rec5[9] += 0.05; rec5[11] += 0.03
    rc, q5 = m.ik(rec5)
    want = q4.copy(); want[9:12] = q5[9:12]                    # only limb 1 moves
    pl1 = np.array([float(v) for v in re.search(r"^place_limb:(.*)$", out.stdout, re.M).group(1).split()])
    d = np.abs(pl1 - want); d[6:] = np.minimum(d[6:], np.abs(d[6:] - 2 * np.pi))
    assert rc == 0 and d.max() < 1e-12
    A5, J5 = m.fk(want)
    dp = np.array([float(v) for v in re.search(r"^dynparts:(.*)$", out.stdout, re.M).group(1).split()]).reshape(m.n, 13)
    assert np.array_equal(dp[:, 0], np.arange(m.n)) and np.array_equal(dp[:, 1], cons["parent"]) and (dp[:, 2] == 1).all()
    assert np.array_equal(np.flatnonzero(dp[:, 3]), np.sort(cons["limb_foot"]))
    for i in range(m.n):
        Ag = A5[i].reshape(4, 4).T
        com = Ag @ np.append(cons["A_body_geom"][i][12:15], 1.0)
        assert np.abs(dp[i, 4:7] - com[:3]).max() < 1e-12
        jointed = cons["jkind"][i] != 0
        jp = J5[i][12:15] if jointed else A5[i][12:15]
        assert np.abs(dp[i, 7:10] - jp).max() < 1e-12
        ax = J5[i][8:11] if jointed else np.zeros(3)
        assert np.abs(dp[i, 10:13] - ax).max() < 1e-12
    feet = np.array([float(v) for v in re.search(r"^feet:(.*)$", out.stdout, re.M).group(1).split()]).reshape(-1, 3)
    for l, fb in enumerate(cons["limb_foot"]):
        fp = A5[fb].reshape(4, 4).T @ np.append(cons["capsule_to_pos"][fb], 1.0)
        assert np.abs(feet[l] - fp[:3]).max() < 1e-12
    assert float(re.search(r"total mass = (\S+)", out.stdout).group(1)) == m.n
