"""GPU tier: the fall / perturbation sweep (BASELINE configs[4]).

hsl_fall_sweep_host integrates many copies of the reference's closed loop (PD position control about the evaluated
gait, one torso kick, fall check) with a batched restatement of ODE's quickstep.  The checker is the reference's own
player.cpp running on the ODE shim (oracle/shim/ode_step.cpp), in a separate process.  Both follow the same arithmetic
row for row, so short horizons agree to round-off; over long horizons the contact dynamics amplify round-off, so fall
outcomes are compared world by world with a small disagreement allowance and as statistics (fall rate, time-to-fall
histogram)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import PRESETS, ROOT, model_xml

pytestmark = pytest.mark.gpu


def cpu_side(tmp_path, xml, params, n_steps, kick_step, kick_dv, n_traj=0, nprocs=None, **kw):
    from oracle import ref
    ref.build()
    if not ref.available():
        pytest.skip("oracle/_ref not built")
    inp, outp = str(tmp_path / "in.npz"), str(tmp_path / "out.npz")
    np.savez(inp, xml=xml, params=params, n_steps=n_steps, kick_step=np.asarray(kick_step, np.int32), kick_dv=np.asarray(kick_dv, np.float64),
             play_dt=kw.get("play_dt", 0.02), t0=kw.get("t0", 0.0), hc=kw.get("hc", 0.7), tmin=kw.get("tmin", 0.1),
             nprocs=nprocs or os.cpu_count() or 4, n_traj=n_traj)
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tests", "fall_cpu_worker.py"), inp, outp], timeout=1500)
    return np.load(outp)


def test_short_horizon_follows_the_cpu_stepper(hsl, tmp_path):
    params, name = hsl.load_preset(PRESETS, 8)
    xml = model_xml(name)
    kicks = np.array([[0, 0, 0], [4, 0, 0], [0, 10, 0], [0, -7, 2], [3, 9, -1], [0, 14, 0]], float)
    ks = np.array([-1, 30, 30, 10, 45, 20], np.int32)
    n_steps = 300
    cpu = cpu_side(tmp_path, xml, params, n_steps, ks, kicks, n_traj=len(ks))
    gpu = hsl.Model(xml).fall_sweep(params, n_steps, ks, kicks, want_traj=True)
    assert (gpu["status"] == 0).all()
    for i in range(len(ks)):
        a, b = gpu["traj"][i], cpu["traj"][i]
        # Same arithmetic row for row (tests/test_hostcheck_fall.py: the kernel source run on the host agrees with the CPU
        # stepper to 1e-13), but the gait puts the stance feet exactly at touching distance, so whether a foot counts as in
        # contact in the very first steps hangs on the last bit of the device's FK: the paths separate by O(g h^2) = 4e-4
        # there and stay within a few millimetres while the robot walks.
        n_ok = int(np.isfinite(b[:, 0]).sum())
        assert np.abs(a[:60] - b[:60]).max() < 1e-2, (i, np.abs(a[:60] - b[:60]).max())
        assert np.abs(a[:n_ok] - b[:n_ok]).max() < 8e-2, i
    assert np.array_equal(gpu["fell"], cpu["fell"])
    both = gpu["fell"].astype(bool)
    assert both.any() and not both.all()          # the kicks bracket the stability limit
    assert np.abs(gpu["t_end"][both] - cpu["t_end"][both]).max() <= 0.1


def test_fall_statistics_against_the_cpu_stepper(hsl, tmp_path):
    """10240 random kicks (direction uniform in the horizontal plane, |dv| U[0, 16], kick time U[0.5 s, 3.5 s)): fall rate and
    time-to-fall histogram of the sweep kernel against the CPU stepper on the same kicks (HSL_FALL_STAT_WORLDS overrides the count)."""
    params, name = hsl.load_preset(PRESETS, 8)
    xml = model_xml(name)
    rng = np.random.default_rng(20261019)
    w, n_steps = int(os.environ.get("HSL_FALL_STAT_WORLDS", "10240")), 300
    mag, th = rng.uniform(0, 16, w), rng.uniform(0, 2 * np.pi, w)
    kicks = np.stack([mag * np.cos(th), mag * np.sin(th), np.zeros(w)], axis=1)
    ks = rng.integers(25, 175, w).astype(np.int32)
    cpu = cpu_side(tmp_path, xml, params, n_steps, ks, kicks)
    gpu = hsl.Model(xml).fall_sweep(params, n_steps, ks, kicks)
    fr_c, fr_g = cpu["fell"].mean(), gpu["fell"].mean()
    print("fall rate cpu %.4f gpu %.4f, per-world agreement %.4f, kernel %.1f ms (%.3g world-steps/s)" %
          (fr_c, fr_g, (cpu["fell"] == gpu["fell"]).mean(), gpu["kernel_ms"], w * n_steps / (gpu["kernel_ms"] * 1e-3)))
    assert 0.05 < fr_c < 0.95                      # the kick range brackets the stability limit
    assert abs(fr_c - fr_g) <= 0.02
    assert (cpu["fell"] == gpu["fell"]).mean() >= 0.97
    hc, _ = np.histogram(cpu["t_end"][cpu["fell"] == 1], bins=12, range=(0, n_steps * 0.02))
    hg, _ = np.histogram(gpu["t_end"][gpu["fell"] == 1], bins=12, range=(0, n_steps * 0.02))
    assert np.abs(hc - hg).sum() <= 0.1 * max(hc.sum(), 1)
    # no kick, no fall; a kick far beyond the limit always falls
    calm = hsl.Model(xml).fall_sweep(params, n_steps, np.full(4, -1, np.int32), np.zeros((4, 3)))
    assert not calm["fell"].any() and np.abs(calm["final_z"] - 0.8).max() < 0.02
    hard = hsl.Model(xml).fall_sweep(params, n_steps, np.full(4, 20, np.int32), np.tile([0.0, 40.0, 0.0], (4, 1)))
    assert hard["fell"].all()


@pytest.mark.parametrize("pid", [8, 1])
def test_warp_per_world_kernel_equals_the_thread_per_world_kernel(hsl, pid):
    """The two kernels of the sweep (a warp per world with the world in registers, hsl_fall_warp.cuh; a thread per world,
    hsl_fall_world.h -- the source the CPU tier emulates) state the same arithmetic in the same order: same torso paths."""
    params, name = hsl.load_preset(PRESETS, pid)
    xml = model_xml(name)
    rng = np.random.default_rng(77 + pid)
    w, n_steps = 96, 200
    mag, th = rng.uniform(0, 16, w), rng.uniform(0, 2 * np.pi, w)
    kicks = np.stack([mag * np.cos(th), mag * np.sin(th), rng.uniform(-2, 2, w)], axis=1)
    ks = rng.integers(5, 150, w).astype(np.int32)
    ks[:3] = -1
    m = hsl.Model(xml)
    m.set_fall_variant(0)
    a = m.fall_sweep(params, n_steps, ks, kicks, want_traj=True)
    m.set_fall_variant(1)
    b = m.fall_sweep(params, n_steps, ks, kicks, want_traj=True)
    d = np.abs(a["traj"] - b["traj"]).max()
    print("preset %d: max |traj difference| %.3g over %d worlds x %d steps; kernel %.2f ms (thread) vs %.2f ms (warp)" %
          (pid, d, w, n_steps, a["kernel_ms"], b["kernel_ms"]))
    assert np.array_equal(a["fell"], b["fell"]) and np.array_equal(a["status"], b["status"])
    assert np.array_equal(a["t_end"], b["t_end"])
    assert d < 1e-9
    assert a["fell"].any() and not a["fell"].all()


@pytest.mark.parametrize("pid,kick_max,hc", [(9, 16.0, 0.6), (24, 64.0, 0.4)])
def test_fall_statistics_on_the_other_models(hsl, tmp_path, pid, kick_max, hc):
    """The statistical comparison on the four-limbed myant and the spider (other hinge pattern, other body count): 2048
    random kicks each, fall rate and per-world outcomes against the CPU stepper."""
    params, name = hsl.load_preset(PRESETS, pid)
    xml = model_xml(name)
    rng = np.random.default_rng(1000 + pid)
    w, n_steps = 2048, 300
    mag, th = rng.uniform(0, kick_max, w), rng.uniform(0, 2 * np.pi, w)
    kicks = np.stack([mag * np.cos(th), mag * np.sin(th), np.zeros(w)], axis=1)
    ks = rng.integers(25, 175, w).astype(np.int32)
    cpu = cpu_side(tmp_path, xml, params, n_steps, ks, kicks, hc=hc)
    gpu = hsl.Model(xml).fall_sweep(params, n_steps, ks, kicks, hc=hc)
    fr_c, fr_g = cpu["fell"].mean(), gpu["fell"].mean()
    print("preset %d (%s): fall rate cpu %.4f gpu %.4f, per-world agreement %.4f, kernel %.1f ms" %
          (pid, name, fr_c, fr_g, (cpu["fell"] == gpu["fell"]).mean(), gpu["kernel_ms"]))
    assert (gpu["status"] == 0).all()
    assert 0.03 < fr_c < 0.97
    assert abs(fr_c - fr_g) <= 0.03
    assert (cpu["fell"] == gpu["fell"]).mean() >= 0.95


def test_fall_sweep_argument_edges(hsl, tmp_path):
    """No kicks at all (null arrays), a start inside the period (t0 > 0: initial pose from the gait's frame there), one world,
    bad arguments; both kernels agree on each, and the t0 > 0 run follows the CPU stepper."""
    params, name = hsl.load_preset(PRESETS, 8)
    xml = model_xml(name)
    m = hsl.Model(xml)
    res = {}
    for v in (0, 1):
        m.set_fall_variant(v)
        calm = m.fall_sweep(params, 80, n_worlds=3, want_traj=True)                 # kick_step = kick_dv = NULL
        assert not calm["fell"].any() and np.array_equal(calm["traj"][0], calm["traj"][2])
        late = m.fall_sweep(params, 120, [10], [[0.0, 12.0, 0.0]], t0=1.0, want_traj=True)   # one world, started at t = 1 s
        res[v] = (calm, late)
    assert np.array_equal(res[0][0]["traj"], res[1][0]["traj"]) and np.array_equal(res[0][1]["traj"], res[1][1]["traj"])
    assert np.array_equal(res[0][1]["t_end"], res[1][1]["t_end"])
    cpu = cpu_side(tmp_path, xml, params, 120, [10], [[0.0, 12.0, 0.0]], n_traj=1, t0=1.0)
    n_ok = int(np.isfinite(cpu["traj"][0][:, 0]).sum())
    assert np.abs(res[1][1]["traj"][0][:min(n_ok, 60)] - cpu["traj"][0][:min(n_ok, 60)]).max() < 1e-2
    assert bool(res[1][1]["fell"][0]) == bool(cpu["fell"][0])
    with pytest.raises(hsl.HslError):
        m.fall_sweep(params, 0, [5], [[0, 1, 0]])
    with pytest.raises(hsl.HslError):
        m.fall_sweep(params, 10, [5], [[0, 1, 0]], play_dt=0.0)
    with pytest.raises(hsl.HslError):
        m.fall_sweep(params, 10, [5], [[0, 1, 0]], t0=100.0)             # beyond the recorded trajectory
