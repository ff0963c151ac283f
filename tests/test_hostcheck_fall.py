"""CPU tier: the fall-sweep kernel's per-world code (hslabs_b200/csrc/hsl_fall_world.h, compiled for the host by
tests/hostcheck) against the reference's own closed loop (player.cpp position control, kicks, fall check) running on the
ODE shim's world stepper (oracle/shim/ode_step.cpp) -- in a separate process, because the reference keeps one static ODE
world per process.  Both follow the same arithmetic row for row: torso trajectories agree to round-off."""
import os
import subprocess
import sys

import numpy as np
import pytest

import hostlib
from conftest import PRESETS, ROOT, model_xml


def cpu_side(tmp_path, xml, params, n_steps, kick_step, kick_dv, n_traj, nprocs=4):
    from oracle import ref
    ref.build()
    if not ref.available():
        pytest.skip("oracle/_ref not built and /root/reference not present")
    inp, outp = str(tmp_path / "in.npz"), str(tmp_path / "out.npz")
    np.savez(inp, xml=xml, params=params, n_steps=n_steps, kick_step=np.asarray(kick_step, np.int32), kick_dv=np.asarray(kick_dv, np.float64),
             play_dt=0.02, t0=0.0, hc=0.7, tmin=0.1, nprocs=nprocs, n_traj=n_traj)
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tests", "fall_cpu_worker.py"), inp, outp], timeout=900)
    return np.load(outp)


@pytest.mark.parametrize("pid", [8, 9])
def test_sweep_world_code_equals_the_cpu_stepper(orc, pid, tmp_path):
    params, name = orc.load_preset(PRESETS, pid)
    xml = model_xml(name)
    kicks = np.array([[0, 0, 0], [4, 0, 0], [0, 10, 0], [0, -7, 2], [3, 9, -1], [0, 14, 0]], float)
    ks = np.array([-1, 30, 30, 10, 45, 20], np.int32)
    n_steps = 200
    cpu = cpu_side(tmp_path, xml, params, n_steps, ks, kicks, n_traj=len(ks))
    emu = hostlib.fall_sweep(xml, params, n_steps, ks, kicks, want_traj=True)
    assert (emu["status"] == 0).all()
    assert np.array_equal(emu["fell"], cpu["fell"])
    # The gait puts the stance feet exactly at touching distance (foot z = rcap), so whether a foot counts as in contact in
    # the first step hangs on the last bit of the forward kinematics: the reference multiplies 4x4 matrices, the kernel code
    # composes rotations and offsets.  Hexapod, preset 8: every sign agrees and the two runs stay equal to round-off for the
    # whole horizon.  Myant, preset 9: one foot starts 1e-17 the other way, the runs separate by O(g h^2) = 4e-4 in step 0
    # and stay within centimetres; the outcomes agree, the fall times to two steps.
    tol_traj, tol_t = (1e-9, 1e-12) if pid == 8 else (5e-2, 0.05)
    assert np.abs(emu["t_end"] - cpu["t_end"]).max() <= tol_t
    for i in range(len(ks)):
        n_ok = min(int(np.isfinite(cpu["traj"][i][:, 0]).sum()), int(round(emu["t_end"][i] / 0.02)))
        assert n_ok >= 20
        assert np.abs(emu["traj"][i][:n_ok] - cpu["traj"][i][:n_ok]).max() < tol_traj, i
    assert cpu["fell"].any() and not cpu["fell"].all()   # the kicks bracket the stability limit
