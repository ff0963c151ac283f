"""CPU tier: the kernel's per-thread phase functions (hslabs_b200/csrc/hsl_frame.h), compiled for the host and run
as a serial emulation of the CUDA block (tests/hostcheck), against the CPU oracle.  This proves the restructured
algorithm (Newton-Euler recursion + per-contact 3x3 blocks + 6x6 Schur solve) equals the reference-shaped one
(6n x 6n force-torque matrix, QR, null space, two-level perturbation solve) before any GPU time is spent; the
`-m gpu` tier then checks the same code running on the device through the C ABI.
"""
import numpy as np
import pytest

import hostlib
from conftest import PRESETS, model_xml, rel_err

TOL = 1e-9


def angle_err(a, b):
    d = np.abs(a - b)
    return float(np.minimum(d, np.abs(d - 2 * np.pi)).max())


@pytest.mark.parametrize("pid", [0, 1, 2, 3, 7, 8, 9, 10, 15, 17, 20, 23, 24, 25, 26, 27])
def test_presets(orc, pid):
    params, name = orc.load_preset(PRESETS, pid)
    xml = model_xml(name)
    n_t = 20
    ref = orc.Model(xml).measure_cot(params, n_t, detail=True)
    got = hostlib.eval_gaits(xml, params, n_t)
    assert ref["status"] == 0 and got["status"][0] == 0
    assert angle_err(got["traj"][0], ref["traj"][:n_t + 4]) < 1e-12
    for key in ("x", "z", "tau"):
        assert rel_err(got[key][0], ref[key]) < TOL, key
    assert abs(got["cot"][0] - ref["cot"]) <= TOL * abs(ref["cot"])
    assert abs(got["work"][0] - ref["work"]) <= TOL * abs(ref["work"])


def test_random_batch_hexapod(orc):
    """Config-2 style random candidates (small batch): costs, ranking and status agree."""
    rng = np.random.default_rng(20261018)
    n = 24
    p = np.zeros((n, 13))
    p[:, 2] = rng.uniform(-0.15, -0.05, n); p[:, 6] = rng.uniform(0, 1, n); p[:, 7] = rng.uniform(1, 6, n)
    p[:, 8] = rng.uniform(0.1, 0.5, n); p[:, 9] = rng.uniform(0.02, 0.12, n); p[:, 11] = -1
    xml = model_xml("hexapod")
    n_t = 32
    ref = orc.Model(xml).eval_batch(p, n_t, nthreads=4)
    got = hostlib.eval_gaits(xml, p, n_t)
    ok = ref["status"] == 0
    assert np.array_equal(got["status"] == 0, ok)
    assert ok.sum() >= n // 2
    assert rel_err(got["cot"][ok], ref["cot"][ok]) < TOL
    assert np.array_equal(np.argsort(got["cot"][ok]), np.argsort(ref["cot"][ok]))


def test_unreachable_and_ignore_reach(orc):
    xml = model_xml("hexapod")
    p, _ = orc.load_preset(PRESETS, 8)
    p = p.copy()
    p[2] = 0.3  # torso too high: feet cannot reach the ground
    m = orc.Model(xml)
    assert m.measure_cot(p, 20)["status"] == 1
    got = hostlib.eval_gaits(xml, p, 20)
    assert got["status"][0] & 2 and np.isnan(got["cot"][0])
    m.set_ignore_reach(True)
    ref = m.measure_cot(p, 20, detail=True)
    got = hostlib.eval_gaits(xml, p, 20, flags=1)
    # every foot hangs in the air: the reference aborts in Eigen's comma initialiser (tests/test_ref_pins.py), the oracle
    # reports the breakdown (2), the kernel math flags HSL_ST_FEW_CONTACTS (8); the IK trajectory is still comparable
    assert ref["status"] == 2 and not (got["status"][0] & 2) and (got["status"][0] & 8)
    assert angle_err(got["traj"][0], ref["traj"][:24]) < 1e-7  # legs at full extension: acos is ill-conditioned there
    # a stride too long for the stance ends only: the stretched feet leave the contact set, the rest carries the body
    q = orc.make_params(torso_pos=(0, 0, -0.1), period=3.0, step_length=1.5, step_height=0.1, step_duration=0.5)
    m.set_ignore_reach(False)
    assert m.measure_cot(q, 20)["status"] == 1
    m.set_ignore_reach(True)
    ref = m.measure_cot(q, 20, detail=True)
    got = hostlib.eval_gaits(xml, q, 20, flags=1)
    assert ref["status"] == 0 and (got["status"][0] & 15) == 0
    assert abs(got["cot"][0] - ref["cot"]) <= 1e-7 * abs(ref["cot"])


def test_bad_step_duration():
    got = hostlib.eval_gaits(model_xml("hexapod"), np.array([0, 0, -.1, 0, 0, 0, 1.5, 3, .5, .1, 0, -1, 0.]), 20)
    assert got["status"][0] & 1 and np.isnan(got["cot"][0])


@pytest.mark.parametrize("pid", [8, 9, 26])
def test_trajectory_entry(orc, pid):
    """L2 entry: supplied joint trajectories -> x, z, motor torques, work (periodic.cpp:149-160, 192-202, 285-307)."""
    params, name = orc.load_preset(PRESETS, pid)
    xml = model_xml(name)
    n_t = 20
    m = orc.Model(xml)
    ref = m.measure_cot(params, n_t, detail=True)
    dt = params[7] / n_t
    ref2 = m.eval_trajectory(ref["traj"], n_t, dt)
    got = hostlib.eval_trajectories(xml, ref["traj"][None], dt, n_t)
    for key in ("x", "z", "tau"):
        assert rel_err(got[key][0], ref2[key]) < TOL, key
    assert abs(got["work"][0] - ref2["work"]) <= TOL * abs(ref2["work"])
    assert abs(ref2["work"] - ref["work"]) <= 1e-12 * abs(ref["work"])


@pytest.mark.parametrize("pid", [8, 9, 24])
def test_frame_solve_entry(orc, pid):
    """L1 entry: populated dynrecords -> x, z, motor torques (forcetorquesolver::solve_forcetorques, ftsolver.cpp:78-102)."""
    params, name = orc.load_preset(PRESETS, pid)
    xml = model_xml(name)
    n_t = 20
    m = orc.Model(xml)
    ref = m.measure_cot(params, n_t, detail=True)
    fields = m.frame_fields(params, n_t)
    got = hostlib.solve_frames(xml, fields)
    assert (got["status"] == 0).all()
    for key in ("x", "z", "tau"):
        assert rel_err(got[key], ref[key]) < TOL, key


@pytest.mark.parametrize("model,fb,grid", [("hexapod", 64, 3), ("hexapod", 32, 2), ("myant", 64, 2), ("spider", 32, 5)])
def test_pipelined_kernel_emulation(orc, model, fb, grid):
    """The persistent software-pipelined cost-only kernel (hsl_pipe.h), emulated serially, against the plain kernel
    emulation and the oracle: costs, work, contact statistics, status."""
    rng = np.random.default_rng(11)
    n, n_t = 9, 37
    p = np.zeros((n, 13))
    p[:, 6] = rng.uniform(0, 1, n); p[:, 7] = rng.uniform(1, 6, n); p[:, 8] = rng.uniform(0.1, 0.5, n)
    p[:, 9] = rng.uniform(0.02, 0.12, n); p[:, 11] = -1
    p[:, 2] = {"hexapod": -0.1, "myant": -0.07, "spider": 0.05}[model]
    if model == "spider":
        p[:, 11] = 0; p[:, 12] = 0.4
    p[4, 2] = 0.5  # one unreachable candidate
    xml = model_xml(model)
    plain = hostlib.eval_gaits(xml, p, n_t)
    pipe = hostlib.eval_gaits_pipe(xml, p, n_t, fb=fb, grid=grid)
    assert np.array_equal(pipe["status"], plain["status"]) and pipe["status"][4] & 2
    ok = plain["status"] == 0
    for k in ("cot", "work", "min_cfz", "max_mu"):
        # max_mu = |f_xy| / f_z of the lightest-loaded foot: a ratio with a small denominator, so round-off level
        # differences of the two code paths (IEEE quotients vs reciprocal-square-root products in the IK) show more
        tol = 1e-9 if k == "max_mu" else 1e-12
        assert np.abs(pipe[k][ok] - plain[k][ok]).max() <= tol * np.abs(plain[k][ok]).max(), k
        assert np.isnan(pipe[k][~ok]).all()
    ref = orc.Model(xml).eval_batch(p, n_t, nthreads=4)
    assert np.abs(pipe["cot"][ok] - ref["cot"][ok]).max() <= TOL * np.abs(ref["cot"][ok]).max()


REC_CASES = [  # (preset, rec_transl, rec_eas): maps that keep at least two feet on the ground in every frame
    (8, (0, 0, 0), (0, 0, -1.571)),          # main.cpp:38 (commented-out call in the reference's main)
    (1, (0, 0, 0), (0, 0, -1.571)),
    (12, (0.3, -0.2, 0), (0, 0, 0.7)),
    (8, (0, 0, -0.05), (0, 0, 0.3)),         # pushed down: every foot counts as a contact in every frame
    (24, (0, 0, 0), (0.02, 0.03, 0)),        # slight tilt: contact pattern changes, gravity no longer along the legs' z
    (9, (0.1, 0.1, 0), (0, 0, 2.5)),
]


@pytest.mark.parametrize("pid,transl,eas", REC_CASES)
def test_rec_transform(orc, pid, transl, eas):
    """pergensetup::rec_transform (pergen.cpp:309-335): torso pose and foot targets of every frame mapped rigidly."""
    params, name = orc.load_preset(PRESETS, pid)
    xml = model_xml(name)
    n_t = 20
    ref = orc.Model(xml).measure_cot(params, n_t, detail=True, rec_transform=(transl, eas))
    hostlib.set_rec_transform(transl, eas)
    try:
        got = hostlib.eval_gaits(xml, params, n_t)
        pipe = hostlib.eval_gaits_pipe(xml, params, n_t)
    finally:
        hostlib.set_rec_transform()
    assert ref["status"] == 0 and got["status"][0] == 0
    assert angle_err(got["traj"][0], ref["traj"][:n_t + 4]) < 1e-12
    for key in ("x", "z", "tau"):
        assert rel_err(got[key][0], ref[key]) < TOL, key
    assert abs(got["cot"][0] - ref["cot"]) <= TOL * abs(ref["cot"])
    assert abs(pipe["cot"][0] - ref["cot"]) <= TOL * abs(ref["cot"])
    # the map changes the answer (the level-1 weights are axis components in the world frame, ftsolver.cpp:239-246),
    # so this is not a vacuous check
    plain = orc.Model(xml).measure_cot(params, n_t)
    assert abs(plain["cot"] - ref["cot"]) > 1e-6 * abs(ref["cot"])


@pytest.mark.parametrize("pid", [8, 1, 24, 12, 9, 26])
def test_forces_from_torques(orc, pid):
    """forcetorquesolver::solve_forces (ftsolver.cpp:331-378): the 6-parameter pinned-feet residual space + one 3x3
    solve per limb against the oracle's rectangular QR least squares -- for the torques of the forward solve
    (test_dynamics, playerexperim.cpp:95-121) and for arbitrary (inconsistent) torques, gait and dynrecord entries."""
    params, name = orc.load_preset(PRESETS, pid)
    xml = model_xml(name)
    n_t = 20
    om = orc.Model(xml)
    fwd = om.measure_cot(params, n_t, detail=True)
    rng = np.random.default_rng(pid)
    fields = om.frame_fields(params, n_t)
    for tau in (fwd["tau"], fwd["tau"] + rng.normal(0, 1.0, fwd["tau"].shape)):
        ref = om.solve_forces_frames(params, n_t, tau)
        got = hostlib.solve_forces_gait(xml, params, n_t, tau)
        assert got["status"][0] == 0
        assert rel_err(got["z"][0], ref) < TOL
        got2 = hostlib.solve_forces_fields(xml, fields, tau)
        assert (got2["status"] == 0).all()
        assert rel_err(got2["z"], ref) < TOL
    # test_dynamics property: with >= 3 feet on the ground the forward solve's torques give its contact forces back
    k = fields["contacts"].sum(axis=1)
    back = hostlib.solve_forces_gait(xml, params, n_t, fwd["tau"])["z"][0]
    sel = k >= 3
    if sel.any():
        assert np.abs(back[sel] - fwd["z"][sel]).max() < 1e-9 * np.abs(fwd["z"]).max()


@pytest.mark.parametrize("model,pattern", [("hexapod", 1), ("myant", 1), ("spider", 2)])
def test_axis_specialised_kernels_equal_generic(model, pattern):
    """The kernels specialised for the model's hinge-axis pattern (compile-time hinge axes, hsl_frame.h) run the same
    arithmetic as the generic ones up to the order of a few fused multiply-adds (offsets with structural zeros):
    results equal to 1e-13, detail and pipelined paths."""
    from test_gpu_parity import _random_candidates
    xml = model_xml(model)
    assert hostlib.axis_pattern(xml) == pattern
    p = _random_candidates(model, 6, 7)
    try:
        hostlib.set_axis_specialisation(True)
        a = hostlib.eval_gaits(xml, p, 24)
        ap = hostlib.eval_gaits_pipe(xml, p, 24)
        hostlib.set_axis_specialisation(False)
        b = hostlib.eval_gaits(xml, p, 24)
        bp = hostlib.eval_gaits_pipe(xml, p, 24)
    finally:
        hostlib.set_axis_specialisation(True)
    for k in ("cot", "work", "x", "z", "tau", "traj"):
        assert rel_err(a[k], b[k]) < 1e-13, k
    assert rel_err(ap["cot"], bp["cot"]) < 1e-13


@pytest.mark.parametrize("pid", [8, 9, 12, 24, 20])
def test_record_level_entries(orc, pid):
    """pergensetup::set_rec (pergen.cpp:225-239) and kinematicmodel::set_jvalues_with_lik (model.cpp:354-359) on their
    own: frame records of a candidate at arbitrary times, and the joint values of arbitrary records."""
    params, name = orc.load_preset(PRESETS, pid)
    xml = model_xml(name)
    om = orc.Model(xml)
    times = np.array([0.0, 0.137, 1.0, 2.75, 7.3, 19.9])
    got = hostlib.gait_records(xml, params, times)
    assert got["status"][0] == 0
    for k, t in enumerate(times):
        ref = om.gait_rec(params, float(t))
        assert np.abs(got["rec"][0, k] - ref).max() < 1e-12, (k, t)
    # with a rec_transform
    hostlib.set_rec_transform((0.1, -0.2, 0.0), (0, 0, 0.6))
    try:
        tr = hostlib.gait_records(xml, params, times[:2])["rec"][0]
    finally:
        hostlib.set_rec_transform()
    assert np.abs(tr - got["rec"][0, :2]).max() > 1e-3
    # IK of those records: the oracle's joint values (mod 2 pi), unreachable records flagged
    ik = hostlib.ik_records(xml, got["rec"][0])
    for k in range(len(times)):
        rc, q = om.ik(got["rec"][0, k])
        if rc:
            assert ik["status"][k] == 2
        else:
            assert ik["status"][k] == 0
            assert angle_err(ik["q"][k][None, :], q[None, :]) < 1e-12
    far = got["rec"][0, 0].copy(); far[6:9] += 5.0           # a foot five metres away
    assert hostlib.ik_records(xml, far)["status"][0] == 2
