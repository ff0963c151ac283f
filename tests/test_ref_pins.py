"""The oracle restatement pinned against the reference's OWN code (oracle/_ref: /root/reference's sources compiled
unmodified against shim headers for ODE / drawstuff / rapidxml / Eigen; oracle/Makefile target `ref`).

Every stage of the path is compared where the reference exposes it: load-time constants, FK, IK, candidate
construction, frame records, per-frame dynrecord fields, x / z / motor torques, work, COT, contact statistics,
solve_forces, the rec_transform branch, curved gaits, the 1-D sweep of main.cpp:69 and the error behaviour.
Tolerance 1e-12 relative to the max-norm of each quantity (observed: 0 ... 5e-16); the linear-algebra back end is
shared between the two (oracle/orc_linalg.hpp) and is cross-checked against LAPACK in tests/test_oracle_lapack.py.
"""
import numpy as np
import pytest

from conftest import PRESETS, model_xml, ref_presets, ref_xml, rel_err

TOL = 1e-12
_models = {}


def pair(orc, refb, name):
    if name not in _models:
        _models[name] = (orc.Model(ref_xml(name)), refb.Model(ref_xml(name)))
    return _models[name]


@pytest.mark.parametrize("name", ["myant", "hexapod", "spider"])
def test_model_constants(orc, refb, name):
    """The reference's own loader (model.cpp:224-289, visualization.cpp:442-504, lik.cpp:44-78) against the oracle's."""
    o, r = pair(orc, refb, name)
    assert (o.n, o.nf, o.nmj, o.config_dim, o.rcap) == (r.n, r.nf, r.nmj, r.config_dim, r.rcap)
    assert_same_constants(o.constants(), r.constants())


def assert_same_constants(a, b):
    jointed = b["jkind"] != 0
    assert np.array_equal(a["jkind"] != 0, jointed)
    for k in ("parent", "A_pj_body", "A_body_geom", "limb_top", "limb_foot"):
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(a["J_A_parent"][jointed], b["J_A_parent"][jointed])  # bodies without a joint have none
    feet = b["limb_foot"]  # capsule_to_pos is only set for capsule geoms (visualization.cpp:495-504) and only read for feet
    assert np.array_equal(a["capsule_to_pos"][feet], b["capsule_to_pos"][feet])


@pytest.mark.parametrize("name", ["myant", "hexapod", "spider"])
def test_generated_xml_gives_reference_constants(orc, refb, name):
    """The repo's generated model files, read by the reference's loader, equal the reference's own XMLs."""
    import os
    if not os.path.exists("/root/reference/%s.xml" % name):
        pytest.skip("/root/reference not mounted")
    assert_same_constants(refb.Model("/root/reference/%s.xml" % name).constants(), refb.Model(model_xml(name)).constants())


def test_preset_parser(orc, refb):
    """get_rec_str + get_pgs_config_params (player.cpp:170-208,230-244) on every preset row with a model in the tree."""
    r = refb.Model(ref_xml("hexapod"))
    for pid in range(28):
        po, no = orc.load_preset(ref_presets(), pid)
        pr, nr = r.load_preset(ref_presets(), pid)
        assert no == nr and np.array_equal(po, pr), pid
    with pytest.raises(KeyError):
        r.load_preset(ref_presets(), 999)  # the reference prints ERROR and exit(1)s


@pytest.mark.parametrize("name", ["myant", "hexapod", "spider"])
def test_fk_ik_records(orc, refb, name):
    o, r = pair(orc, refb, name)
    rng = np.random.default_rng(5)
    for _ in range(5):
        q = rng.uniform(-1, 1, o.config_dim)
        Ao, Jo = o.fk(q)
        Ar, Jr = r.fk(q)
        jointed = r.constants()["jkind"] != 0
        assert np.abs(Ao - Ar).max() < 1e-15 and np.abs(Jo[jointed] - Jr[jointed]).max() < 1e-15
    pid = {"myant": 1, "hexapod": 8, "spider": 24}[name]
    params, _ = orc.load_preset(ref_presets(), pid)
    so, sr = o.gait_setup(params), r.gait_setup(params)
    for a, b in zip(so[:3], sr[:3]):
        assert np.array_equal(a, b)
    assert np.array_equal(so[3][:2], sr[3][:2])  # (t_step, v); max_radius is only set for curved gaits (pergen.cpp:140-157)
    for t in (0.0, 0.41, 1.3, 2.9, 7.7):
        reco, recr = o.gait_rec(params, t), r.gait_rec(params, t)
        assert np.array_equal(reco, recr)
        (rco, qo), (rcr, qr) = o.ik(reco), r.ik(recr)
        assert rco == rcr == 0 and np.array_equal(qo, qr)


@pytest.mark.parametrize("pid", [0, 1, 2, 8, 9, 15, 24])
@pytest.mark.parametrize("n_t", [20, 150])
def test_presets_equal_reference(orc, refb, pid, n_t):
    """x, z, tau, traj, work of the oracle equal the reference build on these presets."""
    params, name = orc.load_preset(ref_presets(), pid)
    o, r = pair(orc, refb, name[:-4])
    a, b = o.measure_cot(params, n_t, detail=True), r.measure_cot(params, n_t, detail=True)
    assert a["status"] == b["status"] == 0
    assert np.array_equal(a["traj"], b["traj"])
    for k in ("x", "z", "tau"):
        assert rel_err(a[k], b[k]) < TOL, (k, rel_err(a[k], b[k]))
    for k in ("work", "cot"):
        assert abs(a[k] - b[k]) <= TOL * abs(b[k]), k
    # the reference's statistics run over ALL feet (periodic.cpp:347-357), swing-foot round-off included
    assert abs(a["min_cfz"] - b["min_cfz"]) <= 1e-9 * max(1.0, abs(b["min_cfz"]))
    assert abs(a["max_mu"] - b["max_mu"]) <= 1e-6 * abs(b["max_mu"])


@pytest.mark.parametrize("pid", [3, 7, 10, 17, 20, 23, 25, 26, 27])
def test_other_presets_costs(orc, refb, pid):
    params, name = orc.load_preset(ref_presets(), pid)
    o, r = pair(orc, refb, name[:-4])
    a, b = o.measure_cot(params, 20), r.measure_cot(params, 20)
    assert a["status"] == b["status"] == 0
    assert abs(a["cot"] - b["cot"]) <= TOL * abs(b["cot"]) and abs(a["work"] - b["work"]) <= TOL * abs(b["work"])


@pytest.mark.parametrize("pid", [1, 8, 24])
def test_frame_fields_and_complete_traj(orc, refb, pid):
    """dynrecord contents of the solved frames (dynrec.cpp:134-224) and the traj.txt record incl. the q-dot columns
    (periodic.cpp:261-282,408-426)."""
    params, name = orc.load_preset(ref_presets(), pid)
    o, r = pair(orc, refb, name[:-4])
    fo, fr = o.frame_fields(params, 20), r.frame_fields(params, 20)
    assert np.array_equal(fo["contacts"], fr["contacts"])
    for k in ("pos", "jpos", "jzaxis", "fpos"):
        assert np.abs(fo[k] - fr[k]).max() < 1e-15, k
    for k in ("mom_rate", "ang_mom_rate"):
        assert rel_err(fo[k], fr[k]) < TOL, k
    b = r.measure_cot(params, 20, detail=True)
    a = o.measure_cot(params, 20, detail=True)
    cd, nmj, n_t = o.config_dim, o.nmj, 20
    comp = b["complete"]
    dt = params[7] / n_t
    for i in range(n_t):
        tsi = i + n_t if i < 2 else i
        assert np.array_equal(comp[i, :cd], a["traj"][tsi])
        d = a["traj"][tsi + 1] - a["traj"][tsi - 1]
        d = np.where(d > np.pi, d - 2 * np.pi, np.where(d < -np.pi, d + 2 * np.pi, d))
        assert np.abs(comp[i, cd:2 * cd] - d / (2 * dt)).max() < 1e-12
        assert rel_err(comp[i, 2 * cd:], a["tau"][(tsi - 2) % n_t]) < TOL


def test_sweep_of_main_cpp(orc, refb):
    """main.cpp:69: measure_cot_sweep(pgs, 20, "period", 3, 18, 15) on preset 8, parsed from the reference's stdout."""
    params, name = orc.load_preset(ref_presets(), 8)
    o, r = pair(orc, refb, "hexapod")
    vo, co = o.measure_cot_sweep(params, 20, "period", 3, 18, 15)
    vr, cr = r.measure_cot_sweep(params, 20, "period", 3, 18, 15)
    assert np.allclose(vo, vr, rtol=0, atol=1e-13)
    assert rel_err(co, cr) < TOL
    assert np.array_equal(np.argsort(co), np.argsort(cr))


@pytest.mark.parametrize("pid,tr,ea", [(8, (0, 0, 0), (0, 0, -1.571)), (1, (0.1, -0.2, 0.0), (0, 0, 0.6)),
                                       (24, (0, 0, -0.01), (0.02, -0.015, 0.3))])
def test_rec_transform_branch(orc, refb, pid, tr, ea):
    """pergensetup::set_rec_transform / transform_rec (pergen.cpp:309-342)."""
    params, name = orc.load_preset(ref_presets(), pid)
    o, r = pair(orc, refb, name[:-4])
    a = o.measure_cot(params, 20, detail=True, rec_transform=(tr, ea))
    b = r.measure_cot(params, 20, detail=True, rec_transform=(tr, ea))
    # the tilted spider case has frames with ONE foot on the ground: the reference dies in Eigen's comma-initialiser
    # assertion there (ftsolver.cpp:222-223); both report a breakdown (status 2)
    assert a["status"] == b["status"] == (2 if pid == 24 else 0)
    if b["status"] == 0:
        assert np.abs(a["traj"] - b["traj"]).max() < 1e-14
        for k in ("x", "z", "tau"):
            assert rel_err(a[k], b[k]) < 1e-10, k
        assert abs(a["cot"] - b["cot"]) <= 1e-10 * abs(b["cot"])


def test_curved_and_shifted_candidates(orc, refb):
    """Random candidates with curvature and lateral / radial foot shifts (pergen.cpp:160-198,386-397,476-507)."""
    rng = np.random.default_rng(11)
    n_ok = 0
    for name in ("myant", "hexapod", "spider"):
        o, r = pair(orc, refb, name)
        for i in range(6):
            z0 = rng.uniform(0, 0.1) if name == "spider" else rng.uniform(-0.15, -0.05)
            p = orc.make_params(torso_pos=(0, 0, z0), torso_angles=(0, 0, rng.uniform(-0.3, 0.3) if i % 2 else 0.0),
                                step_duration=rng.uniform(0, 1), period=rng.uniform(1, 6), step_length=rng.uniform(0.1, 0.5),
                                step_height=rng.uniform(0.02, 0.12), curvature=rng.uniform(-0.1, 0.1),
                                shift_type=i % 3 - 1, shift_value=rng.uniform(0.3, 0.5) if name == "spider" else rng.uniform(0, 0.3))
            a, b = o.measure_cot(p, 24, detail=True), r.measure_cot(p, 24, detail=True)
            assert (a["status"] == 0) == (b["status"] == 0), (name, i, a["status"], b["status"])
            if b["status"] != 0:
                continue
            n_ok += 1
            assert np.abs(a["traj"] - b["traj"]).max() < 1e-13
            for k in ("x", "z", "tau"):
                assert rel_err(a[k], b[k]) < 1e-9, (name, i, k)
            assert abs(a["cot"] - b["cot"]) <= 1e-9 * abs(b["cot"])
    assert n_ok >= 9


def test_unreachable_target_and_ignore_reach(orc, refb):
    """lik.cpp:142-164: an unreachable foot target makes the reference print ERROR and exit(1) unless the global
    ignore_reach flag is set, in which case the limb is stretched towards the target (the foot then hangs above the
    ground and drops out of the contact set)."""
    o, r = pair(orc, refb, "hexapod")
    for kw in (dict(step_length=1.5, step_duration=0.5), dict(step_length=1.45, step_duration=0.2)):
        p = orc.make_params(torso_pos=(0, 0, -0.1), period=3.0, step_height=0.1, **kw)
        a, b = o.measure_cot(p, 20), r.measure_cot(p, 20)
        assert a["status"] == 1 and b["status"] == 1 and np.isnan(b["cot"])
        try:
            o.set_ignore_reach(True)
            r.set_ignore_reach(True)
            a, b = o.measure_cot(p, 20, detail=True), r.measure_cot(p, 20, detail=True)
            assert a["status"] == b["status"] == 0
            assert np.abs(a["traj"] - b["traj"]).max() < 1e-13
            for k in ("x", "z", "tau"):
                assert rel_err(a[k], b[k]) < 1e-9, k
            assert abs(a["cot"] - b["cot"]) <= 1e-9 * abs(b["cot"])
        finally:
            o.set_ignore_reach(False)
            r.set_ignore_reach(False)


def test_no_foot_on_the_ground_breaks_the_reference(orc, refb):
    """With ignore_reach and the torso too high every foot hangs in the air.  The reference then runs into Eigen's
    comma-initialiser assertion (ftsolver.cpp:222-223 with 0-column matrices) and aborts; the oracle reports a
    breakdown (status 2), the product flags HSL_ST_FEW_CONTACTS."""
    o, r = pair(orc, refb, "hexapod")
    p = orc.make_params(torso_pos=(0, 0, 0.0), period=3.0, step_length=1.3, step_height=0.1)
    try:
        o.set_ignore_reach(True)
        r.set_ignore_reach(True)
        a, b = o.measure_cot(p, 20), r.measure_cot(p, 20)
        assert a["status"] == 2 and b["status"] == 2
    finally:
        o.set_ignore_reach(False)
        r.set_ignore_reach(False)


@pytest.mark.parametrize("pid", [1, 8, 24])
def test_solve_forces_and_test_dynamics(orc, refb, pid):
    """forcetorquesolver::solve_forces (ftsolver.cpp:331-378) and the sequence of modelplayer::test_dynamics
    (playerexperim.cpp:95-121): the contact forces recovered from the solved torques reproduce the solved ones."""
    params, name = orc.load_preset(ref_presets(), pid)
    o, r = pair(orc, refb, name[:-4])
    rco, cfo, cf1o, tauo = o.test_dynamics(params)
    rcr, cfr, cf1r, taur = r.test_dynamics(params)
    assert rco == rcr == 0
    assert rel_err(cfo, cfr) < TOL and rel_err(tauo, taur) < TOL and rel_err(cf1o, cf1r) < 1e-10
    if pid != 1:  # the reference's own "s = ..." check; with two feet down (trot) the torso wrench is not zero and s is not either
        assert np.linalg.norm(cfr - cf1r) < 1e-9 * np.linalg.norm(cfr)
    rng = np.random.default_rng(3)
    tau = rng.normal(size=(20, o.nmj))
    assert rel_err(o.solve_forces_frames(params, 20, tau), r.solve_forces_frames(params, 20, tau)) < 1e-10


def test_external_trajectory_entry(orc, refb):
    params, name = orc.load_preset(ref_presets(), 8)
    o, r = pair(orc, refb, "hexapod")
    n_t = 32
    traj = o.measure_cot(params, n_t, detail=True)["traj"]
    a, b = o.eval_trajectory(traj, n_t, params[7] / n_t), r.eval_trajectory(traj, n_t, params[7] / n_t)
    assert a["status"] == b["status"] == 0
    for k in ("x", "z", "tau"):
        assert rel_err(a[k], b[k]) < TOL
    assert abs(a["work"] - b["work"]) <= TOL * abs(b["work"])


def test_reference_lik_self_check(refb):
    """liksolver::solver_test (lik.cpp:123-128,371-404), the reference's own IK round trip; it exit(1)s on failure."""
    for name in ("myant", "hexapod"):
        assert refb.Model(ref_xml(name)).lik_solver_test(200) == 0


def test_batch_on_worker_processes(orc, refb):
    """ref_eval_batch: forked workers, a failing candidate (reference exit(1)) costs only itself."""
    o, r = pair(orc, refb, "hexapod")
    rng = np.random.default_rng(9)
    ps = []
    for i in range(10):
        ps.append(orc.make_params(torso_pos=(0, 0, 0.3 if i == 4 else rng.uniform(-0.15, -0.05)), step_duration=rng.uniform(0, 1),
                                  period=rng.uniform(1, 6), step_length=rng.uniform(0.1, 0.5), step_height=rng.uniform(0.02, 0.12)))
    ps = np.array(ps)
    a, b = o.eval_batch(ps, 20, nthreads=2), r.eval_batch(ps, 20, nthreads=3)
    assert b["status"][4] == 1 and np.isnan(b["cot"][4])
    ok = np.arange(10) != 4
    assert np.all(b["status"][ok] == 0) and np.all(a["status"][ok] == 0)
    assert rel_err(a["cot"][ok], b["cot"][ok]) < TOL
