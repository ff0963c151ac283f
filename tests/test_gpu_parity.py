"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same inputs.

Tolerances, relative to the max-norm of each quantity over the compared block, FP64:
joint values 1e-12 (modulo 2*pi), x / z / motor torques 1e-9, work / COT 1e-9, contact flags exact.
"""
import numpy as np
import pytest

from conftest import PRESETS, model_xml, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-9
PRESET_IDS = [0, 1, 2, 3, 7, 8, 9, 10, 15, 17, 20, 23, 24, 25, 26, 27]


def _angle_err(a, b):
    d = np.abs(a - b)
    d = np.minimum(d, np.abs(d - 2 * np.pi))
    return float(d.max())


@pytest.mark.parametrize("pid", PRESET_IDS)
@pytest.mark.parametrize("n_t", [20, 150])
def test_presets_detail(hsl, orc, pid, n_t):
    params, name = orc.load_preset(PRESETS, pid)
    xml = model_xml(name)
    ref = orc.Model(xml).measure_cot(params, n_t, detail=True)
    gpu = hsl.Model(xml).eval_gaits_detail(params, n_t)
    assert ref["status"] == 0 and gpu["status"][0] == 0
    assert _angle_err(gpu["traj"][0], ref["traj"][:n_t + 4]) < 1e-12
    zc = ref["z"].reshape(n_t, -1, 3)
    contacts_ref = np.abs(zc).max(axis=2) > 1e-9 * np.abs(zc).max()
    assert np.array_equal(gpu["contacts"][0].astype(bool), contacts_ref)
    for key in ("x", "z", "tau"):
        assert rel_err(gpu[key][0], ref[key]) < TOL, key
    assert abs(gpu["work"][0] - ref["work"]) <= TOL * abs(ref["work"])
    assert abs(gpu["cot"][0] - ref["cot"]) <= TOL * abs(ref["cot"])
    # contact statistics over the feet on the ground (the reference also folds in round-off of swing feet)
    lam = zc[contacts_ref]
    assert abs(gpu["min_cfz"][0] - lam[:, 2].min()) <= TOL * np.abs(lam).max()
    mu = np.sqrt(lam[:, 0] ** 2 + lam[:, 1] ** 2) / lam[:, 2]
    assert abs(gpu["max_mu"][0] - mu.max()) <= 1e-7 * max(1.0, abs(mu.max()))


def test_golden_fixtures(hsl):
    """Committed golden frames (frozen from the oracle, tests/golden/make_golden.py): no oracle call needed."""
    import os
    from conftest import GOLDEN
    for name in ("hexapod", "myant", "spider"):
        g = np.load(os.path.join(GOLDEN, "frames_%s.npz" % name))
        out = hsl.Model(model_xml(name)).eval_gaits_detail(g["params"], 20)
        assert out["status"][0] == 0
        assert _angle_err(out["traj"][0], g["traj"][:24]) < 1e-12
        for key in ("x", "z", "tau"):
            assert rel_err(out[key][0], g[key]) < TOL, (name, key)
        assert abs(out["cot"][0] - float(g["cot"])) <= TOL * abs(float(g["cot"]))


def test_golden_sweep(hsl):
    """measure_cot_sweep(pgs, 20, "period", 3, 18, 15) on preset 8 (main.cpp:69) against the frozen sweep."""
    import json
    import os
    from conftest import GOLDEN
    gold = json.load(open(os.path.join(GOLDEN, "sweep_preset8_period.json")))
    params, name = hsl.load_preset(PRESETS, 8)
    vals, cots = hsl.measure_cot_sweep(hsl.Model(model_xml(name)), params, 20, "period", 3, 18, 15)
    assert np.allclose(vals, gold["vals"], rtol=0, atol=1e-12)
    assert rel_err(cots, gold["cots"]) < TOL
    assert np.array_equal(np.argsort(cots), np.argsort(gold["cots"]))


def _random_candidates(model, n, seed):
    rng = np.random.default_rng(seed)
    p = np.zeros((n, 13))
    p[:, 6] = rng.uniform(0, 1, n); p[:, 7] = rng.uniform(1, 6, n); p[:, 8] = rng.uniform(0.1, 0.5, n)
    p[:, 9] = rng.uniform(0.02, 0.12, n); p[:, 11] = -1
    if model == "hexapod":
        p[:, 2] = rng.uniform(-0.15, -0.05, n)
    elif model == "myant":
        p[:, 2] = rng.uniform(-0.12, -0.04, n)
    else:  # spider: BASELINE config 3
        p[:, 2] = rng.uniform(0.0, 0.1, n); p[:, 11] = 0; p[:, 12] = rng.uniform(0.3, 0.5, n)
    return p


@pytest.mark.parametrize("model,n_t", [("hexapod", 64), ("myant", 48), ("spider", 64)])
def test_random_batch_costs_and_ranking(hsl, orc, model, n_t):
    """Random candidates in the BASELINE ranges: per-candidate costs within 1e-9, identical status and ranking."""
    p = _random_candidates(model, 48, 20261018)
    ref = orc.Model(model_xml(model)).eval_batch(p, n_t, nthreads=8)
    gpu = hsl.Model(model_xml(model)).eval_gaits(p, n_t)
    ok = ref["status"] == 0
    assert np.array_equal((gpu["status"] & 3) == 0, ok)
    assert ok.sum() >= 24
    assert (gpu["status"][ok] == 0).all()
    assert np.abs(gpu["cot"][ok] - ref["cot"][ok]).max() <= TOL * np.abs(ref["cot"][ok]).max()
    assert np.abs(gpu["work"][ok] - ref["work"][ok]).max() <= TOL * np.abs(ref["work"][ok]).max()
    assert np.isnan(gpu["cot"][~ok]).all()
    ro, go = np.argsort(ref["cot"][ok]), np.argsort(gpu["cot"][ok])
    assert np.array_equal(ro, go)


def test_curved_and_shifted_batch(hsl, orc):
    """BASELINE config 2, second batch: curvature U[-.1,.1], lateral foot shift U[0,.3] (curved-path branch)."""
    rng = np.random.default_rng(7)
    p = _random_candidates("hexapod", 24, 11)
    p[:, 10] = rng.uniform(-0.1, 0.1, 24); p[:, 11] = 0; p[:, 12] = rng.uniform(0, 0.3, 24)
    ref = orc.Model(model_xml("hexapod")).eval_batch(p, 40, nthreads=8)
    gpu = hsl.Model(model_xml("hexapod")).eval_gaits(p, 40)
    ok = ref["status"] == 0
    assert ok.sum() >= 8 and np.array_equal((gpu["status"] & 3) == 0, ok)
    assert np.abs(gpu["cot"][ok] - ref["cot"][ok]).max() <= TOL * np.abs(ref["cot"][ok]).max()


def test_status_codes_and_ignore_reach(hsl, orc):
    xml = model_xml("hexapod")
    p, _ = orc.load_preset(PRESETS, 8)
    bad = p.copy(); bad[2] = 0.3           # feet cannot reach the ground
    oob = p.copy(); oob[6] = 1.5           # step_duration out of [0,1]
    m = hsl.Model(xml)
    out = m.eval_gaits(np.stack([p, bad, oob]), 20)
    assert out["status"][0] == 0 and np.isfinite(out["cot"][0])
    assert out["status"][1] & 2 and np.isnan(out["cot"][1])
    assert out["status"][2] & 1 and np.isnan(out["cot"][2])
    om = orc.Model(xml)
    assert om.measure_cot(bad, 20)["status"] == 1
    om.set_ignore_reach(True)
    ref = om.measure_cot(bad, 20, detail=True)
    got = m.eval_gaits_detail(bad, 20, flags=hsl.HSL_FLAG_IGNORE_REACH)
    # every foot hangs in the air: the reference aborts in Eigen's comma initialiser (tests/test_ref_pins.py), the oracle
    # reports the breakdown (2), the library flags HSL_ST_FEW_CONTACTS
    assert ref["status"] == 2 and not (got["status"][0] & 2) and (got["status"][0] & hsl.HSL_ST_FEW_CONTACTS)
    # every leg is at (clamped) or within round-off of full extension here, where acos is infinitely ill-conditioned:
    # a 1-ulp difference in the foot target moves the knee angles by ~sqrt(ulp), so angles agree to 1e-7 only.
    assert _angle_err(got["traj"][0], ref["traj"][:24]) < 1e-7
    # a stride too long for the stance ends only: the stretched feet leave the contact set, the rest carries the body
    q = orc.make_params(torso_pos=(0, 0, -0.1), period=3.0, step_length=1.5, step_height=0.1, step_duration=0.5)
    om.set_ignore_reach(False)
    assert om.measure_cot(q, 20)["status"] == 1 and m.eval_gaits(q[None], 20)["status"][0] & 2
    om.set_ignore_reach(True)
    ref = om.measure_cot(q, 20, detail=True)
    got = m.eval_gaits(q[None], 20, flags=hsl.HSL_FLAG_IGNORE_REACH)
    assert ref["status"] == 0 and (got["status"][0] & 15) == 0
    assert abs(got["cot"][0] - ref["cot"]) <= 1e-7 * abs(ref["cot"])


@pytest.mark.parametrize("pid", [8, 9, 26])
def test_trajectory_entry(hsl, orc, pid):
    """hsl_eval_trajectories_host: supplied joint trajectories -> x, z, tau, work."""
    params, name = orc.load_preset(PRESETS, pid)
    xml = model_xml(name)
    n_t = 20
    om = orc.Model(xml)
    ref = om.measure_cot(params, n_t, detail=True)
    dt = params[7] / n_t
    ref2 = om.eval_trajectory(ref["traj"], n_t, dt)
    got = hsl.Model(xml).eval_trajectories(np.stack([ref["traj"], ref["traj"]]), dt, n_t)
    for c in range(2):
        for key in ("x", "z", "tau"):
            assert rel_err(got[key][c], ref2[key]) < TOL, key
        assert abs(got["work"][c] - ref2["work"]) <= TOL * abs(ref2["work"])


@pytest.mark.parametrize("pid", [8, 9, 24])
def test_frame_solve_entry(hsl, orc, pid):
    """hsl_solve_frames_host: populated dynrecords -> x, z, motor torques (the narrowest entry)."""
    params, name = orc.load_preset(PRESETS, pid)
    xml = model_xml(name)
    om = orc.Model(xml)
    ref = om.measure_cot(params, 20, detail=True)
    f = om.frame_fields(params, 20)
    got = hsl.Model(xml).solve_frames(f["pos"], f["jpos"], f["jzaxis"], f["mom_rate"], f["ang_mom_rate"], f["fpos"], f["contacts"])
    assert (got["status"] == 0).all()
    for key in ("x", "z", "tau"):
        assert rel_err(got[key], ref[key]) < TOL, key


def test_frame_solve_ragged_sizes(hsl, orc):
    """Frame counts that do not fill a block (1, 31, 33 frames) give the same per-frame results."""
    params, name = orc.load_preset(PRESETS, 8)
    xml = model_xml(name)
    f = orc.Model(xml).frame_fields(params, 40)
    m = hsl.Model(xml)
    full = m.solve_frames(f["pos"], f["jpos"], f["jzaxis"], f["mom_rate"], f["ang_mom_rate"], f["fpos"], f["contacts"])
    for k in (1, 31, 33):
        part = m.solve_frames(*(f[a][:k] for a in ("pos", "jpos", "jzaxis", "mom_rate", "ang_mom_rate", "fpos", "contacts")))
        assert np.array_equal(part["x"], full["x"][:k]) and np.array_equal(part["tau"], full["tau"][:k])


def test_kernel_variants_agree(hsl):
    """All occupancy variants of the cost-only kernel and the detail kernel agree to round-off (they are separate
    compilations of the same source: instruction scheduling / FMA contraction may differ in the last bits)."""
    p = _random_candidates("hexapod", 96, 3)
    m = hsl.Model(model_xml("hexapod"))
    p[5, 2] = 0.5  # one unreachable candidate: status and NaNs must agree too
    det = m.eval_gaits_detail(p, 37)
    ref = det["cot"]
    for fb, mr in ((32, 255), (32, 128), (32, 96), (64, 128), (64, 1), (32, 1)):  # maxreg 1 = pipelined persistent kernel
        m.set_tuning(fb, mr)
        full = m.eval_gaits(p, 37)
        got = full["cot"]
        assert np.array_equal(full["status"], det["status"]), (fb, mr)
        for k in ("work", "min_cfz", "max_mu"):
            okk = det["status"] == 0
            tol = 1e-9 if k == "max_mu" else 1e-12  # max_mu is a ratio with a small denominator
            assert np.abs(full[k][okk] - det[k][okk]).max() <= tol * np.abs(det[k][okk]).max(), (fb, mr, k)
        assert np.array_equal(np.isnan(got), np.isnan(ref)), (fb, mr)
        assert np.nanmax(np.abs(got - ref) / np.abs(ref)) < 1e-12, (fb, mr)


def test_full_size_properties(hsl, orc):
    """BASELINE config 2 at full size (4096 x 256): determinism, batch-order invariance, agreement with the oracle on
    a sampled subset, ranking of the sample identical, static-limit rows carry the weight."""
    n, n_t = 4096, 256
    p = _random_candidates("hexapod", n, 20261018)
    p[:8, 8] = 1e-6; p[:8, 9] = 1e-6        # static-limit rows
    m = hsl.Model(model_xml("hexapod"))
    a = m.eval_gaits(p, n_t)
    b = m.eval_gaits(p, n_t)
    for k in ("cot", "work", "min_cfz", "max_mu"):
        assert np.array_equal(a[k], b[k], equal_nan=True)          # bit-reproducible run to run
    perm = np.random.default_rng(0).permutation(n)
    c = m.eval_gaits(p[perm], n_t)
    assert np.array_equal(c["cot"], a["cot"][perm], equal_nan=True)  # a candidate's cost does not depend on its neighbours
    ok = np.where(a["status"] == 0)[0]
    assert ok.size > n // 2
    sample = ok[np.linspace(0, ok.size - 1, 12).astype(int)]
    ref = orc.Model(model_xml("hexapod")).eval_batch(p[sample], n_t, nthreads=12)
    assert (ref["status"] == 0).all()
    assert np.abs(a["cot"][sample] - ref["cot"]).max() <= TOL * np.abs(ref["cot"]).max()
    assert np.array_equal(np.argsort(a["cot"][sample]), np.argsort(ref["cot"]))
    # static limit: work vanishes, every contact force is a share of the weight (min_cfz > 0)
    st = [i for i in range(8) if a["status"][i] == 0]
    assert len(st) >= 4 and (a["min_cfz"][st] > 0).all() and (a["work"][st] < 1e-4).all()
    # unreachable candidates report NaN and a status bit, never garbage
    badrows = a["status"] & 3 != 0
    assert np.isnan(a["cot"][badrows]).all() and np.isfinite(a["cot"][~badrows]).all()


def test_config3_shard_size_properties(hsl, orc):
    """BASELINE config 3, one GPU's shard: spider.xml, 8192 candidates x 512 frames (4.2 M frame solves): runs,
    is reproducible, agrees with the oracle on a sampled subset and ranks it identically."""
    n, n_t = 8192, 512
    p = _random_candidates("spider", n, 20261019)
    m = hsl.Model(model_xml("spider"))
    a = m.eval_gaits(p, n_t)
    b = m.eval_gaits(p, n_t)
    assert np.array_equal(a["cot"], b["cot"], equal_nan=True)
    ok = np.where(a["status"] == 0)[0]
    assert ok.size > n // 2
    sample = ok[np.linspace(0, ok.size - 1, 8).astype(int)]
    ref = orc.Model(model_xml("spider")).eval_batch(p[sample], n_t, nthreads=8)
    assert (ref["status"] == 0).all()
    assert np.abs(a["cot"][sample] - ref["cot"]).max() <= TOL * np.abs(ref["cot"]).max()
    assert np.abs(a["work"][sample] - ref["work"]).max() <= TOL * np.abs(ref["work"]).max()
    assert np.array_equal(np.argsort(a["cot"][sample]), np.argsort(ref["cot"]))


def test_small_and_ragged_batches(hsl, orc):
    """Edge sizes: one candidate, n_t = 1 and 2 (fewer frames than the finite-difference stencil is wide), batches that
    do not fill a block, many short candidates per block."""
    xml = model_xml("hexapod")
    p, _ = orc.load_preset(PRESETS, 8)
    m = hsl.Model(xml)
    om = orc.Model(xml)
    for n_t in (1, 2, 3, 5, 27, 29, 61):
        ref = om.measure_cot(p, n_t)
        got = m.eval_gaits(p, n_t)
        assert got["status"][0] == 0 and ref["status"] == 0
        # n_t = 1, 2: every frame is a whole number of periods apart, the joints do not move and the work is 0 up to
        # round-off (FMA contraction on the device leaves ~1e-17 where the host gets an exact 0)
        assert abs(got["cot"][0] - ref["cot"]) <= TOL * max(abs(ref["cot"]), 1e-3), n_t
    batch = np.tile(p, (37, 1))
    batch[:, 7] = np.linspace(2.0, 6.0, 37)
    ref = om.eval_batch(batch, 6, nthreads=8)
    got = m.eval_gaits(batch, 6)      # 37 candidates x 10 slots: several candidates per block
    assert np.abs(got["cot"] - ref["cot"]).max() <= TOL * np.abs(ref["cot"]).max()


def test_pinned_and_preallocated_outputs(hsl, orc):
    """hsl_pinned_alloc / `out=`: the per-frame entries write straight into caller-owned (page-locked) arrays; same bits
    as the default pageable path; wrong shapes are refused before the call."""
    p, name = orc.load_preset(PRESETS, 8)
    m = hsl.Model(model_xml(name))
    n_t = 23
    batch = np.tile(p, (3, 1))
    batch[:, 7] = (2.5, 3.0, 4.0)
    ref = m.eval_gaits_detail(batch, n_t)
    pin = {k: hsl.pinned_empty(ref[k].shape, ref[k].dtype) for k in ("traj", "x", "z", "tau", "contacts", "cot")}
    got = m.eval_gaits_detail(batch, n_t, out=pin)
    for k in pin:
        assert got[k] is pin[k] and np.array_equal(got[k], ref[k]), k
    traj = hsl.pinned_empty((3, n_t + 5, m.config_dim))
    traj[:, :n_t + 4] = ref["traj"]; traj[:, -1] = ref["traj"][:, -1]
    t = m.eval_trajectories(traj, batch[:, 7] / n_t, n_t, out={"x": pin["x"]})
    assert t["x"] is pin["x"] and rel_err(t["x"], ref["x"]) < TOL and rel_err(t["tau"], ref["tau"]) < TOL
    with pytest.raises(ValueError):
        m.eval_gaits_detail(batch, n_t, out={"x": np.empty((3, n_t, 5))})
    with pytest.raises(ValueError):
        m.eval_gaits_detail(batch, n_t, out={"z": np.empty((3, n_t, 3 * m.nf), np.float32)})


def test_chunked_batches_equal_single_launch(hsl, orc):
    """hsl_set_max_slots: a batch evaluated as several consecutive launches (bounded workspace; also what keeps the
    kernels' 32-bit slot indices exact for huge batches) gives bit-identical costs, statistics and status, including
    a failed candidate in a later chunk and a last chunk that is not full."""
    xml = model_xml("hexapod")
    p, _ = orc.load_preset(PRESETS, 8)
    batch = np.tile(p, (45, 1))
    batch[:, 7] = np.linspace(2.0, 6.0, 45)
    batch[:, 9] = np.linspace(0.02, 0.12, 45)
    batch[31, 2] = 5.0            # torso far above the ground: unreachable (lik.cpp:142-146)
    n_t = 20
    m = hsl.Model(xml)
    one = m.eval_gaits(batch, n_t)
    launches = m.launch_count()
    assert one["status"][31] != 0 and (np.delete(one["status"], 31) == 0).all()
    for max_slots in (24 * 7, 24 * 44, 24 + 5):   # 7, 2 and 45 chunks (7 / 44 / 1 candidates each)
        m.set_max_slots(max_slots)
        before = m.launch_count()
        got = m.eval_gaits(batch, n_t)
        per = max_slots // (n_t + 4)
        assert m.launch_count() - before == 3 * -(-45 // per)
        for key in ("cot", "work", "min_cfz", "max_mu", "status"):
            assert np.array_equal(got[key], one[key], equal_nan=(key != "status")), (max_slots, key)
    assert launches == 3
    m.set_max_slots(1 << 26)
    ref = orc.Model(xml).eval_batch(batch, n_t, nthreads=8)
    ok = ref["status"] == 0
    assert np.abs(one["cot"][ok] - ref["cot"][ok]).max() <= TOL * np.abs(ref["cot"][ok]).max()


@pytest.mark.parametrize("pid,transl,eas", [(8, (0, 0, 0), (0, 0, -1.571)), (1, (0, 0, 0), (0, 0, -1.571)), (12, (0.3, -0.2, 0), (0, 0, 0.7)),
                                            (8, (0, 0, -0.05), (0, 0, 0.3)), (24, (0, 0, 0), (0.02, 0.03, 0)), (9, (0.1, 0.1, 0), (0, 0, 2.5))])
def test_rec_transform(hsl, orc, pid, transl, eas):
    """pergensetup::rec_transform (pergen.cpp:309-335) through hsl_set_rec_transform: detail entry against the oracle,
    cost-only kernels (plain and pipelined) against the detail entry, and switching it off restores the plain result."""
    params, name = orc.load_preset(PRESETS, pid)
    xml = model_xml(name)
    n_t = 20
    ref = orc.Model(xml).measure_cot(params, n_t, detail=True, rec_transform=(transl, eas))
    m = hsl.Model(xml)
    plain = m.eval_gaits(params, n_t)["cot"][0]
    m.set_rec_transform(transl, eas)
    gpu = m.eval_gaits_detail(params, n_t)
    assert ref["status"] == 0 and gpu["status"][0] == 0
    assert _angle_err(gpu["traj"][0], ref["traj"][:n_t + 4]) < 1e-12
    for key in ("x", "z", "tau"):
        assert rel_err(gpu[key][0], ref[key]) < TOL, key
    assert abs(gpu["cot"][0] - ref["cot"]) <= TOL * abs(ref["cot"])
    batch = np.tile(np.asarray(params, np.float64), (70, 1))
    for fb, mr in ((64, 128), (32, 128), (64, 1)):
        m.set_tuning(fb, mr)
        c = m.eval_gaits(batch, n_t)["cot"]
        assert np.abs(c - ref["cot"]).max() <= TOL * abs(ref["cot"]), (fb, mr)
    assert abs(plain - ref["cot"]) > 1e-6 * abs(ref["cot"])
    m.set_rec_transform()
    assert abs(m.eval_gaits(params, n_t)["cot"][0] - plain) <= 1e-12 * abs(plain)


@pytest.mark.parametrize("pid", [8, 1, 24, 12, 9, 26])
def test_forces_from_torques(hsl, orc, pid):
    """forcetorquesolver::solve_forces / periodic::solve_contforces_given_torques (ftsolver.cpp:331-378,
    periodic.cpp:369-374) through the C ABI, gait and dynrecord entries, consistent and arbitrary torques."""
    params, name = orc.load_preset(PRESETS, pid)
    xml = model_xml(name)
    n_t = 37
    om = orc.Model(xml)
    m = hsl.Model(xml)
    fwd = om.measure_cot(params, n_t, detail=True)
    fields = om.frame_fields(params, n_t)
    rng = np.random.default_rng(pid)
    for tau in (fwd["tau"], fwd["tau"] + rng.normal(0, 1.0, fwd["tau"].shape)):
        ref = om.solve_forces_frames(params, n_t, tau)
        got = m.solve_forces_gait(params, n_t, tau)
        assert got["status"][0] == 0
        assert rel_err(got["z"][0], ref) < TOL
        got2 = m.solve_forces(fields["pos"], fields["jpos"], fields["jzaxis"], fields["mom_rate"], fields["ang_mom_rate"], fields["fpos"], tau)
        assert (got2["status"] == 0).all()
        assert rel_err(got2["z"], ref) < TOL
    # round trip on the device alone (test_dynamics): forward solve -> torques -> contact forces, frames with >= 3 contacts
    det = m.eval_gaits_detail(params, n_t)
    back = m.solve_forces_gait(params, n_t, det["tau"][0])["z"][0]
    sel = det["contacts"][0].sum(axis=1) >= 3
    if sel.any():
        assert np.abs(back[sel] - det["z"][0][sel]).max() < 1e-9 * np.abs(det["z"]).max()
    # a batch of two candidates with different torques keeps them apart
    p2 = np.stack([params, params]); t2 = np.stack([fwd["tau"], fwd["tau"] * 0.5])
    zb = m.solve_forces_gait(p2, n_t, t2)["z"]
    assert rel_err(zb[0], om.solve_forces_frames(params, n_t, t2[0])) < TOL
    assert rel_err(zb[1], om.solve_forces_frames(params, n_t, t2[1])) < TOL


@pytest.mark.parametrize("pid", [8, 9, 12, 24, 20])
def test_record_level_entries(hsl, orc, pid):
    """pergensetup::set_rec and kinematicmodel::set_jvalues_with_lik through the C ABI (hsl_gait_records_host,
    hsl_ik_records_host) against the oracle: records at arbitrary times, joint values of arbitrary records."""
    params, name = orc.load_preset(PRESETS, pid)
    xml = model_xml(name)
    om = orc.Model(xml)
    m = hsl.Model(xml)
    times = np.array([0.0, 0.137, 1.0, 2.75, 7.3, 19.9])
    p2 = np.stack([params, params]); p2[1, 8] *= 0.5
    got = m.gait_records(p2, times)
    assert (got["status"] == 0).all()
    for c in range(2):
        for k, t in enumerate(times):
            assert np.abs(got["rec"][c, k] - om.gait_rec(p2[c], float(t))).max() < 1e-12, (c, k)
    recs = got["rec"].reshape(-1, got["rec"].shape[-1])
    far = recs[0].copy(); far[6:9] += 5.0
    ik = m.ik_records(np.vstack([recs, far[None]]))
    for k in range(recs.shape[0]):
        rc, q = om.ik(recs[k])
        assert (ik["status"][k] == 2) == (rc != 0)
        if rc == 0:
            d = np.abs(ik["q"][k] - q); d[6:] = np.minimum(d[6:], np.abs(d[6:] - 2 * np.pi))
            assert d.max() < 1e-12
    assert ik["status"][-1] == 2
    assert m.ik_records(far, flags=hsl.HSL_FLAG_IGNORE_REACH)["status"][0] == 0
