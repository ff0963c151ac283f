"""GPU tier: the device-pointer forms of the per-frame entries (hsl_solve_frames, hsl_eval_trajectories: SURVEY.md 8b L1 / L2
with a stream) and the FK record entry (hsl_fk_records_host), against the oracle and the host-buffer forms."""
import numpy as np
import pytest

from conftest import PRESETS, model_xml, ref_xml, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.mark.parametrize("pid", [8, 9, 24])
def test_frame_solve_on_device_pointers(hsl, orc, pid):
    import torch
    params, name = orc.load_preset(PRESETS, pid)
    om = orc.Model(model_xml(name))
    n_t = 24
    f = om.frame_fields(params, n_t)
    ref = om.measure_cot(params, n_t, detail=True)
    m = hsl.Model(model_xml(name))
    dev = torch.device("cuda")
    d = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in f.items()}
    x = torch.empty((n_t, 6 * m.n), dtype=torch.float64, device=dev)
    z = torch.empty((n_t, 3 * m.nf), dtype=torch.float64, device=dev)
    tau = torch.empty((n_t, m.nmj), dtype=torch.float64, device=dev)
    st = torch.full((n_t,), 7, dtype=torch.int32, device=dev)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    m.solve_frames_device(n_t, d["pos"].data_ptr(), d["jpos"].data_ptr(), d["jzaxis"].data_ptr(), d["mom_rate"].data_ptr(),
                          d["ang_mom_rate"].data_ptr(), d["fpos"].data_ptr(), d["contacts"].data_ptr(), x.data_ptr(), z.data_ptr(),
                          tau.data_ptr(), st.data_ptr(), s.cuda_stream)
    s.synchronize()
    assert (st.cpu().numpy() == 0).all()
    assert rel_err(x.cpu().numpy(), ref["x"]) < TOL and rel_err(z.cpu().numpy(), ref["z"]) < TOL and rel_err(tau.cpu().numpy(), ref["tau"]) < TOL
    host = m.solve_frames(f["pos"], f["jpos"], f["jzaxis"], f["mom_rate"], f["ang_mom_rate"], f["fpos"], f["contacts"])
    assert np.array_equal(host["x"], x.cpu().numpy()) and np.array_equal(host["tau"], tau.cpu().numpy())


@pytest.mark.parametrize("pid", [8, 1])
def test_trajectory_entry_on_device_pointers(hsl, orc, pid):
    import torch
    params, name = orc.load_preset(PRESETS, pid)
    om = orc.Model(model_xml(name))
    n_t, c = 20, 3
    ps = np.tile(params, (c, 1)); ps[1, 8] *= 0.8; ps[2, 7] *= 1.3
    refs = [om.measure_cot(p, n_t, detail=True) for p in ps]
    traj = np.stack([r["traj"] for r in refs])
    dt = ps[:, 7] / n_t
    m = hsl.Model(model_xml(name))
    dev = torch.device("cuda")
    d_traj, d_dt = torch.from_numpy(traj).to(dev), torch.from_numpy(dt).to(dev)
    work = torch.empty(c, dtype=torch.float64, device=dev); mn = torch.empty_like(work); mx = torch.empty_like(work)
    st = torch.empty(c, dtype=torch.int32, device=dev)
    x = torch.empty((c, n_t, 6 * m.n), dtype=torch.float64, device=dev)
    z = torch.empty((c, n_t, 3 * m.nf), dtype=torch.float64, device=dev)
    tau = torch.empty((c, n_t, m.nmj), dtype=torch.float64, device=dev)
    m.eval_trajectories_device(c, n_t, d_traj.data_ptr(), d_dt.data_ptr(), work.data_ptr(), mn.data_ptr(), mx.data_ptr(), st.data_ptr(),
                               x.data_ptr(), z.data_ptr(), tau.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert (st.cpu().numpy() == 0).all()
    for i, r in enumerate(refs):
        assert rel_err(x[i].cpu().numpy(), r["x"]) < TOL and rel_err(z[i].cpu().numpy(), r["z"]) < TOL and rel_err(tau[i].cpu().numpy(), r["tau"]) < TOL
        assert abs(work[i].item() - r["work"]) <= TOL * abs(r["work"])
    # costs only: no per-frame dump requested
    work2 = torch.empty(c, dtype=torch.float64, device=dev)
    m.eval_trajectories_device(c, n_t, d_traj.data_ptr(), d_dt.data_ptr(), work2.data_ptr(), 0, 0, 0, 0, 0, 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert torch.equal(work, work2)


@pytest.mark.parametrize("name", ["myant", "hexapod", "spider"])
def test_fk_records(hsl, orc, refb, name):
    """kinematicmodel::set_jvalues + recompute_modelnodes: body and joint frames against the oracle and the reference build."""
    om, rm = orc.Model(model_xml(name)), refb.Model(ref_xml(name))
    m = hsl.Model(model_xml(name))
    rng = np.random.default_rng(4)
    q = rng.uniform(-1.2, 1.2, (7, m.config_dim))
    got = m.fk_records(q)
    jointed = om.constants()["jkind"] != 0
    for i in range(q.shape[0]):
        for src in (om, rm):
            A, J = src.fk(q[i])
            assert np.abs(got["A_ground"][i] - A).max() < 1e-13
            assert np.abs(got["J_A_ground"][i][jointed] - J[jointed]).max() < 1e-13
        assert not got["J_A_ground"][i][~jointed].any()   # (spider has no jointless body)
    t = m.tables()
    cons = om.constants()
    assert np.array_equal(t["parent"], cons["parent"]) and np.array_equal(t["footis"], cons["limb_foot"]) and np.array_equal(t["limb_top"], cons["limb_top"])
    assert (t["masses"] == 1).all()
    assert np.array_equal(t["com_offset"], cons["A_body_geom"][:, 12:15])
    assert np.array_equal(t["foot_offset"], cons["capsule_to_pos"][cons["limb_foot"]])


def test_gather_object_with_one_rank(hsl):
    """hsl_eval_gaits_gather on a job of one rank (no peer to map): the gather buffer holds this rank's costs and status,
    NaN / 0 behind them, the two buffers alternate, and an empty shard leaves only padding."""
    import torch
    from hslabs_b200 import api
    from hslabs_b200.search import _DeviceArray
    rng = np.random.default_rng(5)
    n, per, n_t = 37, 50, 24
    p = np.zeros((n, 13))
    p[:, 2] = rng.uniform(-0.15, -0.05, n); p[:, 6] = rng.uniform(0, 1, n); p[:, 7] = rng.uniform(1, 6, n)
    p[:, 8] = rng.uniform(0.1, 0.5, n); p[:, 9] = rng.uniform(0.02, 0.12, n); p[:, 11] = -1
    p[5, 2] = 0.4   # unreachable
    m = hsl.Model(hsl.model_path("hexapod"))
    want = m.eval_gaits(p, n_t)
    g = api.Gather(0, 1, per, lambda h: [h])
    d = torch.from_numpy(p).cuda()
    cot = torch.empty(n, dtype=torch.float64, device="cuda")
    ptrs = []
    for k in (n, 0, 11):
        pc, ps = m.eval_gaits_gather(g, k, n_t, d.data_ptr() if k else 0, d_cot=cot.data_ptr() if k else 0, stream=torch.cuda.current_stream().cuda_stream)
        ptrs.append(pc)
        allc = torch.as_tensor(_DeviceArray(pc, per, "<f8"), device="cuda").cpu().numpy()
        alls = torch.as_tensor(_DeviceArray(ps, per, "<i4"), device="cuda").cpu().numpy()
        assert np.array_equal(allc[:k], want["cot"][:k], equal_nan=True) and np.isnan(allc[k:]).all()
        assert np.array_equal(alls[:k], want["status"][:k]) and (alls[k:] == 0).all()
        if k:
            assert np.array_equal(cot.cpu().numpy()[:k], want["cot"][:k], equal_nan=True)
    assert ptrs[0] != ptrs[1] and ptrs[0] == ptrs[2]
    best = torch.empty(1, dtype=torch.int64, device="cuda")
    bval = torch.empty(1, dtype=torch.float64, device="cuda")
    m.eval_gaits_scatter(g, n, n_t, d.data_ptr())
    g.select_best(best.data_ptr(), bval.data_ptr())
    key = np.where(np.isnan(want["cot"]), np.inf, want["cot"])
    assert int(best.item()) == int(np.argmin(key)) and float(bval.item()) == float(key.min())
    with pytest.raises(hsl.HslError):
        m.eval_gaits_gather(g, per + 1, n_t, d.data_ptr())
    g.free()


def test_back_to_back_calls_on_one_stream(hsl):
    """The kernels of a call are chained by programmatic dependent launch (the next kernel's blocks become resident before the
    previous one has drained), and consecutive calls reuse the handle's workspace: 60 calls alternating between two batches of
    different size, queued without any synchronisation, must each give exactly what the batch gives on its own."""
    import torch
    rng = np.random.default_rng(17)

    def batch(n):
        p = np.zeros((n, 13))
        p[:, 2] = rng.uniform(-0.15, -0.05, n); p[:, 6] = rng.uniform(0, 1, n); p[:, 7] = rng.uniform(1, 6, n)
        p[:, 8] = rng.uniform(0.1, 0.5, n); p[:, 9] = rng.uniform(0.02, 0.12, n); p[:, 11] = -1
        return p
    m = hsl.Model(hsl.model_path("hexapod"))
    pa, pb = batch(777), batch(1501)
    want = [m.eval_gaits(pa, 40), m.eval_gaits(pb, 24)]
    da, db = torch.from_numpy(pa).cuda(), torch.from_numpy(pb).cuda()
    outs = []
    s = torch.cuda.current_stream().cuda_stream
    best = torch.empty(60, dtype=torch.int64, device="cuda")
    from hslabs_b200 import api
    for it in range(60):
        which = it % 2
        d, n, n_t = (da, 777, 40) if which == 0 else (db, 1501, 24)
        cot = torch.empty(n, dtype=torch.float64, device="cuda")
        st = torch.empty(n, dtype=torch.int32, device="cuda")
        m.eval_gaits_device(n, n_t, d.data_ptr(), cot.data_ptr(), 0, 0, 0, st.data_ptr(), 0, s)
        api.select_best_device(cot.data_ptr(), n, best[it:].data_ptr(), 0, s)
        outs.append((which, cot, st))
    torch.cuda.synchronize()
    for it, (which, cot, st) in enumerate(outs):
        w = want[which]
        assert np.array_equal(cot.cpu().numpy(), w["cot"], equal_nan=True), it
        assert np.array_equal(st.cpu().numpy(), w["status"]), it
        key = np.where(np.isnan(w["cot"]), np.inf, w["cot"])
        assert int(best[it]) == int(np.argmin(key)), it
