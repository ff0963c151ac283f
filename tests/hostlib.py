"""ctypes access to tests/hostcheck (the CPU emulation of the kernel math; test infrastructure)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "hostcheck", "_build", "libhslhost.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-C", os.path.join(HERE, "hostcheck"), "-s"])
        _lib = C.CDLL(LIB)
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def model_dims(xml):
    pod = np.zeros(16384, np.uint8)
    rc = lib().hc_model_pod(xml.encode(), _p(pod))
    assert rc == 0, rc
    d = pod[:24].view(np.int32)
    return dict(n=int(d[0]), nf=int(d[1]), nmj=int(d[2]), ntrunk=int(d[3]), config_dim=int(d[4]), lik_index=int(d[5]))


def eval_gaits(xml, params, n_t, flags=0):
    params = np.ascontiguousarray(params, np.float64).reshape(-1, 13)
    c = params.shape[0]
    d = model_dims(xml)
    out = dict(cot=np.zeros(c), work=np.zeros(c), min_cfz=np.zeros(c), max_mu=np.zeros(c), status=np.zeros(c, np.int32),
               traj=np.zeros((c, n_t + 4, d["config_dim"])), x=np.zeros((c, n_t, 6 * d["n"])), z=np.zeros((c, n_t, 3 * d["nf"])),
               tau=np.zeros((c, n_t, d["nmj"])), contacts=np.zeros((c, n_t, d["nf"]), np.uint8))
    rc = lib().hc_eval_gaits(xml.encode(), C.c_int64(c), C.c_int(n_t), _p(params), C.c_int(flags),
                             *[_p(out[k]) for k in ("cot", "work", "min_cfz", "max_mu", "status", "traj", "x", "z", "tau", "contacts")])
    assert rc == 0, rc
    return out


def set_rec_transform(transl=None, eas=None):
    """Process-wide mirror of hsl_set_rec_transform for the emulation; call with no arguments to switch it off."""
    tr = None if transl is None else np.ascontiguousarray(transl, np.float64)
    ea = None if eas is None else np.ascontiguousarray(eas, np.float64)
    lib().hc_set_rec_transform(_p(tr), _p(ea))


def set_axis_specialisation(on=True):
    """Emulate the kernels specialised for the model's hinge-axis pattern (default) or the generic ones."""
    lib().hc_set_axis_specialisation(C.c_int(1 if on else 0))


def axis_pattern(xml):
    return lib().hc_axis_pattern(xml.encode())


def eval_trajectories(xml, traj, dt, n_t):
    d = model_dims(xml)
    traj = np.ascontiguousarray(traj, np.float64).reshape(-1, n_t + 5, d["config_dim"])
    c = traj.shape[0]
    dt = np.ascontiguousarray(np.broadcast_to(np.asarray(dt, np.float64), (c,)))
    out = dict(work=np.zeros(c), min_cfz=np.zeros(c), max_mu=np.zeros(c), status=np.zeros(c, np.int32),
               x=np.zeros((c, n_t, 6 * d["n"])), z=np.zeros((c, n_t, 3 * d["nf"])), tau=np.zeros((c, n_t, d["nmj"])))
    rc = lib().hc_eval_trajectories(xml.encode(), C.c_int64(c), C.c_int(n_t), _p(traj), _p(dt), _p(out["work"]), _p(out["min_cfz"]),
                                    _p(out["max_mu"]), _p(out["status"]), _p(out["x"]), _p(out["z"]), _p(out["tau"]))
    assert rc == 0, rc
    return out


def solve_frames(xml, f):
    d = model_dims(xml)
    nfr = f["pos"].shape[0]
    out = dict(x=np.zeros((nfr, 6 * d["n"])), z=np.zeros((nfr, 3 * d["nf"])), tau=np.zeros((nfr, d["nmj"])), status=np.zeros(nfr, np.int32))
    arrs = [np.ascontiguousarray(f[k], np.float64) for k in ("pos", "jpos", "jzaxis", "mom_rate", "ang_mom_rate", "fpos")]
    con = np.ascontiguousarray(f["contacts"], np.uint8)
    rc = lib().hc_solve_frames(xml.encode(), C.c_int64(nfr), *[_p(a) for a in arrs], _p(con), _p(out["x"]), _p(out["z"]), _p(out["tau"]),
                               _p(out["status"]))
    assert rc == 0, rc
    return out


def solve_forces_gait(xml, params, n_t, tau, flags=0):
    """Emulation of the forces-from-torques kernel along generated gaits: tau [C][n_t][nmj] -> z [C][n_t][3nf]."""
    params = np.ascontiguousarray(params, np.float64).reshape(-1, 13)
    c = params.shape[0]
    d = model_dims(xml)
    tau = np.ascontiguousarray(tau, np.float64).reshape(c, n_t, d["nmj"])
    z = np.zeros((c, n_t, 3 * d["nf"])); status = np.zeros(c, np.int32)
    rc = lib().hc_solve_forces_gait(xml.encode(), C.c_int64(c), C.c_int(n_t), _p(params), C.c_int(flags), _p(tau), _p(z), _p(status))
    assert rc == 0, rc
    return dict(z=z, status=status)


def solve_forces_fields(xml, f, tau):
    d = model_dims(xml)
    nfr = f["pos"].shape[0]
    arrs = [np.ascontiguousarray(f[k], np.float64) for k in ("pos", "jpos", "jzaxis", "mom_rate", "ang_mom_rate", "fpos")]
    tau = np.ascontiguousarray(tau, np.float64).reshape(nfr, d["nmj"])
    z = np.zeros((nfr, 3 * d["nf"])); status = np.zeros(nfr, np.int32)
    rc = lib().hc_solve_forces_fields(xml.encode(), C.c_int64(nfr), *[_p(a) for a in arrs], _p(tau), _p(z), _p(status))
    assert rc == 0, rc
    return dict(z=z, status=status)


def gait_records(xml, params, times, flags=0):
    params = np.ascontiguousarray(params, np.float64).reshape(-1, 13)
    times = np.ascontiguousarray(times, np.float64).reshape(-1)
    d = model_dims(xml)
    c = params.shape[0]
    rec = np.zeros((c, times.shape[0], 6 + 3 * d["nf"])); status = np.zeros(c, np.int32)
    rc = lib().hc_gait_records(xml.encode(), C.c_int64(c), _p(params), C.c_int(times.shape[0]), _p(times), C.c_int(flags), _p(rec), _p(status))
    assert rc == 0, rc
    return dict(rec=rec, status=status)


def ik_records(xml, rec, flags=0):
    d = model_dims(xml)
    rec = np.ascontiguousarray(rec, np.float64).reshape(-1, 6 + 3 * d["nf"])
    n = rec.shape[0]
    q = np.zeros((n, d["config_dim"])); status = np.zeros(n, np.int32)
    rc = lib().hc_ik_records(xml.encode(), C.c_int64(n), _p(rec), C.c_int(flags), _p(q), _p(status))
    assert rc == 0, rc
    return dict(q=q, status=status)


def eval_gaits_pipe(xml, params, n_t, flags=0, fb=64, grid=3):
    """Serial emulation of the persistent pipelined cost-only kernel."""
    params = np.ascontiguousarray(params, np.float64).reshape(-1, 13)
    c = params.shape[0]
    out = dict(cot=np.zeros(c), work=np.zeros(c), min_cfz=np.zeros(c), max_mu=np.zeros(c), status=np.zeros(c, np.int32))
    rc = lib().hc_eval_gaits_pipe(xml.encode(), C.c_int64(c), C.c_int(n_t), _p(params), C.c_int(flags), C.c_int(fb), C.c_int(grid),
                                  *[_p(out[k]) for k in ("cot", "work", "min_cfz", "max_mu", "status")])
    assert rc == 0, rc
    return out


def fall_sweep(xml, params, n_steps, kick_step, kick_dv, play_dt=0.02, t0=0.0, hc=0.7, tmin=0.1, want_traj=False):
    """hsl_fall_sweep_host with the sweep kernel's per-world code run serially on the host."""
    params = np.ascontiguousarray(params, np.float64)
    ks = np.ascontiguousarray(kick_step, np.int32)
    kv = np.ascontiguousarray(kick_dv, np.float64).reshape(-1, 3)
    w = ks.shape[0]
    out = dict(fell=np.zeros(w, np.uint8), t_end=np.zeros(w), final_z=np.zeros(w), status=np.zeros(w, np.int32))
    traj = np.zeros((w, n_steps, 3)) if want_traj else None
    rc = lib().hc_fall_sweep(xml.encode(), C.c_int64(w), _p(params), C.c_double(play_dt), C.c_double(t0), C.c_int(n_steps), _p(ks), _p(kv),
                             C.c_double(hc), C.c_double(tmin), _p(out["fell"]), _p(out["t_end"]), _p(out["final_z"]), _p(out["status"]), _p(traj))
    assert rc == 0, rc
    if want_traj:
        out["traj"] = traj
    return out
