"""ctypes binding of oracle/_ref/libhslref.so: the reference's OWN sources, compiled unmodified against shim
headers (oracle/Makefile target `ref`, oracle/ref_driver.cpp).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__ and bench.py's cpu_baseline / --impl reference legs may import this module; nothing
under hslabs_b200/ does.  The library is built in the container that has /root/reference and travels to the GPU
box as a prebuilt file (git-ignored, not gpurun-ignored); `available()` says whether it is there.

`Model` has the same methods as `oracle.orc.Model`, so a test can run one body against either.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libhslref.so")
REFERENCE_DIR = os.environ.get("HSL_REFERENCE_DIR", "/root/reference")
NPARAM = 13
_lib = None


def build():
    """(Re)build _ref from /root/reference when the sources are present; no-op otherwise."""
    if os.path.exists(os.path.join(REFERENCE_DIR, "periodic.cpp")):
        subprocess.check_call(["make", "-C", HERE, "-s", "ref", "REF=" + REFERENCE_DIR])
    return LIB if os.path.exists(LIB) else None


def available():
    return os.path.exists(LIB)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.ref_model_load.restype = C.c_void_p
        _lib.ref_model_load.argtypes = [C.c_char_p, C.c_char_p]
        _lib.ref_model_rcap.restype = C.c_double
        _lib.ref_model_rcap.argtypes = [C.c_void_p]
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Model:
    """xml_path: path of a model file; its base name must be myant.xml / hexapod.xml / spider.xml (lik.cpp:8-11)."""

    def __init__(self, xml_path):
        self.path = xml_path
        d, name = os.path.split(os.path.abspath(xml_path))
        self.dir, self.xml = d, name
        self.h = C.c_void_p(lib().ref_model_load(d.encode(), name.encode()))
        if not self.h:
            raise RuntimeError("reference build could not load " + xml_path)
        dims = (C.c_int * 4)()
        lib().ref_model_dims(self.h, dims)
        self.n, self.nf, self.nmj, self.config_dim = dims[0], dims[1], dims[2], dims[3]
        self.rcap = lib().ref_model_rcap(self.h)

    def set_ignore_reach(self, flag):
        lib().ref_set_ignore_reach(self.h, C.c_int(int(flag)))

    def constants(self):
        n, nf = self.n, self.nf
        out = dict(parent=np.zeros(n, np.int32), jkind=np.zeros(n, np.int32), A_pj_body=np.zeros((n, 16)),
                   J_A_parent=np.zeros((n, 16)), A_body_geom=np.zeros((n, 16)), capsule_to_pos=np.zeros((n, 3)),
                   limb_top=np.zeros(nf, np.int32), limb_foot=np.zeros(nf, np.int32))
        lib().ref_model_constants(self.h, _p(out["parent"]), _p(out["jkind"]), _p(out["A_pj_body"]), _p(out["J_A_parent"]),
                                  _p(out["A_body_geom"]), _p(out["capsule_to_pos"]), _p(out["limb_top"]), _p(out["limb_foot"]))
        return out

    def fk(self, q):
        q = np.ascontiguousarray(q, np.float64)
        A = np.zeros((self.n, 16)); J = np.zeros((self.n, 16))
        lib().ref_fk(self.h, _p(q), _p(A), _p(J))
        return A, J

    def ik(self, rec):
        rec = np.ascontiguousarray(rec, np.float64)
        q = np.zeros(self.config_dim)
        rc = lib().ref_ik(self.h, _p(rec), _p(q))
        return rc, q

    def gait_setup(self, params):
        params = np.ascontiguousarray(params, np.float64)
        pos0 = np.zeros((self.nf, 3)); ts = np.zeros(self.nf); xs = np.zeros(self.nf); scal = np.zeros(3)
        rc = lib().ref_gait_setup(self.h, _p(params), _p(pos0), _p(ts), _p(xs), _p(scal))
        if rc:
            raise ValueError("bad gait parameters")
        return pos0, ts, xs, scal

    def gait_rec(self, params, t):
        params = np.ascontiguousarray(params, np.float64)
        rec = np.zeros(6 + 3 * self.nf)
        rc = lib().ref_gait_rec(self.h, _p(params), C.c_double(t), _p(rec))
        if rc:
            raise ValueError("bad gait parameters")
        return rec

    def measure_cot(self, params, n_t, detail=False, rec_transform=None):
        params = np.ascontiguousarray(params, np.float64)
        out = np.zeros(4)
        traj = x = z = tau = complete = None
        if detail:
            traj = np.zeros((n_t + 5, self.config_dim)); x = np.zeros((n_t, 6 * self.n))
            z = np.zeros((n_t, 3 * self.nf)); tau = np.zeros((n_t, self.nmj))
            complete = np.zeros((n_t, 2 * self.config_dim + self.nmj))
        if rec_transform is not None:
            tr = np.ascontiguousarray(rec_transform[0], np.float64); ea = np.ascontiguousarray(rec_transform[1], np.float64)
            rc = lib().ref_measure_cot_rect(self.h, _p(params), _p(tr), _p(ea), C.c_int(n_t), _p(out), _p(traj), _p(x), _p(z), _p(tau))
            complete = None
        else:
            rc = lib().ref_measure_cot(self.h, _p(params), C.c_int(n_t), _p(out), _p(traj), _p(x), _p(z), _p(tau), _p(complete))
        if rc:
            out[:] = np.nan
        res = dict(status=rc, cot=out[0], work=out[1], min_cfz=out[2], max_mu=out[3])
        if detail:
            res.update(traj=traj, x=x, z=z, tau=tau, complete=complete)
        return res

    def frame_fields(self, params, n_t):
        params = np.ascontiguousarray(params, np.float64)
        n, nf = self.n, self.nf
        f = dict(pos=np.zeros((n_t, n, 3)), jpos=np.zeros((n_t, n, 3)), jzaxis=np.zeros((n_t, n, 3)),
                 mom_rate=np.zeros((n_t, n, 3)), ang_mom_rate=np.zeros((n_t, n, 3)), fpos=np.zeros((n_t, nf, 3)),
                 contacts=np.zeros((n_t, nf), np.uint8))
        rc = lib().ref_frame_fields(self.h, _p(params), C.c_int(n_t), _p(f["pos"]), _p(f["jpos"]), _p(f["jzaxis"]),
                                    _p(f["mom_rate"]), _p(f["ang_mom_rate"]), _p(f["fpos"]), _p(f["contacts"]))
        if rc:
            raise ValueError("frame_fields failed rc=%d" % rc)
        return f

    def eval_trajectory(self, q, n_t, dt):
        q = np.ascontiguousarray(q, np.float64)
        out = np.zeros(3); x = np.zeros((n_t, 6 * self.n)); z = np.zeros((n_t, 3 * self.nf)); tau = np.zeros((n_t, self.nmj))
        rc = lib().ref_eval_trajectory(self.h, _p(q), C.c_int(n_t), C.c_double(dt), _p(out), _p(x), _p(z), _p(tau))
        return dict(status=rc, work=out[0], min_cfz=out[1], max_mu=out[2], x=x, z=z, tau=tau)

    def test_dynamics(self, params, n_t=20, frame=2):
        params = np.ascontiguousarray(params, np.float64)
        cf = np.zeros(3 * self.nf); cf1 = np.zeros(3 * self.nf); tau = np.zeros(self.nmj)
        rc = lib().ref_test_dynamics(self.h, _p(params), C.c_int(n_t), C.c_int(frame), _p(cf), _p(cf1), _p(tau))
        return rc, cf, cf1, tau

    def solve_forces_frames(self, params, n_t, tau):
        params = np.ascontiguousarray(params, np.float64)
        tau = np.ascontiguousarray(tau, np.float64).reshape(n_t, self.nmj)
        cf = np.zeros((n_t, 3 * self.nf))
        rc = lib().ref_solve_forces_frames(self.h, _p(params), C.c_int(n_t), _p(tau), _p(cf))
        if rc:
            raise ValueError("solve_forces_frames failed: %d" % rc)
        return cf

    def measure_cot_sweep(self, params, n_t, name, v0, v1, n_val):
        params = np.ascontiguousarray(params, np.float64)
        vals = np.zeros(n_val + 1); cots = np.zeros(n_val + 1)
        rc = lib().ref_measure_cot_sweep(self.h, _p(params), C.c_int(n_t), name.encode(), C.c_double(v0), C.c_double(v1),
                                         C.c_int(n_val), _p(vals), _p(cots))
        if rc:
            raise ValueError("sweep failed")
        return vals, cots

    def eval_batch(self, params, n_t, nthreads=1):
        """nthreads = number of forked worker processes (the reference is not thread-safe)."""
        params = np.ascontiguousarray(params, np.float64).reshape(-1, NPARAM)
        c = params.shape[0]
        out = dict(cot=np.zeros(c), work=np.zeros(c), min_cfz=np.zeros(c), max_mu=np.zeros(c), status=np.zeros(c, np.int32))
        lib().ref_eval_batch(self.h, C.c_long(c), C.c_int(n_t), _p(params), _p(out["cot"]), _p(out["work"]),
                             _p(out["min_cfz"]), _p(out["max_mu"]), _p(out["status"]), C.c_int(nthreads))
        return out

    def load_preset(self, path, pid):
        """Preset row parsed by the reference's own get_rec_str / get_pgs_config_params (player.cpp:170-244)."""
        params = np.zeros(NPARAM)
        name = C.create_string_buffer(64)
        rc = lib().ref_load_preset(self.h, os.path.abspath(path).encode(), C.c_int(pid), _p(params), name)
        if rc != 0:
            raise KeyError("preset %d not found in %s" % (pid, path))
        return params, name.value.decode()

    def lik_solver_test(self, n=100):
        return lib().ref_lik_solver_test(self.h, C.c_int(n))

    def fall_run(self, params, n_steps, play_dt=0.02, t0=0.0, kick_step=-1, kick_dv=(0.0, 0.0, 0.0), hc=0.7, tmin=0.1, want_traj=False):
        """The reference's position-control loop on the ODE shim's stepper with one torso kick (oracle/ref_driver.cpp
        ref_fall_run).  Only valid in a process where this is the only model loaded."""
        params = np.ascontiguousarray(params, np.float64)
        dv = np.ascontiguousarray(kick_dv, np.float64)
        out = np.zeros(4)
        traj = np.zeros((n_steps, 3)) if want_traj else None
        lib().ref_fall_run.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_double, C.c_double,
                                       C.c_void_p, C.c_void_p]
        rc = lib().ref_fall_run(self.h, _p(params), play_dt, t0, n_steps, kick_step, _p(dv), hc, tmin, _p(out), _p(traj))
        if rc:
            raise RuntimeError("ref_fall_run failed: %d" % rc)
        res = dict(fell=bool(out[0]), t=out[1], z=out[2], steps=int(out[3]))
        if want_traj:
            res["traj"] = traj[:max(res["steps"], 0)]
        return res

    def fall_batch(self, params, n_steps, kick_step, kick_dv, play_dt=0.02, t0=0.0, hc=0.7, tmin=0.1, nprocs=1):
        """fall_run for W worlds on forked worker processes (one pristine reference process per world)."""
        params = np.ascontiguousarray(params, np.float64)
        ks = np.ascontiguousarray(kick_step, np.int32)
        kv = np.ascontiguousarray(kick_dv, np.float64).reshape(-1, 3)
        w = ks.shape[0]
        fell = np.zeros(w, np.uint8); t_end = np.zeros(w); z = np.zeros(w)
        lib().ref_fall_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_long, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_void_p, C.c_double,
                                         C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        rc = lib().ref_fall_batch(self.h, _p(params), w, play_dt, t0, n_steps, _p(ks), _p(kv), hc, tmin, _p(fell), _p(t_end), _p(z), nprocs)
        if rc:
            raise RuntimeError("ref_fall_batch failed: %d" % rc)
        return dict(fell=fell, t_end=t_end, final_z=z)
