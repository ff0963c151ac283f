"""ctypes binding of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; nothing under hslabs_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "liborc.so")
LIBQ = os.path.join(HERE, "_build", "liborcq.so")
NPARAM = 13
_lib = None


def build(force=False):
    srcs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cpp", ".hpp"))]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs
                                                 if os.access(HERE, os.W_OK)):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return LIB


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            build()
        _lib = C.CDLL(LIB)
        _lib.orc_model_load.restype = C.c_void_p
        _lib.orc_model_load.argtypes = [C.c_char_p]
        _lib.orc_model_rcap.restype = C.c_double
        for name in ("orc_model_free", "orc_model_dims", "orc_model_rcap", "orc_set_ignore_reach", "orc_model_constants",
                     "orc_fk", "orc_ik", "orc_gait_setup", "orc_gait_rec", "orc_measure_cot", "orc_measure_cot_rect", "orc_frame_fields",
                     "orc_eval_trajectory", "orc_test_dynamics", "orc_solve_forces_frames", "orc_measure_cot_sweep", "orc_eval_batch"):
            getattr(_lib, name).argtypes = None
    return _lib


_libq = None


def libq():
    """Quad-precision build of the same oracle headers (orc_capi_quad.cpp); conditioning yardstick only."""
    global _libq
    if _libq is None:
        if not os.path.exists(LIBQ):
            build(force=True)
        _libq = C.CDLL(LIBQ)
        _libq.orcq_model_load.restype = C.c_void_p
        _libq.orcq_model_load.argtypes = [C.c_char_p]
        _libq.orcq_model_free.argtypes = [C.c_void_p]
    return _libq


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def make_params(torso_pos=(0, 0, 0), torso_angles=(0, 0, 0), step_duration=1.0, period=3.0, step_length=0.5,
                step_height=0.1, curvature=0.0, shift_type=-1, shift_value=0.0):
    return np.array(list(torso_pos) + list(torso_angles) + [step_duration, period, step_length, step_height,
                                                             curvature, shift_type, shift_value], dtype=np.float64)


def load_preset(path, pid):
    params = np.zeros(NPARAM)
    name = C.create_string_buffer(64)
    rc = lib().orc_load_preset(path.encode(), C.c_int(pid), _p(params), name)
    if rc != 0:
        raise KeyError("preset %d not found in %s" % (pid, path))
    return params, name.value.decode()


class Model:
    def __init__(self, xml_path):
        self.path = xml_path
        self.h = C.c_void_p(lib().orc_model_load(xml_path.encode()))
        if not self.h:
            raise RuntimeError("oracle could not load " + xml_path)
        d = (C.c_int * 4)()
        lib().orc_model_dims(self.h, d)
        self.n, self.nf, self.nmj, self.config_dim = d[0], d[1], d[2], d[3]
        self.rcap = lib().orc_model_rcap(self.h)

    def __del__(self):
        try:
            if self.h:
                lib().orc_model_free(self.h)
                self.h = None
        except Exception:
            pass

    def set_ignore_reach(self, flag):
        lib().orc_set_ignore_reach(self.h, C.c_int(int(flag)))

    def constants(self):
        n, nf = self.n, self.nf
        out = dict(parent=np.zeros(n, np.int32), jkind=np.zeros(n, np.int32), A_pj_body=np.zeros((n, 16)),
                   J_A_parent=np.zeros((n, 16)), A_body_geom=np.zeros((n, 16)), capsule_to_pos=np.zeros((n, 3)),
                   limb_top=np.zeros(nf, np.int32), limb_foot=np.zeros(nf, np.int32))
        lib().orc_model_constants(self.h, _p(out["parent"]), _p(out["jkind"]), _p(out["A_pj_body"]), _p(out["J_A_parent"]),
                                  _p(out["A_body_geom"]), _p(out["capsule_to_pos"]), _p(out["limb_top"]), _p(out["limb_foot"]))
        return out

    def fk(self, q):
        q = np.ascontiguousarray(q, np.float64)
        A = np.zeros((self.n, 16)); J = np.zeros((self.n, 16))
        lib().orc_fk(self.h, _p(q), _p(A), _p(J))
        return A, J

    def ik(self, rec):
        rec = np.ascontiguousarray(rec, np.float64)
        q = np.zeros(self.config_dim)
        rc = lib().orc_ik(self.h, _p(rec), _p(q))
        return rc, q

    def gait_setup(self, params):
        params = np.ascontiguousarray(params, np.float64)
        pos0 = np.zeros((self.nf, 3)); ts = np.zeros(self.nf); xs = np.zeros(self.nf); scal = np.zeros(3)
        rc = lib().orc_gait_setup(self.h, _p(params), _p(pos0), _p(ts), _p(xs), _p(scal))
        if rc:
            raise ValueError("bad gait parameters")
        return pos0, ts, xs, scal

    def gait_rec(self, params, t):
        params = np.ascontiguousarray(params, np.float64)
        rec = np.zeros(6 + 3 * self.nf)
        rc = lib().orc_gait_rec(self.h, _p(params), C.c_double(t), _p(rec))
        if rc:
            raise ValueError("bad gait parameters")
        return rec

    def measure_cot(self, params, n_t, detail=False, rec_transform=None):
        """rec_transform = (transl[3], eas[3]) switches pergensetup::rec_transform on (pergen.cpp:309-335)."""
        params = np.ascontiguousarray(params, np.float64)
        out = np.zeros(4)
        traj = x = z = tau = None
        if detail:
            traj = np.zeros((n_t + 5, self.config_dim)); x = np.zeros((n_t, 6 * self.n))
            z = np.zeros((n_t, 3 * self.nf)); tau = np.zeros((n_t, self.nmj))
        if rec_transform is not None:
            tr = np.ascontiguousarray(rec_transform[0], np.float64); ea = np.ascontiguousarray(rec_transform[1], np.float64)
            rc = lib().orc_measure_cot_rect(self.h, _p(params), _p(tr), _p(ea), C.c_int(n_t), _p(out), _p(traj), _p(x), _p(z), _p(tau))
        else:
            rc = lib().orc_measure_cot(self.h, _p(params), C.c_int(n_t), _p(out), _p(traj), _p(x), _p(z), _p(tau))
        res = dict(status=rc, cot=out[0], work=out[1], min_cfz=out[2], max_mu=out[3])
        if detail:
            res.update(traj=traj, x=x, z=z, tau=tau)
        return res

    def measure_cot_quad(self, params, n_t):
        """measure_cot evaluated in __float128 (results rounded to double); always returns the detail arrays."""
        params = np.ascontiguousarray(params, np.float64)
        hq = C.c_void_p(libq().orcq_model_load(self.path.encode()))
        if not hq:
            raise RuntimeError("quad oracle could not load %s" % self.path)
        out = np.zeros(4)
        traj = np.zeros((n_t + 5, self.config_dim)); x = np.zeros((n_t, 6 * self.n))
        z = np.zeros((n_t, 3 * self.nf)); tau = np.zeros((n_t, self.nmj))
        try:
            rc = libq().orcq_measure_cot(hq, _p(params), C.c_int(n_t), _p(out), _p(traj), _p(x), _p(z), _p(tau))
        finally:
            libq().orcq_model_free(hq)
        return dict(status=rc, cot=out[0], work=out[1], min_cfz=out[2], max_mu=out[3], traj=traj, x=x, z=z, tau=tau)

    def frame_fields(self, params, n_t):
        params = np.ascontiguousarray(params, np.float64)
        n, nf = self.n, self.nf
        f = dict(pos=np.zeros((n_t, n, 3)), jpos=np.zeros((n_t, n, 3)), jzaxis=np.zeros((n_t, n, 3)),
                 mom_rate=np.zeros((n_t, n, 3)), ang_mom_rate=np.zeros((n_t, n, 3)), fpos=np.zeros((n_t, nf, 3)),
                 contacts=np.zeros((n_t, nf), np.uint8))
        rc = lib().orc_frame_fields(self.h, _p(params), C.c_int(n_t), _p(f["pos"]), _p(f["jpos"]), _p(f["jzaxis"]),
                                    _p(f["mom_rate"]), _p(f["ang_mom_rate"]), _p(f["fpos"]), _p(f["contacts"]))
        if rc:
            raise ValueError("frame_fields failed rc=%d" % rc)
        return f

    def eval_trajectory(self, q, n_t, dt):
        q = np.ascontiguousarray(q, np.float64)
        out = np.zeros(3); x = np.zeros((n_t, 6 * self.n)); z = np.zeros((n_t, 3 * self.nf)); tau = np.zeros((n_t, self.nmj))
        rc = lib().orc_eval_trajectory(self.h, _p(q), C.c_int(n_t), C.c_double(dt), _p(out), _p(x), _p(z), _p(tau))
        return dict(status=rc, work=out[0], min_cfz=out[1], max_mu=out[2], x=x, z=z, tau=tau)

    def test_dynamics(self, params, n_t=20, frame=2):
        params = np.ascontiguousarray(params, np.float64)
        cf = np.zeros(3 * self.nf); cf1 = np.zeros(3 * self.nf); tau = np.zeros(self.nmj)
        rc = lib().orc_test_dynamics(self.h, _p(params), C.c_int(n_t), C.c_int(frame), _p(cf), _p(cf1), _p(tau))
        return rc, cf, cf1, tau

    def solve_forces_frames(self, params, n_t, tau):
        """forcetorquesolver::solve_forces on every solved frame: tau [n_t][nmj] -> contact forces [n_t][3nf]."""
        params = np.ascontiguousarray(params, np.float64)
        tau = np.ascontiguousarray(tau, np.float64).reshape(n_t, self.nmj)
        cf = np.zeros((n_t, 3 * self.nf))
        rc = lib().orc_solve_forces_frames(self.h, _p(params), C.c_int(n_t), _p(tau), _p(cf))
        if rc:
            raise ValueError("solve_forces_frames failed: %d" % rc)
        return cf

    def measure_cot_sweep(self, params, n_t, name, v0, v1, n_val):
        params = np.ascontiguousarray(params, np.float64)
        vals = np.zeros(n_val + 1); cots = np.zeros(n_val + 1)
        rc = lib().orc_measure_cot_sweep(self.h, _p(params), C.c_int(n_t), name.encode(), C.c_double(v0), C.c_double(v1),
                                         C.c_int(n_val), _p(vals), _p(cots))
        if rc:
            raise ValueError("sweep failed")
        return vals, cots

    def eval_batch(self, params, n_t, nthreads=1):
        params = np.ascontiguousarray(params, np.float64).reshape(-1, NPARAM)
        c = params.shape[0]
        out = dict(cot=np.zeros(c), work=np.zeros(c), min_cfz=np.zeros(c), max_mu=np.zeros(c), status=np.zeros(c, np.int32))
        lib().orc_eval_batch(self.h, C.c_long(c), C.c_int(n_t), _p(params), _p(out["cot"]), _p(out["work"]),
                             _p(out["min_cfz"]), _p(out["max_mu"]), _p(out["status"]), C.c_int(nthreads))
        return out
