"""ctypes host layer over include/hsl.h.

Names follow the reference's evaluation interface: a candidate is the scalar content of
``pgsconfigparams`` (pergen.h:137-146), ``measure_cot`` / ``measure_cot_sweep`` are
``modelplayer::measure_cot(_sweep)`` (player.cpp:269-285, 311-321) evaluated in one batch on the GPU,
``Model.eval_gaits_detail`` exposes what ``periodic`` keeps per frame (trajectory, joint force/torque
vector x, contact forces z, motor torques; periodic.cpp:77-96, 361-391).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
HSL_NPARAM = 13
HSL_FLAG_IGNORE_REACH = 1
ST_BAD_PARAMS, ST_UNREACHABLE, ST_SOLVER, ST_FEW_CONTACTS, ST_ILLCOND = 1, 2, 4, 8, 16
HSL_ST_BAD_PARAMS, HSL_ST_UNREACHABLE, HSL_ST_SOLVER, HSL_ST_FEW_CONTACTS, HSL_ST_ILLCOND = 1, 2, 4, 8, 16  # include/hsl.h
SWEEP_NAMES = {"step_duration": 6, "period": 7, "step_length": 8, "step_height": 9}  # pergen.cpp:423

_lib = None


class HslError(RuntimeError):
    pass


def lib_path():
    # HSL_B200_LIB selects another build of the same library (e.g. the -DHSL_PHASE_CLOCKS profiling build)
    return os.environ.get("HSL_B200_LIB") or os.path.join(HERE, "lib", "libhsl_b200.so")


def model_path(name):
    """Path of a bundled model input ('hexapod' -> .../models/hexapod.xml)."""
    if not name.endswith(".xml"):
        name += ".xml"
    return os.path.join(HERE, "models", name)


def _load():
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise HslError("libhsl_b200.so is not built (%s); run `python hslabs_b200/build.py` -- there is no CPU fallback" % path)
    lib = C.CDLL(path)
    lib.hsl_last_error.restype = C.c_char_p
    lib.hsl_model_rcap.restype = C.c_double
    lib.hsl_model_pod.restype = C.c_size_t
    lib.hsl_launch_count.restype = C.c_int64
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    lib.hsl_model_load_xml.argtypes = [C.c_char_p, C.POINTER(vp)]
    lib.hsl_model_free.argtypes = [vp]
    lib.hsl_model_dims.argtypes = [vp, vp]
    lib.hsl_model_rcap.argtypes = [vp]
    lib.hsl_model_pod.argtypes = [vp, vp, C.c_size_t]
    lib.hsl_eval_gaits.argtypes = [vp, i64, i32, vp, i32, vp, vp, vp, vp, vp, vp]
    lib.hsl_eval_gaits_host.argtypes = [vp, i64, i32, vp, i32, vp, vp, vp, vp, vp]
    lib.hsl_eval_gaits_detail_host.argtypes = [vp, i64, i32, vp, i32] + [vp] * 10
    lib.hsl_eval_trajectories_host.argtypes = [vp, i64, i32] + [vp] * 9
    lib.hsl_solve_frames_host.argtypes = [vp, i64] + [vp] * 11
    lib.hsl_solve_forces_host.argtypes = [vp, i64] + [vp] * 9
    lib.hsl_gait_records_host.argtypes = [vp, i64, vp, i32, vp, i32, vp, vp]
    lib.hsl_ik_records_host.argtypes = [vp, i64, vp, i32, vp, vp]
    lib.hsl_solve_forces_gait_host.argtypes = [vp, i64, i32, vp, i32, vp, vp, vp]
    lib.hsl_set_tuning.argtypes = [vp, i32, i32]
    lib.hsl_set_max_slots.argtypes = [vp, i64]
    lib.hsl_get_tuning.argtypes = [vp, vp, vp]
    lib.hsl_set_kernel_timing.argtypes = [vp, i32]
    lib.hsl_last_kernel_ms.argtypes = [vp, vp]
    lib.hsl_pinned_alloc.restype = vp
    lib.hsl_pinned_alloc.argtypes = [C.c_size_t]
    lib.hsl_pinned_free.argtypes = [vp]
    lib.hsl_set_rec_transform.argtypes = [vp, vp, vp]
    lib.hsl_launch_count.argtypes = [vp]
    lib.hsl_dfma_probe.argtypes = [i32, i32, i32, vp, vp]
    lib.hsl_math_selftest.argtypes = [i32, vp, vp, vp]
    lib.hsl_select_best.argtypes = [vp, i64, vp, vp, vp]
    lib.hsl_eval_trajectories.argtypes = [vp, i64, i32] + [vp] * 10
    lib.hsl_solve_frames.argtypes = [vp, i64] + [vp] * 12
    lib.hsl_fk_records_host.argtypes = [vp, i64, vp, vp, vp]
    lib.hsl_nccl_unique_id.argtypes = [vp]
    lib.hsl_nccl_comm_init.argtypes = [C.POINTER(vp), i32, vp, i32]
    lib.hsl_nccl_comm_destroy.argtypes = [vp]
    lib.hsl_allgather_costs.argtypes = [vp, vp, i64, vp, vp]
    lib.hsl_allgather_costs_host.argtypes = [vp, i32, vp, i64, vp]
    lib.hsl_model_tables.argtypes = [vp] + [vp] * 6
    lib.hsl_fall_sweep_host.argtypes = [vp, i64, vp, C.c_double, C.c_double, i32, vp, vp, C.c_double, C.c_double] + [vp] * 6
    lib.hsl_select_topk.argtypes = [vp, i64, i32, vp, vp, vp]
    lib.hsl_set_fall_variant.argtypes = [vp, i32]
    lib.hsl_gather_create.argtypes = [i32, i32, i64, vp, vp]
    lib.hsl_gather_connect.argtypes = [vp, vp]
    lib.hsl_gather_free.argtypes = [vp]
    lib.hsl_eval_gaits_gather.argtypes = [vp, vp, i64, i32, vp, i32] + [vp] * 8
    lib.hsl_eval_gaits_scatter.argtypes = [vp, vp, i64, i32, vp, i32] + [vp] * 6
    lib.hsl_gather_wait.argtypes = [vp, vp, vp, vp]
    lib.hsl_gather_select_best.argtypes = [vp, vp, vp, vp]
    lib.hsl_gather_size.argtypes = [vp]
    lib.hsl_gather_check.argtypes = [vp]
    lib.hsl_gather_size.restype = i64
    lib.hsl_eval_gaits_gather_host.argtypes = [vp, vp, i64, i32, vp, i32, vp, vp]
    _lib = lib
    return lib


def exported_symbols():
    """Every entry point include/hsl.h declares (checked by the CPU test tier)."""
    return ["hsl_model_load_xml", "hsl_model_free", "hsl_model_dims", "hsl_model_rcap", "hsl_model_pod", "hsl_last_error",
            "hsl_device_count", "hsl_eval_gaits", "hsl_eval_gaits_host", "hsl_eval_gaits_detail_host",
            "hsl_eval_trajectories_host", "hsl_solve_frames_host", "hsl_gait_records_host", "hsl_ik_records_host", "hsl_solve_forces_host", "hsl_solve_forces_gait_host", "hsl_set_rec_transform", "hsl_set_tuning", "hsl_get_tuning", "hsl_set_kernel_timing", "hsl_last_kernel_ms", "hsl_set_max_slots", "hsl_pinned_alloc", "hsl_pinned_free", "hsl_launch_count",
            "hsl_dfma_probe", "hsl_math_selftest", "hsl_select_best", "hsl_select_topk", "hsl_eval_trajectories", "hsl_solve_frames",
            "hsl_fk_records_host", "hsl_nccl_unique_id", "hsl_nccl_comm_init", "hsl_nccl_comm_destroy", "hsl_allgather_costs", "hsl_allgather_costs_host", "hsl_model_tables", "hsl_fall_sweep_host", "hsl_set_fall_variant",
            "hsl_gather_create", "hsl_gather_connect", "hsl_gather_free", "hsl_eval_gaits_gather", "hsl_eval_gaits_scatter",
            "hsl_gather_wait", "hsl_gather_select_best", "hsl_gather_size", "hsl_eval_gaits_gather_host", "hsl_set_device", "hsl_gather_check"]


class _Pinned:
    """Owner of one hsl_pinned_alloc block (freed when the last numpy view goes away)."""

    def __init__(self, nbytes):
        self.ptr = _load().hsl_pinned_alloc(nbytes)
        if not self.ptr:
            raise HslError("hsl_pinned_alloc(%d): %s" % (nbytes, _load().hsl_last_error().decode()))
        self.nbytes = nbytes

    def __del__(self):
        try:
            if self.ptr:
                _load().hsl_pinned_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


def pinned_empty(shape, dtype=np.float64):
    """numpy array in page-locked host memory (hsl_pinned_alloc): pass such arrays as `out=` of the per-frame entries,
    or as their inputs, for copies at the full PCIe rate."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) * dtype.itemsize
    owner = _Pinned(max(n, 1))
    buf = (C.c_char * owner.nbytes).from_address(owner.ptr)
    buf._hsl_owner = owner  # the array references buf (buffer protocol), buf keeps the allocation alive
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


def _out(out, key, shape, dtype=np.float64):
    """Caller-supplied output array (checked) or a fresh pageable one."""
    if out is not None and key in out:
        a = out[key]
        if a.shape != tuple(shape) or a.dtype != np.dtype(dtype) or not a.flags.c_contiguous:
            raise ValueError("out[%r] must be a C-contiguous %s array of shape %s" % (key, np.dtype(dtype), tuple(shape)))
        return a
    return np.empty(shape, dtype)


def _check(rc):
    if rc != 0:
        raise HslError("hsl error %d: %s" % (rc, _load().hsl_last_error().decode()))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def make_params(torso_pos=(0, 0, 0), torso_angles=(0, 0, 0), step_duration=1.0, period=3.0, step_length=0.5,
                step_height=0.1, curvature=0.0, shift_type=-1, shift_value=0.0):
    """One candidate row (13 doubles), the scalars of the reference's pgsconfigparams."""
    return np.array(list(torso_pos) + list(torso_angles) + [step_duration, period, step_length, step_height, curvature,
                                                             shift_type, shift_value], dtype=np.float64)


def load_preset(config_fname, setup_id):
    """modelplayer::get_rec_str + get_pgs_config_params (player.cpp:170-208, 230-244): preset row -> (params, xml name)."""
    with open(config_fname) as f:
        for line in f:
            tok = line.split()
            if not tok or int(tok[0]) != setup_id:
                continue
            kw, xml, i = {}, None, 1
            while i < len(tok):
                key = tok[i]
                if key == "xml_file":
                    xml = tok[i + 1]; i += 2
                elif key in ("torso_pos", "torso_angles"):
                    kw[key] = tuple(float(v) for v in tok[i + 1:i + 4]); i += 4
                elif key in ("step_duration", "period", "step_length", "step_height", "curvature"):
                    kw[key] = float(tok[i + 1]); i += 2
                elif key == "lateral_foot_shift":
                    kw["shift_type"], kw["shift_value"] = 0, float(tok[i + 1]); i += 2
                elif key == "radial_foot_shift":
                    kw["shift_type"], kw["shift_value"] = 1, float(tok[i + 1]); i += 2
                else:
                    raise HslError("ERROR: unknown key " + key)
            return make_params(**kw), xml
    raise HslError("ERROR: no string with rec_id = %d" % setup_id)


class Model:
    """kinematicmodel + liksolver + periodic's static data, flattened and resident for the kernels."""

    def __init__(self, xml_path):
        lib = _load()
        h = C.c_void_p()
        _check(lib.hsl_model_load_xml(os.fspath(xml_path).encode(), C.byref(h)))
        self._h = h
        d = np.zeros(6, np.int32)
        _check(lib.hsl_model_dims(h, _p(d)))
        self.n, self.nf, self.nmj, self.config_dim, self.ntrunk, self.lik_index = (int(v) for v in d)
        self.rcap = lib.hsl_model_rcap(h)
        self.xml_path = os.fspath(xml_path)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                _load().hsl_model_free(self._h)
                self._h = None
        except Exception:
            pass

    def set_rec_transform(self, transl=None, eas=None):
        """pergensetup::set_rec_rotation / set_rec_transform (pergen.cpp:309-320): rigid map applied to every generated
        frame record of the following eval_gaits* calls; no arguments switch it off."""
        tr = None if transl is None else np.ascontiguousarray(transl, np.float64)
        ea = None if eas is None else np.ascontiguousarray(eas, np.float64)
        _check(_load().hsl_set_rec_transform(self._h, _p(tr), _p(ea)))

    # ---- measurement helpers
    def set_tuning(self, fb=None, maxreg=None):
        """Cost-only kernel variant: frame slots per block (32 | 64) and register cap (64..255, or 1 for the persistent
        pipelined kernel).  None keeps the value in use (the per-model library default unless changed before)."""
        cur_fb, cur_mr = self.get_tuning()
        _check(_load().hsl_set_tuning(self._h, cur_fb if fb is None else int(fb), cur_mr if maxreg is None else int(maxreg)))

    def get_tuning(self):
        fb, mr = C.c_int(0), C.c_int(0)
        _check(_load().hsl_get_tuning(self._h, C.byref(fb), C.byref(mr)))
        return fb.value, mr.value

    def set_kernel_timing(self, on=True):
        _check(_load().hsl_set_kernel_timing(self._h, int(bool(on))))

    def last_kernel_ms(self):
        """(setup, per-frame kernel, finish) durations in ms of the last gait-evaluation chunk on this handle."""
        ms = (C.c_float * 3)()
        _check(_load().hsl_last_kernel_ms(self._h, ms))
        return float(ms[0]), float(ms[1]), float(ms[2])

    def set_max_slots(self, max_slots):
        """Frame slots per launch: larger batches run as consecutive chunks (bounded workspace, same results)."""
        _check(_load().hsl_set_max_slots(self._h, int(max_slots)))

    def launch_count(self):
        return int(_load().hsl_launch_count(self._h))

    def pod_bytes(self):
        size = _load().hsl_model_pod(self._h, None, 0)
        buf = np.zeros(size, np.uint8)
        _load().hsl_model_pod(self._h, _p(buf), size)
        return buf

    # ---- batch evaluation, host buffers (the reference-facing call)
    def eval_gaits(self, params, n_t, flags=0, out=None):
        """hsl_eval_gaits_host.  params in page-locked memory (pinned_empty) are copied to the device from where they lie;
        out: optional dict of preallocated result arrays."""
        params = np.ascontiguousarray(params, np.float64).reshape(-1, HSL_NPARAM)
        c = params.shape[0]
        out = dict(cot=_out(out, "cot", (c,)), work=_out(out, "work", (c,)), min_cfz=_out(out, "min_cfz", (c,)), max_mu=_out(out, "max_mu", (c,)),
                   status=_out(out, "status", (c,), np.int32))
        _check(_load().hsl_eval_gaits_host(self._h, c, n_t, _p(params), flags, _p(out["cot"]), _p(out["work"]),
                                           _p(out["min_cfz"]), _p(out["max_mu"]), _p(out["status"])))
        return out

    def eval_gaits_detail(self, params, n_t, flags=0, out=None):
        """out: optional dict of preallocated arrays (e.g. from pinned_empty) for any of the result keys."""
        params = np.ascontiguousarray(params, np.float64).reshape(-1, HSL_NPARAM)
        c = params.shape[0]
        shapes = dict(cot=(c,), work=(c,), min_cfz=(c,), max_mu=(c,), traj=(c, n_t + 4, self.config_dim), x=(c, n_t, 6 * self.n),
                      z=(c, n_t, 3 * self.nf), tau=(c, n_t, self.nmj))
        res = {k: _out(out, k, sh) for k, sh in shapes.items()}
        res["status"] = _out(out, "status", (c,), np.int32)
        res["contacts"] = _out(out, "contacts", (c, n_t, self.nf), np.uint8)
        _check(_load().hsl_eval_gaits_detail_host(self._h, c, n_t, _p(params), flags, *[_p(res[k]) for k in (
            "cot", "work", "min_cfz", "max_mu", "status", "traj", "x", "z", "tau", "contacts")]))
        return res

    def eval_trajectories(self, traj, dt, n_t, out=None):
        traj = np.ascontiguousarray(traj, np.float64).reshape(-1, n_t + 5, self.config_dim)
        c = traj.shape[0]
        dt = np.ascontiguousarray(np.broadcast_to(np.asarray(dt, np.float64), (c,)))
        shapes = dict(work=(c,), min_cfz=(c,), max_mu=(c,), x=(c, n_t, 6 * self.n), z=(c, n_t, 3 * self.nf), tau=(c, n_t, self.nmj))
        res = {k: _out(out, k, sh) for k, sh in shapes.items()}
        res["status"] = _out(out, "status", (c,), np.int32)
        _check(_load().hsl_eval_trajectories_host(self._h, c, n_t, _p(traj), _p(dt), _p(res["work"]), _p(res["min_cfz"]),
                                                  _p(res["max_mu"]), _p(res["status"]), _p(res["x"]), _p(res["z"]), _p(res["tau"])))
        return res

    def solve_frames(self, pos, jpos, jzaxis, mom_rate, ang_mom_rate, fpos, contacts, out=None):
        """forcetorquesolver::solve_forcetorques + get_motor_torques on populated dynrecords."""
        arrs = [np.ascontiguousarray(a, np.float64) for a in (pos, jpos, jzaxis, mom_rate, ang_mom_rate, fpos)]
        contacts = np.ascontiguousarray(contacts, np.uint8)
        f = arrs[0].shape[0]
        res = dict(x=_out(out, "x", (f, 6 * self.n)), z=_out(out, "z", (f, 3 * self.nf)), tau=_out(out, "tau", (f, self.nmj)),
                   status=_out(out, "status", (f,), np.int32))
        _check(_load().hsl_solve_frames_host(self._h, f, *[_p(a) for a in arrs], _p(contacts), _p(res["x"]), _p(res["z"]),
                                             _p(res["tau"]), _p(res["status"])))
        return res

    def gait_records(self, params, times, flags=0):
        """pergensetup::set_rec (pergen.cpp:225-239): records [C][len(times)][6+3nf] of the candidates at the given times."""
        params = np.ascontiguousarray(params, np.float64).reshape(-1, HSL_NPARAM)
        times = np.ascontiguousarray(times, np.float64).reshape(-1)
        c = params.shape[0]
        rec = np.empty((c, times.shape[0], 6 + 3 * self.nf)); status = np.empty(c, np.int32)
        _check(_load().hsl_gait_records_host(self._h, c, _p(params), times.shape[0], _p(times), flags, _p(rec), _p(status)))
        return dict(rec=rec, status=status)

    def ik_records(self, rec, flags=0):
        """kinematicmodel::set_jvalues_with_lik + get_jvalues: joint values [n][config_dim] of records [n][6+3nf]."""
        rec = np.ascontiguousarray(rec, np.float64).reshape(-1, 6 + 3 * self.nf)
        n = rec.shape[0]
        q = np.empty((n, self.config_dim)); status = np.empty(n, np.int32)
        _check(_load().hsl_ik_records_host(self._h, n, _p(rec), flags, _p(q), _p(status)))
        return dict(q=q, status=status)

    def set_fall_variant(self, variant):
        """Kernel of fall_sweep: 1 (default) a warp per world, 0 a thread per world."""
        _check(_load().hsl_set_fall_variant(self._h, int(variant)))

    def fall_sweep(self, params, n_steps, kick_step=None, kick_dv=None, play_dt=0.02, t0=0.0, hc=0.7, tmin=0.1, n_worlds=None, want_traj=False):
        """hsl_fall_sweep_host: n_worlds copies of the reference's position-control loop with one torso kick each
        (kick_step [W], kick_dv [W][3]); returns fell, t_end, final_z, status (+ traj [W][n_steps][3]) and kernel_ms."""
        ks = None if kick_step is None else np.ascontiguousarray(kick_step, np.int32)
        kv = None if kick_dv is None else np.ascontiguousarray(kick_dv, np.float64).reshape(-1, 3)
        w = n_worlds or (ks.shape[0] if ks is not None else kv.shape[0] if kv is not None else 1)
        params = np.ascontiguousarray(params, np.float64)
        out = dict(fell=np.zeros(w, np.uint8), t_end=np.zeros(w), final_z=np.zeros(w), status=np.zeros(w, np.int32))
        traj = np.zeros((w, n_steps, 3)) if want_traj else None
        ms = C.c_float(0)
        _check(_load().hsl_fall_sweep_host(self._h, w, _p(params), play_dt, t0, n_steps, _p(ks), _p(kv), hc, tmin, _p(out["fell"]), _p(out["t_end"]),
                                           _p(out["final_z"]), _p(out["status"]), _p(traj), C.cast(C.byref(ms), C.c_void_p)))
        out["kernel_ms"] = float(ms.value)
        if want_traj:
            out["traj"] = traj
        return out

    def tables(self):
        """periodic::set_dynparts data: parent ids, foot / limb-top body ids, masses, COM and foot-point offsets."""
        out = dict(parent=np.zeros(self.n, np.int32), footis=np.zeros(self.nf, np.int32), limb_top=np.zeros(self.nf, np.int32),
                   masses=np.zeros(self.n), com_offset=np.zeros((self.n, 3)), foot_offset=np.zeros((self.nf, 3)))
        _check(_load().hsl_model_tables(self._h, *[_p(out[k]) for k in ("parent", "footis", "limb_top", "masses", "com_offset", "foot_offset")]))
        return out

    def fk_records(self, q):
        """kinematicmodel::set_jvalues + recompute_modelnodes: ground frames of the bodies and of their joints, column-major
        4x4 `affine`s [n][bodies][16] each, for joint values q [n][config_dim]."""
        q = np.ascontiguousarray(q, np.float64).reshape(-1, self.config_dim)
        n = q.shape[0]
        A = np.empty((n, self.n, 16)); J = np.empty((n, self.n, 16))
        _check(_load().hsl_fk_records_host(self._h, n, _p(q), _p(A), _p(J)))
        return dict(A_ground=A, J_A_ground=J)

    def solve_forces(self, pos, jpos, jzaxis, mom_rate, ang_mom_rate, fpos, torques):
        """forcetorquesolver::solve_forces (ftsolver.cpp:331-378): contact forces of all feet for given motor torques."""
        arrs = [np.ascontiguousarray(a, np.float64) for a in (pos, jpos, jzaxis, mom_rate, ang_mom_rate, fpos)]
        f = arrs[0].shape[0]
        torques = np.ascontiguousarray(torques, np.float64).reshape(f, self.nmj)
        out = dict(z=np.empty((f, 3 * self.nf)), status=np.empty(f, np.int32))
        _check(_load().hsl_solve_forces_host(self._h, f, *[_p(a) for a in arrs], _p(torques), _p(out["z"]), _p(out["status"])))
        return out

    def solve_forces_gait(self, params, n_t, torques, flags=0):
        """periodic::solve_contforces_given_torques (periodic.cpp:369-374) on every solved frame of generated gaits:
        torques [C][n_t][nmj] -> z [C][n_t][3nf]."""
        params = np.ascontiguousarray(params, np.float64).reshape(-1, HSL_NPARAM)
        c = params.shape[0]
        torques = np.ascontiguousarray(torques, np.float64).reshape(c, n_t, self.nmj)
        out = dict(z=np.empty((c, n_t, 3 * self.nf)), status=np.empty(c, np.int32))
        _check(_load().hsl_solve_forces_gait_host(self._h, c, n_t, _p(params), flags, _p(torques), _p(out["z"]), _p(out["status"])))
        return out

    # ---- device-pointer entry (inputs already in HBM); arguments are integer device addresses
    def eval_gaits_device(self, n_cand, n_t, d_params, d_cot=0, d_work=0, d_min_cfz=0, d_max_mu=0, d_status=0, flags=0,
                          stream=0):
        _check(_load().hsl_eval_gaits(self._h, n_cand, n_t, d_params, flags, d_cot or None, d_work or None,
                                      d_min_cfz or None, d_max_mu or None, d_status or None, stream or None))


    def eval_gaits_scatter(self, gather, n_cand, n_t, d_params, d_cot=0, d_work=0, d_min_cfz=0, d_max_mu=0, d_status=0, flags=0, stream=0):
        """hsl_eval_gaits_scatter: evaluate this rank's candidates; the finish kernel stores costs and status into every rank's
        gather buffer and raises this rank's flags.  Read with gather.select_best() or gather.wait()."""
        _check(_load().hsl_eval_gaits_scatter(self._h, gather._g, n_cand, n_t, d_params or None, flags, d_cot or None, d_work or None,
                                              d_min_cfz or None, d_max_mu or None, d_status or None, stream or None))

    def eval_gaits_gather(self, gather, n_cand, n_t, d_params, d_cot=0, d_work=0, d_min_cfz=0, d_max_mu=0, d_status=0, flags=0, stream=0):
        """hsl_eval_gaits_gather: evaluate this rank's n_cand candidates and all-gather costs and status over peer memory in
        the same launches.  Returns the integer device addresses of the [world][n_per_rank] cost (float64) and status (int32)
        arrays in this rank's gather buffer; they are complete for work queued on `stream` afterwards and stay valid until
        the second-next call on `gather`."""
        all_cot, all_st = C.c_void_p(), C.c_void_p()
        _check(_load().hsl_eval_gaits_gather(self._h, gather._g, n_cand, n_t, d_params or None, flags, d_cot or None, d_work or None,
                                             d_min_cfz or None, d_max_mu or None, d_status or None, C.byref(all_cot), C.byref(all_st),
                                             stream or None))
        return all_cot.value, all_st.value

    def eval_trajectories_device(self, n_cand, n_t, d_traj, d_dt, d_work=0, d_min_cfz=0, d_max_mu=0, d_status=0, d_x=0, d_z=0, d_tau=0,
                                 stream=0):
        """hsl_eval_trajectories: the L2 entry on device-resident trajectories (integer device addresses, row-major layouts of
        the host form; queued on `stream`, not synchronised)."""
        _check(_load().hsl_eval_trajectories(self._h, n_cand, n_t, d_traj, d_dt, d_work or None, d_min_cfz or None, d_max_mu or None,
                                             d_status or None, d_x or None, d_z or None, d_tau or None, stream or None))

    def solve_frames_device(self, n_frames, d_pos, d_jpos, d_jzaxis, d_mom_rate, d_ang_mom_rate, d_fpos, d_contacts, d_x=0, d_z=0, d_tau=0,
                            d_status=0, stream=0):
        """hsl_solve_frames: the L1 entry on device-resident dynrecord arrays (integer device addresses)."""
        _check(_load().hsl_solve_frames(self._h, n_frames, d_pos, d_jpos, d_jzaxis, d_mom_rate, d_ang_mom_rate, d_fpos, d_contacts,
                                        d_x or None, d_z or None, d_tau or None, d_status or None, stream or None))


class Gather:
    """The peer-memory all-gather of the costs (hsl_gather_create / _connect): this rank's gather buffer for `n_per_rank`
    candidates per rank, mapped by every rank of the job over NVLink.  `exchange(handle_bytes) -> [handle_bytes of every
    rank, in rank order]` is the launcher's side channel (torch.distributed.all_gather_object, MPI, files ...); it also acts
    as the barrier between creating and mapping the buffers."""

    def __init__(self, rank, world, n_per_rank, exchange):
        self._g = C.c_void_p()
        mine = (C.c_char * 64)()
        rc = _load().hsl_gather_create(world, rank, n_per_rank, C.byref(self._g), mine)
        self.rank, self.world, self.n_per_rank = rank, world, int(n_per_rank)
        handles = exchange(bytes(mine.raw) if rc == 0 else b"")   # a rank that failed still takes part in the exchange
        _check(rc)
        if len(handles) != world or any(h is None or len(h) != 64 for h in handles):
            self.free()
            raise HslError("exchange() must return the 64-byte handle of every rank (a rank failed to create its buffer?)")
        rc = _load().hsl_gather_connect(self._g, (C.c_char * (64 * world)).from_buffer_copy(b"".join(handles)))
        if rc != 0:
            msg = _load().hsl_last_error().decode()
            self.free()
            raise HslError("hsl error %d: %s" % (rc, msg))

    def select_best(self, d_index=0, d_value=0, stream=0):
        """hsl_gather_select_best: argmin over the costs of the latest scatter; the kernel itself waits for the ranks' flags."""
        _check(_load().hsl_gather_select_best(self._g, d_index or None, d_value or None, stream or None))

    def wait(self, stream=0):
        """hsl_gather_wait: device addresses of the gathered [world][n_per_rank] cost / status arrays of the latest scatter."""
        all_cot, all_st = C.c_void_p(), C.c_void_p()
        _check(_load().hsl_gather_wait(self._g, C.byref(all_cot), C.byref(all_st), stream or None))
        return all_cot.value, all_st.value

    def check(self):
        """hsl_gather_check: synchronises the device; raises if a wait on this object timed out (a rank never arrived)."""
        _check(_load().hsl_gather_check(self._g))

    def free(self):
        """Unmap and free.  All ranks must be done with the buffers (barrier before)."""
        if self._g:
            _check(_load().hsl_gather_free(self._g))
            self._g = C.c_void_p()


def torch_gather(dist, n_per_rank, fallback=False):
    """Gather object for the ranks of an initialised torch.distributed job (handles exchanged with all_gather_object).
    fallback=True: when any rank cannot create or map the buffers (no peer access between the GPUs, IPC not permitted in
    the container ...) every rank frees what it has and None is returned, so that the caller can use the NCCL form --
    the decision is collective.  fallback=False: raise on this rank."""
    def exchange(mine):
        box = [None] * dist.get_world_size()
        dist.all_gather_object(box, mine)
        return box
    if not fallback:
        return Gather(dist.get_rank(), dist.get_world_size(), n_per_rank, exchange)
    g, err = None, None
    try:
        g = Gather(dist.get_rank(), dist.get_world_size(), n_per_rank, exchange)
    except Exception as e:   # HslError, or a malformed exchange because another rank failed before it
        err = e
    oks = [None] * dist.get_world_size()
    dist.all_gather_object(oks, err is None)
    if all(oks):
        return g
    if g is not None:
        g.free()
    return None


class NcclComm:
    """An ncclComm_t made through the C ABI (hsl_nccl_unique_id / hsl_nccl_comm_init), for hsl_allgather_costs.
    `exchange(id_bytes_or_None) -> id_bytes` hands rank 0's 128-byte id to every rank (any side channel: a file, MPI,
    torch.distributed.broadcast_object_list ...)."""

    def __init__(self, rank, world, exchange):
        ident = (C.c_char * 128)()
        if rank == 0:
            _check(_load().hsl_nccl_unique_id(ident))
        raw = exchange(bytes(ident.raw) if rank == 0 else None)
        ident = (C.c_char * 128).from_buffer_copy(raw)
        self._c = C.c_void_p()
        _check(_load().hsl_nccl_comm_init(C.byref(self._c), world, ident, rank))
        self.rank, self.world = rank, world

    def allgather_costs(self, d_local, n_per_rank, d_all, stream=0):
        _check(_load().hsl_allgather_costs(self._c, d_local, n_per_rank, d_all, stream or None))

    def destroy(self):
        if self._c:
            _check(_load().hsl_nccl_comm_destroy(self._c))
            self._c = C.c_void_p()


def nccl_allgather_selftest(rank, world, dist, dev):
    """hsl_allgather_costs over a communicator made through the C ABI, against torch.distributed's all-gather."""
    import torch

    def exchange(raw):
        box = [raw]
        dist.broadcast_object_list(box, src=0)
        return box[0]
    comm = NcclComm(rank, world, exchange)
    n = 1000
    local = torch.arange(n, dtype=torch.float64, device=dev) + 1e6 * rank
    mine = torch.empty(world * n, dtype=torch.float64, device=dev)
    ref = torch.empty(world * n, dtype=torch.float64, device=dev)
    comm.allgather_costs(local.data_ptr(), n, mine.data_ptr(), torch.cuda.current_stream().cuda_stream)
    dist.all_gather_into_tensor(ref, local)
    torch.cuda.synchronize()
    assert torch.equal(mine, ref), "hsl_allgather_costs differs from torch.distributed.all_gather_into_tensor"
    comm.destroy()


def select_best_device(d_cost, n, d_index, d_value=0, stream=0):
    """Argmin over device-resident costs (NaN skipped); arguments are integer device addresses."""
    _check(_load().hsl_select_best(d_cost, n, d_index or None, d_value or None, stream or None))


def select_topk_device(d_cost, n, k, d_index, d_value=0, stream=0):
    """The k cheapest valid candidates in stable ascending order (NaN skipped, -1 / NaN past the valid ones); integer device addresses."""
    _check(_load().hsl_select_topk(d_cost, n, k, d_index or None, d_value or None, stream or None))


def dfma_probe(blocks=148 * 8, threads=256, iters=4096):
    """Measured FP64 FMA throughput of the current device in TFLOP/s (roofline denominator of this path)."""
    tf = C.c_double()
    ms = C.c_float()
    _check(_load().hsl_dfma_probe(blocks, threads, iters, C.byref(tf), C.byref(ms)))
    return tf.value, ms.value


def math_selftest(a, b):
    """Device results of the kernels' branch-free primitives next to the library ones: dict of (ours, reference) pairs."""
    a = np.ascontiguousarray(a, np.float64)
    b = np.ascontiguousarray(b, np.float64)
    out = np.empty((10, a.size))
    _check(_load().hsl_math_selftest(a.size, _p(a), _p(b), _p(out)))
    return dict(div=(out[0], out[1]), sqrt=(out[2], out[3]), atan2=(out[4], out[5]), sin=(out[6], out[7]), cos=(out[8], out[9]))


def measure_cot(model, params, n_t, flags=0):
    """modelplayer::measure_cot (player.cpp:269-285) for one candidate."""
    r = model.eval_gaits(params, n_t, flags)
    return float(r["cot"][0])


def measure_cot_sweep(model, params, n_t, param_name, val0, val1, n_val, flags=0, verbose=False):
    """modelplayer::measure_cot_sweep (player.cpp:311-321) with pgssweeper::sweep/next (pergen.cpp:417-449):
    n_val+1 candidates val0 + i*(val1-val0)/n_val, evaluated as one batch."""
    if param_name not in SWEEP_NAMES:
        raise HslError("ERROR: cannot sweep over " + param_name)
    delval = (val1 - val0) / n_val
    vals = np.array([val0 + i * delval for i in range(n_val + 1)])
    batch = np.tile(np.asarray(params, np.float64), (n_val + 1, 1))
    batch[:, SWEEP_NAMES[param_name]] = vals
    r = model.eval_gaits(batch, n_t, flags)
    if verbose:
        for v, c in zip(vals, r["cot"]):
            print("val = %g COT = %g" % (v, c))
    return vals, r["cot"]
