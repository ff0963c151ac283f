"""CPU tier: the C-ABI library builds, loads, exports every symbol include/hsl.h declares, flattens the models
exactly as the oracle's tree says, and fails loudly (no CPU fallback) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import hostlib
from conftest import ROOT, model_xml


def test_library_exports_header_symbols(hsl):
    from hslabs_b200 import api, build
    build.build()
    lib = C.CDLL(api.lib_path())
    header = open(os.path.join(ROOT, "include", "hsl.h")).read()
    declared = set(re.findall(r"\b(hsl_[a-z_0-9]+)\s*\(", header))
    assert declared == set(api.exported_symbols())
    for sym in declared:
        assert hasattr(lib, sym), sym


@pytest.mark.parametrize("name,dims", [("hexapod", (22, 6, 18, 24, 4)), ("myant", (17, 4, 12, 18, 5)), ("spider", (19, 6, 18, 24, 1))])
def test_model_dims(hsl, name, dims):
    m = hsl.Model(hsl.model_path(name))
    assert (m.n, m.nf, m.nmj, m.config_dim, m.ntrunk) == dims
    assert m.rcap == 0.08


@pytest.mark.parametrize("name", ["hexapod", "myant", "spider"])
def test_packed_model_matches_oracle_tree(hsl, orc, name):
    """The flat constant block (hsl_model.h) against the oracle's restatement of the reference's pointer tree."""
    cons = orc.Model(model_xml(name)).constants()
    pod = hsl.Model(hsl.model_path(name)).pod_bytes()
    host = np.zeros(16384, np.uint8)
    hostlib.lib().hc_model_pod(model_xml(name).encode(), host.ctypes.data_as(C.c_void_p))
    assert np.array_equal(pod, host[:pod.size])  # same loader in the library and in the host emulation
    ints = pod[:24].view(np.int32)
    n, nf = int(ints[0]), int(ints[1])
    dbl = pod[24:].view(np.float64)
    assert dbl[0] == 0.08 and dbl[1] == 1.0
    # limbs: HslLimb = 3 x HslHinge (30 doubles each) + 11 doubles + 6 int32 ; trunk: 8 doubles + 2 int32 = 9 doubles
    trunk0 = 2 + 6
    limb0 = trunk0 + 8 * 9
    for l in range(nf):
        base = limb0 + l * (3 * 30 + 11 + 3)
        body = int(cons["limb_top"][l])
        for h in range(3):
            hb = base + 30 * h
            Rjp, tjp = dbl[hb:hb + 9].reshape(3, 3).T, dbl[hb + 9:hb + 12]
            Rpb, tpb = dbl[hb + 12:hb + 21].reshape(3, 3).T, dbl[hb + 21:hb + 24]
            com = dbl[hb + 24:hb + 27]
            A = cons["J_A_parent"][body].reshape(4, 4).T
            B = cons["A_pj_body"][body].reshape(4, 4).T
            assert np.array_equal(Rjp, A[:3, :3]) and np.array_equal(tjp, A[:3, 3])
            assert np.array_equal(Rpb, B[:3, :3]) and np.array_equal(tpb, B[:3, 3])
            assert np.array_equal(com, cons["A_body_geom"][body].reshape(4, 4).T[:3, 3])
            if h < 2:
                body = int(np.where(cons["parent"] == body)[0][0])
        assert body == cons["limb_foot"][l]
        assert np.array_equal(dbl[base + 90:base + 93], cons["capsule_to_pos"][body])


def test_bad_inputs_report_errors(hsl, tmp_path):
    with pytest.raises(hsl.HslError):
        hsl.Model(str(tmp_path / "missing.xml"))
    bad = tmp_path / "hexapod.xml"
    bad.write_text("<notmujoco/>")
    with pytest.raises(hsl.HslError):
        hsl.Model(str(bad))
    other = tmp_path / "walker.xml"
    other.write_text(open(hsl.model_path("hexapod")).read())
    with pytest.raises(hsl.HslError):  # no LIK solver for unknown model names (lik.cpp:12-16)
        hsl.Model(str(other))


def test_size_bounds_are_argument_errors(hsl):
    """Checked before any CUDA call: chunk bound limits, a candidate longer than one launch may be, per-frame entries
    beyond the kernels' 32-bit frame-slot range."""
    m = hsl.Model(hsl.model_path("hexapod"))
    for bad in (0, 4, 1 << 31):
        with pytest.raises(hsl.HslError, match="frame slots per launch"):
            m.set_max_slots(bad)
    m.set_max_slots(16)
    with pytest.raises(hsl.HslError, match="exceed the launch bound"):
        m.eval_gaits(hsl.make_params(torso_pos=(0, 0, -.1)), 20)
    m.set_max_slots(1 << 26)
    lib = hsl.api._load()
    one = np.zeros(13)
    assert lib.hsl_solve_forces_gait_host(m._h, 1 << 26, 60, one.ctypes.data, 0, one.ctypes.data, one.ctypes.data, None) == -1
    assert b"split the batch" in lib.hsl_last_error()
    assert lib.hsl_eval_trajectories_host(m._h, 1 << 26, 60, one.ctypes.data, one.ctypes.data, None, None, None, None, None, None, None) == -1
    assert b"split the batch" in lib.hsl_last_error()


def test_no_cpu_fallback(hsl):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    m = hsl.Model(hsl.model_path("hexapod"))
    p = hsl.make_params(torso_pos=(0, 0, -.1))
    with pytest.raises(hsl.HslError, match="CUDA"):
        m.eval_gaits(p, 20)
    # every compute entry of the Python mirror: the wrapper code runs up to the C call, which must refuse
    n, nf, nmj = m.n, m.nf, m.nmj
    z3 = np.zeros((2, n, 3)); zf = np.zeros((2, nf, 3))
    m.set_rec_transform((0, 0, 0), (0, 0, 0.1))
    m.set_rec_transform()
    for call in (lambda: m.eval_gaits_detail(p, 20),
                 lambda: m.eval_trajectories(np.zeros((1, 25, m.config_dim)), 0.1, 20),
                 lambda: m.solve_frames(z3, z3, z3, z3, z3, zf, np.ones((2, nf), np.uint8)),
                 lambda: m.solve_forces(z3, z3, z3, z3, z3, zf, np.zeros((2, nmj))),
                 lambda: m.solve_forces_gait(p, 20, np.zeros((1, 20, nmj))),
                 lambda: m.gait_records(p, [0.0, 0.5]),
                 lambda: m.ik_records(np.zeros((2, 6 + 3 * nf)))):
        with pytest.raises(hsl.HslError, match="CUDA"):
            call()


def test_preset_parser_matches_oracle(hsl, orc):
    from conftest import PRESETS
    for pid in range(28):
        a, xa = hsl.load_preset(PRESETS, pid)
        b, xb = orc.load_preset(PRESETS, pid)
        assert xa == xb and np.array_equal(a, b)


def test_product_never_touches_the_oracle():
    """No file of the product package may include, import, link or load anything under oracle/ or tests/hostcheck
    (comments may mention them)."""
    import re
    pkg = os.path.join(ROOT, "hslabs_b200")
    bad = re.compile(r'(#\s*include\s*[<"][^>"]*(oracle|orc_|hostcheck)|^\s*(from|import)\s+\S*(oracle|hostcheck)|liborc|libhslhost|CDLL\([^)]*(oracle|hostcheck))', re.M)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".h", ".cu", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                m = bad.search(txt)
                assert m is None, (f, m.group(0))
    # and the build recipe of the library lists no oracle source
    from hslabs_b200 import build
    assert all("oracle" not in s and "hostcheck" not in s for s in build.SOURCES + build.HEADERS)
