"""Conditioning yardstick (pin 8): the oracle headers compiled with real = __float128 against the FP64
oracle, and the kernel math (host emulation of the device code) against the same quad results.

What this pins:
  * the reference algorithm restated in FP64 carries ~1e-14 (n_t = 20) to ~2e-12 (n_t = 256) of round-off in x, z and
    tau (the +-2 finite-difference stencil amplifies by 1/dt^2), far inside the 1e-9 parity tolerance;
  * the device formulation (backward recursion + Schur complement) is as close to the quad result as the FP64
    restatement of the reference is -- it is not a less accurate algorithm;
  * the reference's min_cfz / max_mu statistics run over ALL feet (periodic.cpp:347-357); for swing feet the contact
    force is pure round-off (1e-16 in FP64, 1e-34 in quad), so those two numbers are not reproducible even between
    two precisions of the same code.  Over the feet on the ground (the product's definition) they are.
"""
import numpy as np
import pytest

import hostlib
from conftest import PRESETS, model_xml

CASES = [(8, 20), (1, 20), (24, 20), (12, 20), (9, 64)]


def _rel(a, b):
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def _angle_err(a, b):
    d = np.abs(a - b)
    d[:, 3:] = np.minimum(d[:, 3:], np.abs(2 * np.pi - d[:, 3:]))  # joint angles are defined modulo 2 pi (spider sits near +-pi)
    return float(d.max())


def _ground_stats(z, nf):
    zc = z.reshape(z.shape[0], nf, 3)
    on = np.abs(zc).max(axis=2) > 1e-9 * np.abs(zc).max()
    lam = zc[on]
    return lam[:, 2].min(), (np.hypot(lam[:, 0], lam[:, 1]) / lam[:, 2]).max(), on


@pytest.mark.parametrize("pid,n_t", CASES)
def test_double_vs_quad(orc, pid, n_t):
    params, name = orc.load_preset(PRESETS, pid)
    xml = model_xml(name)
    m = orc.Model(xml)
    d = m.measure_cot(params, n_t, detail=True)
    q = m.measure_cot_quad(params, n_t)
    assert d["status"] == 0 and q["status"] == 0
    assert abs(d["cot"] - q["cot"]) <= 1e-12 * abs(q["cot"])
    assert abs(d["work"] - q["work"]) <= 1e-12 * abs(q["work"])
    assert _angle_err(d["traj"], q["traj"]) < 1e-13
    for key in ("x", "z", "tau"):
        assert _rel(d[key], q[key]) < 1e-10, key
    cz_d, mu_d, on_d = _ground_stats(d["z"], m.nf)
    cz_q, mu_q, on_q = _ground_stats(q["z"], m.nf)
    assert np.array_equal(on_d, on_q)
    assert abs(cz_d - cz_q) <= 1e-10 * np.abs(q["z"]).max()
    assert abs(mu_d - mu_q) <= 1e-9 * max(1.0, abs(mu_q))
    # the kernel math against the same yardstick: no further from quad than a small multiple of the FP64 reference
    h = hostlib.eval_gaits(xml, params, n_t)
    assert h["status"][0] == 0
    for key in ("x", "z", "tau"):
        e_ref = _rel(d[key], q[key])
        e_dev = _rel(h[key][0], q[key])
        assert e_dev < max(8 * e_ref, 1e-13), (key, e_dev, e_ref)
    assert abs(h["cot"][0] - q["cot"]) <= 1e-12 * abs(q["cot"])
    assert abs(h["min_cfz"][0] - cz_q) <= 1e-10 * np.abs(q["z"]).max()
    assert abs(h["max_mu"][0] - mu_q) <= 1e-9 * max(1.0, abs(mu_q))


def test_reference_stats_over_swing_feet_are_roundoff(orc):
    """Documents why min_cfz / max_mu are defined over the feet on the ground: the reference's own numbers change
    with the arithmetic precision whenever a foot is in the air."""
    params, name = orc.load_preset(PRESETS, 8)
    m = orc.Model(model_xml(name))
    d = m.measure_cot(params, 20, detail=True)
    q = m.measure_cot_quad(params, 20)
    assert abs(d["min_cfz"]) < 1e-12 and abs(q["min_cfz"]) < 1e-28      # both are noise of the respective precision
    zc = d["z"].reshape(20, m.nf, 3)
    swing = np.abs(zc).max(axis=2) <= 1e-9 * np.abs(zc).max()
    assert swing.any() and np.abs(zc[swing]).max() < 1e-12


def test_illconditioned_candidates_follow_the_exact_algorithm(orc):
    """Almost collinear contact points make level 0 nearly rank deficient.  The reference's FP64 arithmetic then trips
    its rank-threshold retry loop (ftsolver.cpp:208-232) in some frames and returns something else than its own
    algorithm does in exact arithmetic; the kernel math agrees with the __float128 evaluation of the reference
    algorithm, not with the FP64 one, and flags the candidate (HSL_ST_ILLCOND: a level-0 LDL^T pivot below 1e-4 of
    the trace)."""
    from test_gpu_parity import _random_candidates
    n, n_t = 512, 48
    p = _random_candidates("myant", n, 20261018 + n_t)
    rng = np.random.default_rng(5)
    p[:, 10] = rng.uniform(-0.1, 0.1, n); p[:, 11] = 0; p[:, 12] = rng.uniform(0, 0.3, n)
    xml = model_xml("myant")
    m = orc.Model(xml)
    for c, deviates in ((91, True), (113, False), (114, True)):
        d = m.measure_cot(p[c], n_t)
        q = m.measure_cot_quad(p[c], n_t)
        h = hostlib.eval_gaits(xml, p[c], n_t)
        assert d["status"] == 0 and q["status"] == 0
        assert h["status"][0] == 16                                       # informational flag only
        assert abs(h["cot"][0] - q["cot"]) <= 1e-9 * abs(q["cot"])        # the exact algorithm's value
        assert (abs(d["cot"] - q["cot"]) > 1e-3 * abs(q["cot"])) == deviates
    # a benign neighbour: all three agree, no flag
    d = m.measure_cot(p[0], n_t); q = m.measure_cot_quad(p[0], n_t); h = hostlib.eval_gaits(xml, p[0], n_t)
    assert h["status"][0] == 0 and abs(h["cot"][0] - q["cot"]) <= 1e-12 * abs(q["cot"]) and abs(d["cot"] - q["cot"]) <= 1e-12 * abs(q["cot"])
