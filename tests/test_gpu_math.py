"""GPU tier: accuracy of the kernels' branch-free FP64 primitives (hslabs_b200/csrc/hsl_fastmath.h) against the CUDA
library / IEEE operations, on the argument ranges the evaluation path produces."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def ulps(a, b):
    return np.abs(a - b) / np.spacing(np.abs(b))


def test_div_sqrt(hsl):
    from hslabs_b200 import api
    rng = np.random.default_rng(0)
    a = rng.uniform(-50, 50, 200000) * 10.0 ** rng.integers(-6, 3, 200000)
    b = rng.uniform(0.01, 50, 200000) * rng.choice([-1.0, 1.0], 200000) * 10.0 ** rng.integers(-4, 3, 200000)
    r = api.math_selftest(a, b)
    assert ulps(*r["div"]).max() <= 1.0
    assert ulps(*r["sqrt"]).max() <= 1.0
    z = api.math_selftest(np.zeros(4), np.ones(4))
    assert (z["sqrt"][0] == 0).all() and (z["div"][0] == 0).all()


def test_sincos_0_pi(hsl):
    from hslabs_b200 import api
    x = np.concatenate([np.linspace(0, np.pi, 200001), np.pi * np.random.default_rng(1).uniform(0, 1, 100000), [0.0, np.pi / 2, np.pi]])
    r = api.math_selftest(x, np.ones_like(x))
    assert np.abs(r["sin"][0] - r["sin"][1]).max() <= 2.3e-16   # absolute: what the foot target needs
    assert np.abs(r["cos"][0] - r["cos"][1]).max() <= 2.3e-16
    ref = np.sin(x)
    big = np.abs(ref) > 1e-3
    assert ulps(r["sin"][0][big], r["sin"][1][big]).max() <= 2.0


def test_small_angle_and_fallback(hsl):
    from hslabs_b200 import api
    rng = np.random.default_rng(2)
    d = np.concatenate([rng.uniform(-0.49, 0.49, 200000), rng.uniform(-3.1, 3.1, 50000), [0.0, 1e-9, -1e-9, 1e-300]])
    rad = rng.uniform(0.5, 2.0, d.size)
    r = api.math_selftest(rad * np.sin(d), rad * np.cos(d))
    ours, ref = r["atan2"]
    assert np.abs(ours - ref).max() <= 4.5e-16
    small = np.abs(ref) > 1e-6
    assert ulps(ours[small], ref[small]).max() <= 4.0
