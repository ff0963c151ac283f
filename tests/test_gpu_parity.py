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
