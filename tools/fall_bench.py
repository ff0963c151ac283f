"""Fall sweep: time both kernels (thread per world / warp per world) on the same kicks.  Run on the GPU box."""
import os
import sys

import numpy as np

sys.path.insert(0, os.getcwd())
sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import hslabs_b200 as hsl
from conftest import PRESETS, model_xml

params, name = hsl.load_preset(PRESETS, 8)
m = hsl.Model(model_xml(name))
rng = np.random.default_rng(1)
for w in (int(a) for a in (sys.argv[1:] or ["16384", "131072"])):
    mag, th = rng.uniform(0, 16, w), rng.uniform(0, 2 * np.pi, w)
    kicks = np.stack([mag * np.cos(th), mag * np.sin(th), np.zeros(w)], axis=1)
    ks = rng.integers(5, 45, w).astype(np.int32)
    for variant in [int(v) for v in os.environ.get("FALL_VARIANTS", "0,1").split(",")]:
        m.set_fall_variant(variant)
        m.fall_sweep(params, 50, ks[:256], kicks[:256])
        r = m.fall_sweep(params, 50, ks, kicks)
        print("worlds %d variant %d: %.2f ms, %.4g world-steps/s, fall rate %.4f" % (w, variant, r["kernel_ms"], w * 50 / (r["kernel_ms"] * 1e-3), r["fell"].mean()), flush=True)
