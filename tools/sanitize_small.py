"""Small invocation of every kernel variant, for compute-sanitizer runs (one tool per call)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hslabs_b200 as hsl  # noqa: E402

for name, pid in (("hexapod", 8), ("myant", 9), ("spider", 24)):
    m = hsl.Model(hsl.model_path(name))
    p, _ = hsl.load_preset(os.path.join(ROOT, "hslabs_b200", "models", "pgs_presets.txt"), pid)
    batch = np.tile(p, (5, 1))
    batch[:, 7] = np.linspace(2, 4, 5)
    for fb, mr in ((64, 128), (32, 128), (32, 255), (32, 96), (64, 1), (32, 1)):
        m.set_tuning(fb, mr)
        r = m.eval_gaits(batch, 23)
        assert (r["status"] == 0).all()
    d = m.eval_gaits_detail(batch[:2], 23)
    t = m.eval_trajectories(np.concatenate([d["traj"], d["traj"][:, -1:]], axis=1), p[7] / 23, 23)
    print(name, r["cot"][:2], d["cot"], t["work"])
print("done")
