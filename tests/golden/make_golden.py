#!/usr/bin/env python3
"""Freezes golden fixtures from the REFERENCE's own code: oracle/_ref (the sources of /root/reference compiled unmodified
against shim headers, oracle/Makefile target `ref`), run on the reference's own hexapod.xml / myant.xml / spider.xml and
pgs_config.txt.  /root/reference does not exist on the GPU box, so the vectors are committed; this is the script that
made them (round 1 froze them from the oracle restatement; the two agree to 1e-15, tests/test_ref_pins.py).

  sweep_preset8_period.json   COT vs period 3..18, 16 values, n_t=20, preset 8 (the call at main.cpp:69)
  preset_costs.json           COT / work of the presets at n_t = 20 and 150 (config 1 uses ids 0,1,2,9,15)
  frames_<model>.npz          per-frame trajectory / x / z / motor torques of one preset per model (n_t=20)

Run from the repo root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref as orc  # noqa: E402  (same Model interface as oracle.orc)

HERE = os.path.dirname(os.path.abspath(__file__))
MODELS = "/root/reference"
PRESETS = os.path.join(MODELS, "pgs_config.txt")
SOURCE = "oracle/_ref: /root/reference sources (matrix, core, model, visualization, lik, pergen, dynrec, ftsolver, periodic, player .cpp) " \
         "compiled unmodified against oracle/shim, run on the reference's own XML models and pgs_config.txt"


def load_preset(pid):
    m = orc.Model(os.path.join(MODELS, "hexapod.xml"))
    return m.load_preset(PRESETS, pid)


def main():
    params, name = load_preset(8)
    vals, cots = orc.Model(os.path.join(MODELS, name)).measure_cot_sweep(params, 20, "period", 3, 18, 15)
    json.dump({"source": SOURCE, "preset": 8, "n_t": 20, "param": "period", "vals": list(vals), "cots": list(cots)},
              open(os.path.join(HERE, "sweep_preset8_period.json"), "w"), indent=1)
    rows = []
    for pid in (0, 1, 2, 3, 7, 8, 9, 10, 15, 17, 20, 23, 24, 25, 26, 27):
        params, name = load_preset(pid)
        m = orc.Model(os.path.join(MODELS, name))
        for n_t in (20, 150):
            r = m.measure_cot(params, n_t)
            assert r["status"] == 0
            rows.append({"id": pid, "model": name, "n_t": n_t, "cot": r["cot"], "work": r["work"]})
    json.dump({"source": SOURCE, "rows": rows}, open(os.path.join(HERE, "preset_costs.json"), "w"), indent=1)
    for pid in (8, 9, 24):
        params, name = load_preset(pid)
        r = orc.Model(os.path.join(MODELS, name)).measure_cot(params, 20, detail=True)
        np.savez_compressed(os.path.join(HERE, "frames_%s.npz" % name[:-4]), preset=pid, params=params, traj=r["traj"], x=r["x"],
                            z=r["z"], tau=r["tau"], cot=r["cot"], work=r["work"], complete=r["complete"], source=SOURCE)


if __name__ == "__main__":
    main()
