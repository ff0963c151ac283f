"""BASELINE configs[4] at its per-GPU size: 125 000 perturbed worlds (1e6 over 8 GPUs) x 500 steps of 0.02 s, hexapod preset 8
under PD position control, one random torso kick each (|dv| U[0, 16], direction uniform, kick time U[0.5 s, 3.5 s)).
Run on the GPU box: python tools/fall_config5.py [worlds] [steps]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.getcwd())
sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import hslabs_b200 as hsl
from conftest import PRESETS, model_xml

w = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 500
params, name = hsl.load_preset(PRESETS, 8)
m = hsl.Model(model_xml(name))
rng = np.random.default_rng(20261019)
mag, th = rng.uniform(0, 16, w), rng.uniform(0, 2 * np.pi, w)
kicks = np.stack([mag * np.cos(th), mag * np.sin(th), np.zeros(w)], axis=1)
ks = rng.integers(25, 175, w).astype(np.int32)
m.fall_sweep(params, 5, ks[:1024], kicks[:1024])
t0 = time.perf_counter()
r = m.fall_sweep(params, n_steps, ks, kicks)
wall = time.perf_counter() - t0
steps_done = np.where(r["fell"] == 1, np.round(r["t_end"] / 0.02), n_steps).sum()
hist, edges = np.histogram(r["t_end"][r["fell"] == 1], bins=10, range=(0, n_steps * 0.02))
by_mag, _ = np.histogram(mag[r["fell"] == 1], bins=8, range=(0, 16))
all_mag, _ = np.histogram(mag, bins=8, range=(0, 16))
print(json.dumps({"worlds": w, "steps": n_steps, "kernel_ms": r["kernel_ms"], "wall_s": wall, "fall_rate": float(r["fell"].mean()),
                  "world_steps_simulated": float(steps_done), "world_steps_per_s": float(steps_done / (r["kernel_ms"] * 1e-3)),
                  "status_all_zero": bool((r["status"] == 0).all()), "time_to_fall_hist_1s_bins": hist.tolist(),
                  "fall_rate_by_kick_magnitude_2_wide_bins": (by_mag / np.maximum(all_mag, 1)).round(4).tolist()}))
