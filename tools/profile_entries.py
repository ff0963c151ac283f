#!/usr/bin/env python3
"""Exercises the non-headline kernels once each for an ncu capture (profiles/r02_ncu_entries_*):
the per-frame dump kernel (hsl_frames_kernel<..., DUMP>) behind hsl_eval_gaits_detail_host, the forces-from-torques
kernel, the top-k sort and the fall-sweep kernel.  Also prints their event-timed rates (run without ncu for those)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hslabs_b200 as hsl  # noqa: E402
import bench  # noqa: E402


def main():
    import torch
    m = hsl.Model(hsl.model_path("hexapod"))
    p = bench.synth_candidates(600, 1)
    p = p[m.eval_gaits(p, 20)["status"] == 0][:512]
    n_t = 256
    out = None
    for _ in range(2):
        t0 = time.perf_counter()
        out = m.eval_gaits_detail(p, n_t)
        dt = time.perf_counter() - t0
    frames = p.shape[0] * n_t
    print("detail entry (pageable host arrays): %.3g frame solves/s, %d B written per frame" % (frames / dt, 8 * (6 * m.n + 3 * m.nf + m.nmj) + m.nf))
    t0 = time.perf_counter()
    z = m.solve_forces_gait(p, n_t, out["tau"])
    print("forces-from-torques entry: %.3g frames/s" % (frames / (time.perf_counter() - t0)))
    c = torch.rand(65536, dtype=torch.float64, device="cuda")
    idx = torch.empty(4096, dtype=torch.int64, device="cuda")
    hsl.api.select_topk_device(c.data_ptr(), 65536, 4096, idx.data_ptr(), 0, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    params, _ = hsl.load_preset(os.path.join(os.path.dirname(hsl.model_path("hexapod")), "pgs_presets.txt"), 8)
    w = 16384
    rng = np.random.default_rng(3)
    kv = np.stack([rng.uniform(-10, 10, w), rng.uniform(-10, 10, w), np.zeros(w)], axis=1)
    r = m.fall_sweep(params, 10, rng.integers(0, 5, w).astype(np.int32), kv)
    print("fall sweep: %d worlds x 10 steps in %.1f ms = %.3g world-steps/s" % (w, r["kernel_ms"], w * 10 / (r["kernel_ms"] * 1e-3)))


if __name__ == "__main__":
    main()
