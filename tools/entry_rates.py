#!/usr/bin/env python3
"""Rates of the three nested entry levels through their host-buffer C-ABI calls ("also L1-only and
L2-only rates"), hexapod, wall clock around the call (H2D, kernels, D2H and the host-side transposes of the dumps):
  L3 cost-only   hsl_eval_gaits_host            candidates -> COT                       (the headline path)
  L3 detail      hsl_eval_gaits_detail_host     candidates -> COT + traj, x, z, tau, contacts per frame
  L2             hsl_eval_trajectories_host     joint trajectories -> work + x, z, tau per frame
  L1             hsl_solve_frames_host          populated dynrecords -> x, z, tau per frame (synthetic field values)
Usage: python tools/entry_rates.py [--candidates 512] [--frames 256]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def best_of(fn, reps=5):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return float(np.median(ts))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--candidates", type=int, default=512)
    ap.add_argument("--frames", type=int, default=256)
    args = ap.parse_args()
    import hslabs_b200 as hsl
    from bench import synth_candidates
    m = hsl.Model(hsl.model_path("hexapod"))
    p = synth_candidates(4 * args.candidates, 20261018)
    p = np.ascontiguousarray(p[m.eval_gaits(p, 20)["status"] == 0][:args.candidates])
    c, n_t = p.shape[0], args.frames
    nfr = c * n_t
    rows = []
    big = synth_candidates(16 * 4096, 1)
    rows.append(("L3 cost-only (65536 candidates)", big.shape[0] * n_t, best_of(lambda: m.eval_gaits(big, n_t)), big.nbytes, big.shape[0] * 36))
    det = m.eval_gaits_detail(p, n_t)
    out_b = sum(det[k].nbytes for k in ("traj", "x", "z", "tau", "contacts"))
    rows.append(("L3 detail", nfr, best_of(lambda: m.eval_gaits_detail(p, n_t)), p.nbytes, out_b))
    pin = {k: hsl.pinned_empty(det[k].shape, det[k].dtype) for k in ("traj", "x", "z", "tau", "contacts")}
    got = m.eval_gaits_detail(p, n_t, out=pin)
    assert all(np.array_equal(got[k], det[k]) for k in pin)
    rows.append(("L3 detail, pinned outputs", nfr, best_of(lambda: m.eval_gaits_detail(p, n_t, out=pin)), p.nbytes, out_b))
    traj = np.concatenate([det["traj"], det["traj"][:, -1:]], axis=1)  # n_t + 5 rows (the last one is never read by the solve)
    dt = p[:, 7] / n_t
    got = m.eval_trajectories(traj, dt, n_t)
    assert np.abs(got["tau"] - det["tau"]).max() <= 1e-9 * np.abs(det["tau"]).max()
    rows.append(("L2 trajectories", nfr, best_of(lambda: m.eval_trajectories(traj, dt, n_t)), traj.nbytes,
                 sum(got[k].nbytes for k in ("x", "z", "tau"))))
    pin2 = {k: pin[k] for k in ("x", "z", "tau")}
    trajp = hsl.pinned_empty(traj.shape); trajp[...] = traj
    rows.append(("L2 trajectories, pinned in / out", nfr, best_of(lambda: m.eval_trajectories(trajp, dt, n_t, out=pin2)), traj.nbytes,
                 sum(got[k].nbytes for k in ("x", "z", "tau"))))
    rng = np.random.default_rng(5)
    n, nf = m.n, m.nf
    pos = rng.normal(0, 0.3, (nfr, n, 3)); jpos = pos + rng.normal(0, 0.05, (nfr, n, 3))
    jz = rng.normal(0, 1, (nfr, n, 3)); jz /= np.linalg.norm(jz, axis=2, keepdims=True)
    mr = rng.normal(0, 0.1, (nfr, n, 3)); ar = rng.normal(0, 0.01, (nfr, n, 3))
    fpos = rng.normal(0, 0.5, (nfr, nf, 3)); con = (rng.uniform(size=(nfr, nf)) < 0.65).astype(np.uint8)
    f1 = m.solve_frames(pos, jpos, jz, mr, ar, fpos, con)
    rows.append(("L1 frame solve (synthetic fields)", nfr, best_of(lambda: m.solve_frames(pos, jpos, jz, mr, ar, fpos, con)),
                 5 * pos.nbytes + fpos.nbytes + con.nbytes, sum(f1[k].nbytes for k in ("x", "z", "tau"))))
    pin3 = {k: pin[k].reshape(f1[k].shape) for k in ("x", "z", "tau")}
    fin = []
    for a in (pos, jpos, jz, mr, ar, fpos):
        b = hsl.pinned_empty(a.shape); b[...] = a; fin.append(b)
    rows.append(("L1 frame solve, pinned in / out", nfr, best_of(lambda: m.solve_frames(*fin, con, out=pin3)),
                 5 * pos.nbytes + fpos.nbytes + con.nbytes, sum(f1[k].nbytes for k in ("x", "z", "tau"))))
    print("hexapod, %d candidates x %d frames (host buffers, wall clock, median of 5)" % (c, n_t))
    for name, frames, t, bi, bo in rows:
        print("  %-38s %9.3f ms  %.3e frame solves/s   in %.1f MB  out %.1f MB" % (name, t * 1e3, frames / t, bi / 1e6, bo / 1e6))


if __name__ == "__main__":
    main()
