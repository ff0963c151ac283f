#!/usr/bin/env python3
"""Parity evidence on the GPU box: the default cost-only kernels and the detail entry against the CPU oracle on larger
random batches than the test tier uses, all three models, straight and curved gaits.  Writes a plain-text table
(committed as profiles/<round>_parity_report.txt).  The oracle is used here as the checker only.
Usage: python tools/parity_report.py [--candidates 512] [--out gpurun_out/r01_parity_report.txt]"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--candidates", type=int, default=512)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r01_parity_report.txt"))
    a = ap.parse_args()
    import hslabs_b200 as hsl
    from oracle import orc
    from test_gpu_parity import _random_candidates
    threads = os.cpu_count() or 8
    lines = ["# GPU (libhsl_b200, default kernels through the C ABI) vs CPU oracle; relative to the max-norm of the batch",
             "# %-8s %-8s %5s %5s %6s %7s %10s %10s %10s %10s %10s %8s %s" % ("model", "gait", "cand", "n_t", "ok", "flagged", "cot", "work", "x", "z", "tau", "ranking", "oracle_s")]
    for model, n_t in (("hexapod", 64), ("hexapod", 256), ("myant", 48), ("spider", 64)):
        for gait in ("straight", "curved"):
            n = a.candidates if n_t <= 64 else a.candidates // 4
            p = _random_candidates(model, n, 20261018 + n_t)
            if gait == "curved":
                rng = np.random.default_rng(5)
                p[:, 10] = rng.uniform(-0.1, 0.1, n)
                if model != "spider":
                    p[:, 11] = 0; p[:, 12] = rng.uniform(0, 0.3, n)
            xml = hsl.model_path(model)
            t0 = time.time()
            ref = orc.Model(xml).eval_batch(p, n_t, nthreads=threads)
            dt = time.time() - t0
            m = hsl.Model(xml)
            gpu = m.eval_gaits(p, n_t)
            ok = ref["status"] == 0
            assert np.array_equal((gpu["status"] & 3) == 0, ok), "status mismatch"
            # not compared, but counted: frames with fewer than two feet on the ground or a degenerate contact block
            # (outside the reference's domain, its solver asserts there) and ill-conditioned candidates
            # (HSL_ST_ILLCOND: the reference's FP64 result depends on its threshold retry loop)
            flagged = int((ok & ((gpu["status"] & 28) != 0)).sum())
            ok = ok & (gpu["status"] == 0)
            rel = lambda g, r: float(np.abs(g - r).max() / np.abs(r).max())
            e_cot, e_work = rel(gpu["cot"][ok], ref["cot"][ok]), rel(gpu["work"][ok], ref["work"][ok])
            # ranking: identical outside ties closer than 1e-9 relative
            ro, go = np.argsort(ref["cot"][ok], kind="stable"), np.argsort(gpu["cot"][ok], kind="stable")
            same = np.array_equal(ro, go)
            if not same:
                c = ref["cot"][ok]
                bad = [i for i in range(len(ro)) if ro[i] != go[i] and abs(c[ro[i]] - c[go[i]]) > 1e-9 * abs(c[ro[i]])]
                same = not bad
            # detail entry on a few candidates
            idx = np.where(ok)[0][:6]
            ex = ez = et = 0.0
            om = orc.Model(xml)
            det = m.eval_gaits_detail(p[idx], n_t)
            for j, c in enumerate(idx):
                r = om.measure_cot(p[c], n_t, detail=True)
                ex = max(ex, rel(det["x"][j], r["x"])); ez = max(ez, rel(det["z"][j], r["z"])); et = max(et, rel(det["tau"][j], r["tau"]))
            lines.append("  %-8s %-8s %5d %5d %6d %7d %10.2e %10.2e %10.2e %10.2e %10.2e %8s %.1f" % (
                model, gait, n, n_t, int(ok.sum()), flagged, e_cot, e_work, ex, ez, et, "same" if same else "DIFFERS", dt))
            print(lines[-1], flush=True)
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
    open(a.out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
