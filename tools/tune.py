#!/usr/bin/env python3
"""Times the cost-only gait evaluation for each kernel variant (frame slots per block x register cap).
Usage: python tools/tune.py [--model hexapod] [--candidates 4096] [--frames 256]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="hexapod")
    ap.add_argument("--candidates", type=int, default=4096)
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--curved", action="store_true", help="BASELINE config 2, second batch: curvature U[-.1,.1], lateral foot shift U[0,.3]")
    ap.add_argument("--variants", default="32:128,64:128,64:1,32:1")
    args = ap.parse_args()
    import torch
    import hslabs_b200 as hsl
    from bench import synth_candidates
    m = hsl.Model(hsl.model_path(args.model))
    p = synth_candidates(2 * args.candidates, 20261018)
    if args.model == "spider":
        p[:, 2] = np.random.default_rng(1).uniform(0, 0.1, p.shape[0])
        p[:, 11] = 0
        p[:, 12] = np.random.default_rng(2).uniform(0.3, 0.5, p.shape[0])
    if args.model == "myant":
        p[:, 2] = -0.07
    if args.curved:
        rc = np.random.default_rng(3)
        p[:, 10] = rc.uniform(-0.1, 0.1, p.shape[0]); p[:, 11] = 0; p[:, 12] = rc.uniform(0, 0.3, p.shape[0])
    st = m.eval_gaits(p, 20)["status"]
    p = np.ascontiguousarray(p[st == 0][:args.candidates])
    n = p.shape[0]
    dp = torch.from_numpy(p).cuda()
    cot = torch.empty(n, dtype=torch.float64, device="cuda")
    stt = torch.empty(n, dtype=torch.int32, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device="cuda")
    for var in args.variants.split(","):
        fb, mr = (int(v) for v in var.split(":"))
        m.set_tuning(fb, mr)
        ts = []
        for r in range(args.reps + 3):
            flush.fill_(r)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            m.eval_gaits_device(n, args.frames, dp.data_ptr(), cot.data_ptr(), 0, 0, 0, stt.data_ptr(), 0, stream)
            b.record()
            torch.cuda.synchronize()
            if r >= 3:
                ts.append(a.elapsed_time(b))
        ms = float(np.median(ts))
        print("%s fb=%d maxreg=%d  %.3f ms  %.3e frame solves/s  (cot checksum %.12g, bad %d)" % (
            args.model, fb, mr, ms, n * args.frames / (ms * 1e-3), float(torch.nan_to_num(cot).sum()), int((stt != 0).sum())))


if __name__ == "__main__":
    main()
