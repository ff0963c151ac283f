#!/usr/bin/env python3
"""bench.py -- frame solves/sec (IK + inverse dynamics + contact solve) of the gait-evaluation hot path.

A "step" is one pass of the hot path over one batch of synthetic candidate gaits: BASELINE.json configs[1],
hexapod.xml, 4096 candidates x 256 frames per GPU (weak scaling: every rank evaluates its own 4096-candidate
shard, then the per-candidate costs are all-gathered over NCCL for selection).

    python bench.py --gpus N --steps K --warmup W          # N>1: launched by torch.distributed.run
    python bench.py --impl reference ...                    # the CPU port of the reference path (oracle/)

Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20261018
N_CAND, N_T, MODEL = 4096, 256, "hexapod"
FLOPS_BASE, FLOPS_PER_CONTACT = 6920.0, 596.0  # BASELINE.md section 4, hexapod


def synth_candidates(n, seed, rng=None):
    """torso z U[-.15,-.05], step_duration U[0,1], period U[1,6], step_length U[.1,.5],
    step_height U[.02,.12], straight walking, no foot shift."""
    rng = rng or np.random.default_rng(seed)
    p = np.zeros((n, 13))
    p[:, 2] = rng.uniform(-0.15, -0.05, n)
    p[:, 6] = rng.uniform(0.0, 1.0, n)
    p[:, 7] = rng.uniform(1.0, 6.0, n)
    p[:, 8] = rng.uniform(0.1, 0.5, n)
    p[:, 9] = rng.uniform(0.02, 0.12, n)
    p[:, 11] = -1
    return p


class ClockSampler:
    """nvidia-smi clocks / throttle reasons streamed every 50 ms during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.15)
        except Exception:
            self.proc = None

    def stop(self):
        rows = []
        if self.proc is not None:
            time.sleep(0.06)
            self.proc.terminate()
            try:
                out, _ = self.proc.communicate(timeout=5)
            except Exception:
                self.proc.kill()
                out = ""
            rows = [[v.strip() for v in line.split(",")] for line in out.splitlines() if line.strip()]
        sm = [float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons), "samples": len(rows)}


def cpu_port_rate(params, n_t, nthreads, budget_s=15.0, max_rounds=8):
    """Times the CPU port of the reference path (oracle/, dense-QR stand-in for Eigen's sparse QR) on a bounded sample."""
    from oracle import orc
    m = orc.Model(os.path.join(ROOT, "hslabs_b200", "models", MODEL + ".xml"))
    t0 = time.perf_counter()
    m.eval_batch(params[:nthreads], n_t, nthreads)       # one candidate per thread: calibrates the sample size
    t1 = time.perf_counter() - t0
    per_round = max(t1, 1e-3)
    rounds = int(max(1, min(max_rounds, budget_s / per_round - 1, params.shape[0] // nthreads)))
    n = nthreads * rounds
    t0 = time.perf_counter()
    out = m.eval_batch(params[:n], n_t, nthreads)
    dt = time.perf_counter() - t0
    ok = int((out["status"] == 0).sum())
    return n * n_t / dt, n, dt, ok


_JSON_FD = None


def _claim_stdout():
    """stdout must carry the one JSON line and nothing else, but native libraries (NCCL prints its version banner at
    NCCL_DEBUG=VERSION/WARN) write to fd 1 directly: keep a private duplicate of fd 1 for the result line and point
    fd 1 at stderr for everything else."""
    global _JSON_FD
    if _JSON_FD is None:
        sys.stdout.flush()
        _JSON_FD = os.dup(1)
        os.dup2(2, 1)


def _emit(line):
    data = (json.dumps(line) + "\n").encode()
    sys.stdout.flush()
    if _JSON_FD is None:
        os.write(1, data)
    else:
        os.write(_JSON_FD, data)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    params = synth_candidates(max(64, cores * 8), SEED)
    rates = []
    sample = ""
    for step in range(args.warmup + args.steps):
        rate, n, dt, ok = cpu_port_rate(params, N_T, cores, budget_s=10.0)
        if step >= args.warmup:
            rates.append(rate)
        sample = "%d candidates x %d frames per step on %d threads (%.1f s)" % (n, N_T, cores, dt)
    v = float(np.mean(rates))
    line = {"impl": "reference", "metric": "frame solves/sec (IK+ID+contact QP)", "value": v, "unit": "frame solves/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * (n * N_T) / v, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "hexapod.xml, %d candidates x %d frames per GPU (BASELINE configs[1]); CPU arm runs a bounded sample" % (N_CAND, N_T)},
            "cpu_baseline": {"value": v, "unit": "frame solves/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "frame solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--candidates", type=int, default=N_CAND)
    ap.add_argument("--frames", type=int, default=N_T)
    ap.add_argument("--fb", type=int, default=0, help="frame slots per block (32/64/96), 0 = library default")
    ap.add_argument("--maxreg", type=int, default=128, help="register cap per thread of the cost-only kernel variant")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    _claim_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import hslabs_b200 as hsl
    from hslabs_b200 import api

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    n_cand, n_t = args.candidates, args.frames
    model = hsl.Model(hsl.model_path(MODEL))
    if args.fb:
        model.set_tuning(args.fb, args.maxreg)

    # synthetic candidates for this rank's shard; unreachable draws are redrawn
    rng = np.random.default_rng(SEED + rank)
    params = synth_candidates(2 * n_cand, SEED, rng)
    st = model.eval_gaits(params, 20)["status"]
    params = np.ascontiguousarray(params[st == 0][:n_cand])
    assert params.shape[0] == n_cand, "not enough reachable candidates"

    d_params = torch.from_numpy(params).to(dev)
    d_cot = torch.empty(n_cand, dtype=torch.float64, device=dev)
    d_work = torch.empty_like(d_cot)
    d_status = torch.empty(n_cand, dtype=torch.int32, device=dev)
    gathered = torch.empty(world * n_cand, dtype=torch.float64, device=dev) if world > 1 else None
    best = torch.empty(1, dtype=torch.int64, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # 256 MiB > 126 MB L2
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        model.eval_gaits_device(n_cand, n_t, d_params.data_ptr(), d_cot.data_ptr(), d_work.data_ptr(), 0, 0,
                                d_status.data_ptr(), 0, stream)
        costs = d_cot
        if world > 1:
            dist.all_gather_into_tensor(gathered, d_cot)
            costs = gathered
        api.select_best_device(costs.data_ptr(), costs.numel(), best.data_ptr(), 0, stream)   # selection on every rank
        return best

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        flush.fill_(1.0)
        step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = model.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for a, b in ev:
        flush.fill_(float(rank))      # L2 flush between timed iterations (outside the event pair)
        a.record()
        step()
        b.record()
    barrier()
    clocks = sampler.stop()
    launches = model.launch_count() - launches0 + args.steps  # setup + frames + finish per step, + the selection kernel
    step_ms = [a.elapsed_time(b) for a, b in ev]
    ms = sum(step_ms)
    tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms_total = float(tmax.item())
    frames_per_step = world * n_cand * n_t
    value = frames_per_step * args.steps / (ms_total * 1e-3)

    # kernel-only timing of the dominant kernel (hsl_gait_pipe_kernel, the default cost-only variant for six-limbed
    # models) for the roofline: the three launches of a step are timed together here; the setup/finish kernels are
    # < 3 % (profiles/r01_launches.txt).
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in kev:
        flush.fill_(2.0)
        a.record()
        model.eval_gaits_device(n_cand, n_t, d_params.data_ptr(), d_cot.data_ptr(), d_work.data_ptr(), 0, 0,
                                d_status.data_ptr(), 0, stream)
        b.record()
    torch.cuda.synchronize()
    k_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))

    # end to end through the host-buffer C-ABI call (H2D of the candidates + D2H of the costs inside the timed region)
    e2e_t = []
    for i in range(max(3, min(args.steps, 10))):
        t0 = time.perf_counter()
        out = model.eval_gaits(params, n_t)
        e2e_t.append(time.perf_counter() - t0)
    e2e_local = n_cand * n_t / float(np.median(e2e_t[1:] if len(e2e_t) > 1 else e2e_t))
    e2e_rate = torch.tensor([e2e_local], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_rate, op=dist.ReduceOp.MIN)
        e2e_value = float(e2e_rate.item()) * world
    else:
        e2e_value = e2e_local

    if rank == 0:
        # mean number of feet on the ground (for the algorithmic flop count), from a small detailed evaluation
        det = model.eval_gaits_detail(params[:64], n_t)
        kbar = float(det["contacts"].sum(axis=2).mean())
        flops_per_frame = FLOPS_BASE + FLOPS_PER_CONTACT * kbar
        dfma_tf, _ = api.dfma_probe()
        achieved_tf = (n_cand * n_t / (k_ms * 1e-3)) * flops_per_frame / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        alg_bytes = n_cand * (13 * 8 + 4 * 8 + 4) + n_cand * n_t * 0  # cost-only mode: candidates in, costs out
        line = {
            "metric": "frame solves/sec (IK+ID+contact QP)", "value": value, "unit": "frame solves/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "hexapod.xml, %d candidates x %d frames per GPU (BASELINE configs[1])" % (n_cand, n_t),
                       "seed": SEED, "n_t": n_t, "candidates_per_gpu": n_cand, "l2": "flushed between timed iterations (256 MiB fill)",
                       "parallelism": "candidates sharded, costs all-gathered (NCCL)" if world > 1 else "single GPU",
                       "mean_contacts": kbar},
            "gpu_launches": launches,
            # best and median of the timed steps (this rank's CUDA events), next to the mean above
            "ms_per_step_best": float(min(step_ms)), "ms_per_step_median": float(np.median(step_ms)),
            "e2e": {"value": e2e_value, "unit": "frame solves/s", "h2d_bytes_per_step": int(params.nbytes),
                    "d2h_bytes_per_step": int(n_cand * (4 * 8 + 4))},
            "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": dfma_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / dfma_tf if dfma_tf else None,
                         # dram__bytes_read.sum + dram__bytes_write.sum of hsl_gait_pipe_kernel, one ncu --set full capture of
                         # this configuration (profiles/r01_ncu_pipe_kernel_summary.txt); null for other configurations
                         "traffic": 11584000 if (n_cand, n_t, MODEL) == (4096, 256, "hexapod") and not args.fb else None,
                         "peak_source": "measured live: register-resident DFMA probe kernel (MEASURED_PEAKS.json has no FP64 figure)",
                         "flops_per_frame": flops_per_frame, "kernel_ms": k_ms,
                         "hbm": {"achieved_gbs": alg_bytes / (k_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                                 "note": "cost-only mode moves ~%d B per candidate; HBM is not the bound" % (13 * 8 + 36)}},
            "clocks": clocks,
        }
        if not args.no_cpu and world == 1:  # reported at N = 1 only (the CPU path does not change with the GPU count)
            cores = os.cpu_count() or 1
            rate, n, dt, ok = cpu_port_rate(params, n_t, cores, budget_s=15.0, max_rounds=40)  # ~10-15 s of CPU work
            line["cpu_baseline"] = {"value": rate, "unit": "frame solves/s", "cores": cores, "kind": "port",
                                    "sample": "%d candidates x %d frames on %d threads (%.1f s)" % (n, n_t, cores, dt)}
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
