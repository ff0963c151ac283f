#!/usr/bin/env python3
"""bench.py -- frame solves/sec (IK + inverse dynamics + contact solve) of the gait-evaluation hot path.

A "step" is one pass of the hot path over one batch of synthetic candidate gaits.  Headline workload: BASELINE.json
configs[1], hexapod.xml, 4096 candidates x 256 frames per GPU (weak scaling: every rank evaluates its own
4096-candidate shard, the per-candidate costs are all-gathered over NCCL and the cheapest candidate is selected on
every rank).  The same JSON line carries sub-records:

  sustained   the headline step looped for >= 2 s with the SM clock sampled (the 15 ms timed region is a burst)
  config3     BASELINE.json configs[2]: spider.xml, 65536 candidates x 512 frames strong-sharded over the N ranks
              (at N > 1 rank 0 also times the whole configuration alone, for the in-run scaling efficiency)
  parity      after the timed loops: every status is 0 and 64 sampled costs equal the CPU reference arm to 1e-9
  cpu_baseline (N = 1)  the reference's own CPU code (oracle/_ref) or, without it, the oracle port, all cores + 1 thread

    python bench.py --gpus N --steps K --warmup W            # N>1: launched by torch.distributed.run
    python bench.py --workload config3 ...                   # config 3 as the headline (strong scaling)
    python bench.py --impl reference ...                     # the reference's CPU implementation of the path

Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 20261018
# algorithmic flops per frame solve = base + per_contact * (mean feet on the ground)
WORKLOADS = {
    "config2": dict(model="hexapod", n_cand=4096, n_t=256, flops=(6920.0, 596.0), scaling="weak",
                    name="hexapod.xml, %d candidates x %d frames per GPU (BASELINE configs[1])"),
    "config3": dict(model="spider", n_cand=65536, n_t=512, flops=(6440.0, 578.0), scaling="strong",
                    name="spider.xml, %d candidates x %d frames sharded over the GPUs (BASELINE configs[2])"),
}
METRIC = "frame solves/sec (IK+ID+contact QP)"


def synth_candidates(n, seed, rng=None, model="hexapod"):
    """step_duration U[0,1], period U[1,6], step_length U[.1,.5], step_height U[.02,.12],
    straight walking; hexapod: torso z U[-.15,-.05], no foot shift; spider: torso z U[0,.1], lateral foot shift U[.3,.5]."""
    rng = rng or np.random.default_rng(seed)
    p = np.zeros((n, 13))
    p[:, 2] = rng.uniform(-0.15, -0.05, n) if model != "spider" else rng.uniform(0.0, 0.1, n)
    p[:, 6] = rng.uniform(0.0, 1.0, n)
    p[:, 7] = rng.uniform(1.0, 6.0, n)
    p[:, 8] = rng.uniform(0.1, 0.5, n)
    p[:, 9] = rng.uniform(0.02, 0.12, n)
    p[:, 11] = -1
    if model == "spider":
        p[:, 11] = 0
        p[:, 12] = rng.uniform(0.3, 0.5, n)
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
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_min_mhz": min(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "power_w_max": max(pw) if pw else None, "reasons": sorted(reasons),
                "samples": len(rows)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_model(model):
    """The reference's own CPU code when its build travelled here (oracle/_ref, kind "reference"), else the oracle port
    (kind "port": dense QR standing in for Eigen's sparse QR in both)."""
    xml = os.path.join(ROOT, "hslabs_b200", "models", model + ".xml")
    from oracle import ref
    if ref.available():
        return ref.Model(xml), "reference"
    from oracle import orc
    return orc.Model(xml), "port"


def cpu_rate(m, params, n_t, workers, budget_s=15.0, max_rounds=128):
    """Times m.eval_batch on a bounded sample: `workers` worker processes (reference build) or threads (port)."""
    t0 = time.perf_counter()
    m.eval_batch(params[:workers], n_t, workers)           # one candidate per worker: calibrates the sample size
    per_round = max(time.perf_counter() - t0, 1e-3)
    rounds = int(max(1, min(max_rounds, budget_s / per_round - 1, params.shape[0] // workers)))
    n = workers * rounds
    t0 = time.perf_counter()
    out = m.eval_batch(params[:n], n_t, workers)
    dt = time.perf_counter() - t0
    return n * n_t / dt, n, dt, int((out["status"] == 0).sum())


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
    """--impl reference: the reference's CPU implementation of the path on all host cores (worker processes: the reference
    is single-threaded and not re-entrant), a bounded sample of the workload per step.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    w = WORKLOADS[args.workload]
    n_t = args.frames or w["n_t"]
    cores = os.cpu_count() or 1
    m, kind = cpu_model(w["model"])
    params = synth_candidates(max(256, cores * 64), SEED, model=w["model"])
    rates, sample, n = [], "", 0
    for step in range(args.warmup + args.steps):
        rate, n, dt, ok = cpu_rate(m, params, n_t, cores, budget_s=10.0, max_rounds=64)   # ~10 s of all-core CPU work per step
        if step >= args.warmup:
            rates.append(rate)
        sample = "%d candidates x %d frames per step on %d %s (%.1f s)" % (n, n_t, cores, "worker processes" if kind == "reference" else "threads", dt)
    v = float(np.mean(rates))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "frame solves/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * (n * n_t) / v, "higher_is_better": True,
            "scaling": w["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": (w["name"] % (w["n_cand"], n_t)) + "; CPU arm runs a bounded sample"},
            "cpu_baseline": {"value": v, "unit": "frame solves/s", "cores": cores, "kind": kind, "sample": sample,
                             "note": "reference sources compiled unmodified against shim headers (oracle/_ref); Eigen's sparse QR is a Householder QR without column reordering behind the facade (zero rows of the reflectors skipped)"
                             if kind == "reference" else "oracle port of the reference path (oracle/_ref not present)"},
            "e2e": {"value": v, "unit": "frame solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--workload", default="config2", choices=sorted(WORKLOADS))
    ap.add_argument("--candidates", type=int, default=0, help="candidates per GPU (config2) / in total (config3); 0 = the workload's")
    ap.add_argument("--frames", type=int, default=0)
    ap.add_argument("--fb", type=int, default=0, help="frame slots per block (32/64), 0 = library default")
    ap.add_argument("--maxreg", type=int, default=0, help="register cap of the cost-only kernel variant (1 = pipelined), 0 = library default")
    ap.add_argument("--gather", default="auto", choices=["auto", "peer", "nccl", "none"], help="N > 1: costs all-gathered by the finish kernel over NVLink peer memory (hsl_eval_gaits_gather) or by NCCL after it; none = diagnostic only (no collective: shows the ranks' own step times), its line is not a bench result")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline and parity legs")
    ap.add_argument("--no-extras", action="store_true", help="skip the sustained and config3 sub-records")
    args = ap.parse_args()
    _claim_stdout()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import hslabs_b200 as hsl
    from hslabs_b200 import api
    from hslabs_b200.sharding import shard_bounds

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        import datetime
        # a short watchdog: a mismatched collective must fail in two minutes, not hold N GPUs for the default ten
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    stream = torch.cuda.current_stream().cuda_stream
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # 256 MiB > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reachable_candidates(model, name, n, rng):
        """Unreachable draws are redrawn; reachability does not depend on n_t beyond round-off, the
        status of the timed evaluation is checked again afterwards."""
        out = np.zeros((0, 13))
        while out.shape[0] < n:
            p = synth_candidates(n + n // 4 + 16, SEED, rng, model=name)
            st = model.eval_gaits(p, 20)["status"]
            out = np.concatenate([out, p[st == 0]])
        return np.ascontiguousarray(out[:n])

    class Job:
        """One workload on this rank: local shard resident in HBM, step() = evaluate + all-gather + select."""

        def __init__(self, wname, n_total=None, ranks=world, my_rank=rank, n_t=None):
            w = WORKLOADS[wname]
            self.w, self.n_t = w, n_t or w["n_t"]
            self.model = hsl.Model(hsl.model_path(w["model"]))
            if args.fb or args.maxreg:
                self.model.set_tuning(args.fb or None, args.maxreg or None)
            self.ranks = ranks
            if w["scaling"] == "weak":
                self.n_local = n_total or w["n_cand"]
                self.n_total = self.n_local * ranks
                rng = np.random.default_rng(SEED + my_rank)
                self.params = reachable_candidates(self.model, w["model"], self.n_local, rng)
                self.per = self.n_local
            else:
                self.n_total = n_total or w["n_cand"]
                allp = reachable_candidates(self.model, w["model"], self.n_total, np.random.default_rng(SEED))  # same on every rank
                lo, hi, self.per = shard_bounds(self.n_total, ranks, my_rank)
                self.params = np.ascontiguousarray(allp[lo:hi])
                self.n_local = hi - lo
            self.d_params = torch.from_numpy(self.params).to(dev)
            self.d_cot = torch.full((self.per,), float("nan"), dtype=torch.float64, device=dev)
            self.d_work = torch.empty(self.per, dtype=torch.float64, device=dev)
            self.d_status = torch.zeros(self.per, dtype=torch.int32, device=dev)
            self.gathered = torch.empty(ranks * self.per, dtype=torch.float64, device=dev) if ranks > 1 else None
            self.peer = None
            if ranks > 1 and args.gather in ("auto", "peer"):   # collective: every rank builds the same jobs
                self.peer = api.torch_gather(dist, self.per, fallback=(args.gather == "auto"))
                if self.peer is None and my_rank == 0:
                    sys.stderr.write("bench.py: the ranks cannot map each other's gather buffers; costs go through NCCL\n")
            self.best = torch.empty(1, dtype=torch.int64, device=dev)

        def step(self):
            if self.peer is not None:   # evaluation + all-gather in the same launches; selection on every rank
                self.model.eval_gaits_scatter(self.peer, self.n_local, self.n_t, self.d_params.data_ptr() if self.n_local else 0,
                                              self.d_cot.data_ptr(), self.d_work.data_ptr(), 0, 0, self.d_status.data_ptr(), 0, stream)
                self.peer.select_best(self.best.data_ptr(), 0, stream)   # its first instructions wait for the ranks' flags
                return self.best
            if self.n_local:
                self.model.eval_gaits_device(self.n_local, self.n_t, self.d_params.data_ptr(), self.d_cot.data_ptr(), self.d_work.data_ptr(),
                                             0, 0, self.d_status.data_ptr(), 0, stream)
            costs = self.d_cot
            if self.ranks > 1 and args.gather != "none":   # NCCL form
                dist.all_gather_into_tensor(self.gathered, self.d_cot)
                costs = self.gathered
            api.select_best_device(costs.data_ptr(), costs.numel(), self.best.data_ptr(), 0, stream)   # selection on every rank
            return self.best

        def timed(self, steps, warmup, sample_clocks=True, sync_ranks=True, time_kernels=False):
            for _ in range(max(warmup, 3)):
                flush.fill_(1.0)
                self.step()
            if sync_ranks:
                barrier()
            else:
                torch.cuda.synchronize()
            sampler = ClockSampler(local) if sample_clocks else None
            if sampler:
                sampler.start()
            self.model.set_kernel_timing(time_kernels)   # off for the headline: no extra event records in its timed region
            l0 = self.model.launch_count()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            kms = []
            if sync_ranks:
                barrier()
            for a, b in ev:
                flush.fill_(float(rank))      # L2 flush between timed iterations (outside the event pair)
                a.record()
                self.step()
                b.record()
            if sync_ranks:
                barrier()
            else:
                torch.cuda.synchronize()
            clocks = sampler.stop() if sampler else None
            step_ms = [a.elapsed_time(b) for a, b in ev]
            if self.n_local and time_kernels:
                kms = list(self.model.last_kernel_ms())    # the last step's three kernels (events on the launching stream)
            self.model.set_kernel_timing(False)
            launches = self.model.launch_count() - l0 + steps  # setup + frames + finish per chunk, + the selection kernel
            return step_ms, clocks, launches, kms

    # ------------------------------------------------------------------ headline
    head = Job(args.workload, n_total=args.candidates or None, n_t=args.frames or None)
    n_t = head.n_t
    step_ms, clocks, launches, _ = head.timed(args.steps, args.warmup)
    if world > 1 and os.environ.get("HSL_BENCH_RANK_TIMES"):
        sys.stderr.write("[rank %d] own ms per step: mean %.4f min %.4f max %.4f\n" % (rank, float(np.mean(step_ms)), min(step_ms), max(step_ms)))
    if head.peer is not None:
        head.peer.check()      # no wait of the peer-memory gather timed out
    ms_total = max_over_ranks(sum(step_ms))
    frames_per_step = head.n_total * n_t
    value = frames_per_step * args.steps / (ms_total * 1e-3)

    # kernel-only time of the dominant kernel for the roofline: CUDA events around the per-frame kernel itself
    # (hsl_set_kernel_timing; setup and finish kernels are timed separately), one launch per step, L2 flushed before each
    head.model.set_kernel_timing(True)
    k_all = []
    for _ in range(args.steps):
        flush.fill_(2.0)
        head.model.eval_gaits_device(head.n_local, n_t, head.d_params.data_ptr(), head.d_cot.data_ptr(), head.d_work.data_ptr(), 0, 0,
                                     head.d_status.data_ptr(), 0, stream)
        k_all.append(head.model.last_kernel_ms())
    head.model.set_kernel_timing(False)
    k_setup, k_ms, k_finish = (float(np.mean([k[i] for k in k_all])) for i in range(3))

    # end to end through the host-buffer C-ABI call (H2D of the candidates + D2H of the costs inside the timed region)
    e2e_t = []
    pin_params = hsl.pinned_empty(head.params.shape)          # the step's inputs in pinned host memory, as the contract asks
    pin_params[...] = head.params
    pin_out = dict(cot=hsl.pinned_empty((head.n_local,)), work=hsl.pinned_empty((head.n_local,)), min_cfz=hsl.pinned_empty((head.n_local,)),
                   max_mu=hsl.pinned_empty((head.n_local,)), status=hsl.pinned_empty((head.n_local,), np.int32))
    for i in range(max(3, min(args.steps, 10))):
        t0 = time.perf_counter()
        out = head.model.eval_gaits(pin_params, n_t, out=pin_out)
        e2e_t.append(time.perf_counter() - t0)
    if not np.array_equal(out["cot"], head.d_cot[:head.n_local].cpu().numpy()):
        raise SystemExit("bench.py: the host-buffer call and the device-pointer call disagree on the costs")
    e2e_local = head.n_local * n_t / float(np.median(e2e_t[1:] if len(e2e_t) > 1 else e2e_t))
    if world > 1:
        t = torch.tensor([e2e_local], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN if head.w["scaling"] == "weak" else dist.ReduceOp.SUM)
        e2e_value = float(t.item()) * (world if head.w["scaling"] == "weak" else 1)
    else:
        e2e_value = e2e_local

    # every candidate of the timed batch evaluated cleanly (a NaN batch must not print a rate)
    st_host = head.d_status[:head.n_local].cpu().numpy()
    cot_host = head.d_cot[:head.n_local].cpu().numpy()
    bad = int((st_host != 0).sum()) + int((~np.isfinite(cot_host)).sum())
    badt = max_over_ranks(float(bad))
    if badt:
        raise SystemExit("bench.py: %d candidates of the timed batch have a non-zero status or a non-finite cost" % int(badt))

    line = None
    if rank == 0:
        det = head.model.eval_gaits_detail(head.params[:64], n_t)   # mean feet on the ground, for the flop count
        kbar = float(det["contacts"].sum(axis=2).mean())
        flops_per_frame = head.w["flops"][0] + head.w["flops"][1] * kbar
        dfma_tf, _ = api.dfma_probe()
        achieved_tf = (head.n_local * n_t / (k_ms * 1e-3)) * flops_per_frame / 1e12
        peaks, traffic = {}, {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        try:   # dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, from the committed ncu --set full capture
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.workload, {})
        except Exception:
            pass
        default_variant = not (args.fb or args.maxreg or args.candidates or args.frames)
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        alg_bytes = head.n_local * (13 * 8 + 4 * 8 + 4)  # cost-only mode: candidates in, costs out
        fb, mr = head.model.get_tuning()
        line = {
            "metric": METRIC, "value": value, "unit": "frame solves/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": head.w["scaling"],
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": head.w["name"] % (head.w["n_cand"] if not args.candidates else args.candidates, n_t),
                       "seed": SEED, "n_t": n_t, "candidates_per_gpu": head.per, "candidates_total": head.n_total,
                       "l2": "flushed between timed iterations (256 MiB fill)",
                       "parallelism": ("candidates sharded, costs all-gathered (%s), argmin on every rank" % ("by the finish kernel over NVLink peer memory" if head.peer is not None else "NCCL")) if world > 1 else "single GPU",
                       "mean_contacts": kbar, "kernel_variant": {"frame_slots_per_block": fb, "maxreg": mr,
                                                                 "kernel": "hsl_gait_pipe_kernel" if mr == 1 else "hsl_frames_kernel"}},
            "gpu_launches": launches,
            "ms_per_step_best": float(min(step_ms)), "ms_per_step_median": float(np.median(step_ms)),
            "e2e": {"value": e2e_value, "unit": "frame solves/s", "h2d_bytes_per_step": int(head.params.nbytes),
                    "d2h_bytes_per_step": int(head.n_local * (4 * 8 + 4))},
            "roofline": {"bound": "fp64", "achieved": achieved_tf, "peak": dfma_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / dfma_tf if dfma_tf else None,
                         "traffic": traffic.get("dram_bytes_per_launch") if default_variant else None,
                         "traffic_source": traffic.get("source") if default_variant else None,
                         "peak_source": "measured live: register-resident DFMA probe kernel (MEASURED_PEAKS.json has no FP64 figure)",
                         "flops_per_frame": flops_per_frame, "kernel_ms": k_ms,
                         "kernel_ms_note": "CUDA events around the per-frame kernel alone; candidate setup %.4f ms and per-candidate finish %.4f ms are separate launches" % (k_setup, k_finish),
                         "hbm": {"achieved_gbs": alg_bytes / (k_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                                 "note": "cost-only mode moves ~%d B per candidate; HBM is not the bound" % (13 * 8 + 36)}},
            "clocks": clocks,
        }

    # ------------------------------------------------------------------ sustained: the same step for >= 2 s
    if not args.no_extras:
        # the same count on every rank (each step has a collective): derived from the max-over-ranks time, not a local one
        n_long = int(max(200, 2200.0 / max(ms_total / args.steps, 1e-3)))
        sampler = ClockSampler(local)
        barrier()
        sampler.start()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n_long):
            head.step()
        b.record()
        barrier()
        sclk = sampler.stop()
        s_ms = max_over_ranks(a.elapsed_time(b))
        if rank == 0:
            line["sustained"] = {"value": frames_per_step * n_long / (s_ms * 1e-3), "unit": "frame solves/s", "steps": n_long,
                                 "seconds": s_ms * 1e-3, "ms_per_step": s_ms / n_long, "l2": "not flushed (back-to-back steps)",
                                 "clocks": sclk}

    # ------------------------------------------------------------------ config 3: spider 65536 x 512, strong-sharded
    if not args.no_extras and args.workload != "config3":
        c3 = Job("config3")
        c3_ms, _, _, c3_k = c3.timed(5, 3, sample_clocks=False, time_kernels=True)
        c3_total = max_over_ranks(sum(c3_ms))
        c3_frames = c3.n_total * c3.n_t
        rec = {"workload": c3.w["name"] % (c3.n_total, c3.n_t), "value": c3_frames * 5 / (c3_total * 1e-3), "unit": "frame solves/s",
               "ms_per_step": c3_total / 5, "steps": 5, "candidates_per_gpu": c3.per, "scaling": "strong",
               "kernel_ms": c3_k[1] if c3_k else None}
        st3 = c3.d_status[:c3.n_local].cpu().numpy()
        c3_bad = max_over_ranks(float((st3 != 0).sum()))
        rec["status_all_zero"] = (c3_bad == 0)
        if world > 1:
            # the whole configuration on rank 0's GPU alone, same run, for the scaling efficiency (other ranks wait)
            t1 = 0.0
            if rank == 0:
                solo = Job("config3", ranks=1, my_rank=0)
                solo_ms, _, _, _ = solo.timed(3, 3, sample_clocks=False, sync_ranks=False)
                t1 = float(np.mean(solo_ms))
                del solo
            barrier()
            if rank == 0:
                rec["single_gpu_ms_per_step"] = t1
                rec["efficiency_vs_single_gpu"] = t1 / (world * (c3_total / 5))
        if rank == 0:
            line["config3"] = rec
        del c3

    # ------------------------------------------------------------------ fall / perturbation sweep (BASELINE configs[4])
    if not args.no_extras:
        fm = hsl.Model(hsl.model_path("hexapod"))
        fparams, _ = hsl.load_preset(os.path.join(ROOT, "hslabs_b200", "models", "pgs_presets.txt"), 8)
        fw, fsteps = 131072, 50                      # worlds per GPU (55 waves of 148 SMs x 16 warps, a warp per world), steps of 0.02 s
        frng = np.random.default_rng(SEED + 77 + rank)
        mag, th = frng.uniform(0, 16, fw), frng.uniform(0, 2 * np.pi, fw)
        fk = np.stack([mag * np.cos(th), mag * np.sin(th), np.zeros(fw)], axis=1)
        fks = frng.integers(5, 25, fw).astype(np.int32)
        fm.fall_sweep(fparams, 5, fks[:4096], fk[:4096])                 # warm-up
        barrier()
        fres = fm.fall_sweep(fparams, fsteps, fks, fk)
        f_ms = max_over_ranks(fres["kernel_ms"])
        if rank == 0:
            line["fall_sweep"] = {"workload": "hexapod.xml preset 8 under PD position control, %d worlds per GPU x %d steps of 0.02 s, one random torso kick each (BASELINE configs[4])" % (fw, fsteps),
                                  "value": world * fw * fsteps / (f_ms * 1e-3), "unit": "world-steps/s", "kernel_ms": f_ms,
                                  "fall_rate": float(fres["fell"].mean()), "status_all_zero": bool((fres["status"] == 0).all()),
                                  "note": "one warp per world, lane = body, the world in registers; bound by the dependent chain of a Gauss-Seidel row update (DESIGN.md section 9)"}
        del fm

    # ------------------------------------------------------------------ CPU arm + parity of the timed batch (N = 1 only)
    if rank == 0 and not args.no_cpu and world == 1:
        cores = os.cpu_count() or 1
        m, kind = cpu_model(head.w["model"])
        rate, n, dt, ok = cpu_rate(m, head.params, n_t, cores, budget_s=12.0)
        rate1, n1, dt1, _ = cpu_rate(m, head.params[::-1], n_t, 1, budget_s=4.0, max_rounds=16)
        unit = "worker processes" if kind == "reference" else "threads"
        line["cpu_baseline"] = {"value": rate, "unit": "frame solves/s", "cores": cores, "kind": kind,
                                "sample": "%d candidates x %d frames on %d %s (%.1f s)" % (n, n_t, cores, unit, dt),
                                "single_thread": {"value": rate1, "sample": "%d candidates x %d frames (%.1f s); the reference binary is single-threaded (makefile:1)" % (n1, n_t, dt1)},
                                "note": "the reference's own sources compiled unmodified against shim headers (oracle/_ref); Eigen's sparse QR is a Householder QR without column reordering behind the facade (zero rows of the reflectors skipped)"
                                if kind == "reference" else "oracle port of the reference path"}
        sel = np.linspace(0, head.n_local - 1, 64).astype(int)
        refc = m.eval_batch(head.params[sel], n_t, cores)
        good = refc["status"] == 0
        err = float(np.max(np.abs(cot_host[sel][good] - refc["cot"][good]) / np.abs(refc["cot"][good]))) if good.any() else float("nan")
        same_rank = bool(np.array_equal(np.argsort(cot_host[sel][good], kind="stable"), np.argsort(refc["cot"][good], kind="stable")))
        line["parity"] = {"parity_checked": bool(good.sum() >= 48 and err <= 1e-9 and same_rank), "candidates": int(good.sum()),
                          "max_rel_err_cot": err, "ranking_identical": same_rank, "against": kind, "tolerance": 1e-9,
                          "status_all_zero": True}
        line["parity_checked"] = line["parity"]["parity_checked"]
    if rank == 0:
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
