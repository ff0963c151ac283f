"""Gait search around the batched evaluator: the reference's 1-D `pgssweeper` sweep (pergen.cpp:417-449,
player.cpp:311-321) generalised to N-D grids, random search and a cross-entropy loop over
(step_duration, period, step_length, step_height), with the candidates evaluated on the GPU(s).

Candidates live on the device between generation, evaluation and selection (torch tensors are the plumbing;
the evaluation is `hsl_eval_gaits` through the C ABI).  With torch.distributed initialised every rank evaluates
its shard and the costs are all-gathered (NCCL) before selection, so all ranks agree on the ranking.
"""
import itertools

import numpy as np

from .api import HSL_NPARAM, SWEEP_NAMES

BOUNDS = {"step_duration": (0.0, 1.0), "period": (0.3, 20.0), "step_length": (1e-6, 1.0), "step_height": (1e-6, 0.5)}


def grid_candidates(base, axes):
    """Cartesian product of per-parameter value lists.  axes = {name: values}; for one axis built as
    val0 + i*(val1-val0)/n_val, i = 0..n_val, this is exactly pgssweeper::sweep/next."""
    names = list(axes)
    for n in names:
        if n not in SWEEP_NAMES:
            raise ValueError("ERROR: cannot sweep over " + n)
    combos = list(itertools.product(*[np.asarray(axes[n], np.float64) for n in names]))
    cand = np.tile(np.asarray(base, np.float64), (len(combos), 1))
    for j, n in enumerate(names):
        cand[:, SWEEP_NAMES[n]] = [c[j] for c in combos]
    return cand


class _DeviceArray:
    """A device address as a CUDA-array-interface object (so that torch can wrap memory owned by the C library)."""

    def __init__(self, ptr, n, typestr):
        self.__cuda_array_interface__ = dict(shape=(n,), typestr=typestr, data=(ptr, False), version=3)


class DeviceEvaluator:
    """Evaluates candidate rows resident on the device; shards over ranks when torch.distributed is initialised.
    gather="peer": costs and status are all-gathered by the evaluation's own finish kernel over NVLink peer memory
    (hsl_eval_gaits_gather); gather="nccl": two torch.distributed all-gathers after it; gather="auto" (default): peer memory,
    or NCCL when the ranks cannot map each other's buffers (decided collectively, once per batch size)."""

    def __init__(self, model, n_t, flags=0, gather="auto"):
        import torch
        self.torch, self.model, self.n_t, self.flags = torch, model, n_t, flags
        self.dist = torch.distributed if (torch.distributed.is_available() and torch.distributed.is_initialized()) else None
        self.world = self.dist.get_world_size() if self.dist else 1
        self.rank = self.dist.get_rank() if self.dist else 0
        if gather not in ("peer", "nccl", "auto"):
            raise ValueError("gather must be 'auto', 'peer' or 'nccl'")
        self.gather = gather
        self._gathers = {}   # candidates per rank -> api.Gather

    def close(self):
        """Frees the peer gather buffers (collective: every rank calls it)."""
        if self._gathers:
            self.torch.cuda.synchronize()
            self.dist.barrier()
            for g in self._gathers.values():
                if g is not None:
                    g.check()
                    g.free()
            self._gathers = {}

    def __call__(self, cand):
        """cand: float64 CUDA tensor [C][13] (identical on every rank).  Returns (cot[C], status[C]) CUDA tensors."""
        torch = self.torch
        c = cand.shape[0]
        per = -(-c // self.world)
        lo, hi = min(self.rank * per, c), min(self.rank * per + per, c)
        g = None
        if self.dist and self.gather != "nccl":
            from . import api
            if per not in self._gathers:
                self._gathers[per] = api.torch_gather(self.dist, per, fallback=(self.gather == "auto"))
            g = self._gathers[per]
        if g is not None:
            local = cand[lo:hi].contiguous() if hi > lo else None
            p_cot, p_st = self.model.eval_gaits_gather(g, hi - lo, self.n_t, local.data_ptr() if hi > lo else 0, flags=self.flags,
                                                       stream=torch.cuda.current_stream().cuda_stream)
            n = self.world * per   # the views live in the gather buffer and are overwritten two calls later: hand out copies
            cot = torch.as_tensor(_DeviceArray(p_cot, n, "<f8"), device=cand.device).clone()
            st = torch.as_tensor(_DeviceArray(p_st, n, "<i4"), device=cand.device).clone()
            return cot[:c], st[:c]
        cot = torch.full((per,), float("nan"), dtype=torch.float64, device=cand.device)
        st = torch.zeros(per, dtype=torch.int32, device=cand.device)
        if hi > lo:
            local = cand[lo:hi].contiguous()
            self.model.eval_gaits_device(hi - lo, self.n_t, local.data_ptr(), cot.data_ptr(), 0, 0, 0, st.data_ptr(), self.flags,
                                         torch.cuda.current_stream().cuda_stream)
        if self.dist:
            allc = torch.empty(self.world * per, dtype=torch.float64, device=cand.device)
            alls = torch.empty(self.world * per, dtype=torch.int32, device=cand.device)
            self.dist.all_gather_into_tensor(allc, cot)
            self.dist.all_gather_into_tensor(alls, st)
            cot, st = allc, alls
        return cot[:c], st[:c]


def top_k(cot, k, trim=True):
    """Indices and costs of the k cheapest candidates in stable ascending order (`hsl_select_topk` on the device: a
    multi-block bitonic sort of (cost, index) pairs).  Failed candidates (NaN cost) are never selected: the entries past
    the number of valid candidates are -1 / NaN.  trim=True cuts them off (one host synchronisation, for interactive
    use); trim=False returns the padded length-k tensors without touching the host."""
    import torch
    from . import api
    cot = cot.contiguous()
    k = int(min(k, cot.numel()))
    idx = torch.empty(k, dtype=torch.int64, device=cot.device)
    val = torch.empty(k, dtype=torch.float64, device=cot.device)
    api.select_topk_device(cot.data_ptr(), cot.numel(), k, idx.data_ptr(), val.data_ptr(), torch.cuda.current_stream().cuda_stream)
    if not trim:
        return idx, val
    n_ok = int((idx >= 0).sum())
    return idx[:n_ok], val[:n_ok]


def grid_search(model, base, axes, n_t, k=1, flags=0, device="cuda"):
    import torch
    cand = torch.from_numpy(grid_candidates(base, axes)).to(device)
    ev = DeviceEvaluator(model, n_t, flags)
    cot, st = ev(cand)
    ev.close()
    idx, best = top_k(cot, k)
    return dict(candidates=cand, cot=cot, status=st, best_index=idx, best_cot=best)


def random_search(model, base, ranges, n_cand, n_t, seed=0, k=1, flags=0, device="cuda"):
    """Uniform sampling of the named parameters inside `ranges` = {name: (lo, hi)}; generated on the device."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    cand = torch.from_numpy(np.asarray(base, np.float64)).to(device).repeat(n_cand, 1)
    for name, (lo, hi) in ranges.items():
        cand[:, SWEEP_NAMES[name]] = lo + (hi - lo) * torch.rand(n_cand, dtype=torch.float64, device=device, generator=g)
    ev = DeviceEvaluator(model, n_t, flags)
    cot, st = ev(cand)
    ev.close()
    idx, best = top_k(cot, k)
    return dict(candidates=cand, cot=cot, status=st, best_index=idx, best_cot=best)


def cem_search(model, base, ranges, n_t, pop=1024, elite=64, iters=8, seed=0, flags=0, device="cuda", graph="auto"):
    """Cross-entropy method over the named parameters: sample a Gaussian population, evaluate, refit to the elite.
    The loop never synchronises with the host: population, costs, elite set, the Gaussian's moments and the running best
    stay on the device (an iteration without a single valid candidate leaves the distribution unchanged); results are
    read back once at the end.
    graph: an iteration is ~40 small launches (sampling, the three evaluation kernels, the top-k sort, the refit), i.e.
    launch bound for populations of a few thousand candidates (measured, 2048 candidates x 32 frames: 401 us per iteration);
    with graph=True the second iteration is captured into a CUDA graph and the remaining ones are replays of it (213 us per
    iteration, same random stream and results; capture and instantiation cost ~10 ms once, so "auto" switches it on from 64
    iterations, on a single rank).  Multi-rank searches stay eager: the peer-memory gather numbers its calls on the host."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    names = list(ranges)
    lo = torch.tensor([ranges[n][0] for n in names], dtype=torch.float64, device=device)
    hi = torch.tensor([ranges[n][1] for n in names], dtype=torch.float64, device=device)
    mean, std = ((lo + hi) / 2).clone(), ((hi - lo) / 2).clone()
    cols = torch.tensor([SWEEP_NAMES[n] for n in names], dtype=torch.int64, device=device)
    base_t = torch.from_numpy(np.asarray(base, np.float64)).to(device)
    ev = DeviceEvaluator(model, n_t, flags)
    inf = torch.tensor(float("inf"), dtype=torch.float64, device=device)
    best_cot = inf.clone()
    best_cand = torch.full((HSL_NPARAM,), float("nan"), dtype=torch.float64, device=device)
    history = torch.full((max(iters, 1),), float("inf"), dtype=torch.float64, device=device)
    it = torch.zeros(1, dtype=torch.int64, device=device)

    def iteration():   # everything in place: the same function runs eagerly and under capture
        z = torch.randn(pop, len(names), dtype=torch.float64, device=device, generator=g)
        x = torch.minimum(torch.maximum(mean + std * z, lo), hi)
        cand = base_t.repeat(pop, 1)
        cand.index_copy_(1, cols, x)
        cot, st = ev(cand)
        idx, c = top_k(cot, elite, trim=False)          # padded with -1 / NaN past the valid candidates
        w = (idx >= 0).to(torch.float64)                  # elite membership weights
        cnt = w.sum()
        el = x.index_select(0, idx.clamp(min=0))
        m1 = (w[:, None] * el).sum(0) / cnt.clamp(min=1.0)
        var = (w[:, None] * (el - m1) ** 2).sum(0) / cnt.clamp(min=1.0)
        any_ok = cnt > 0
        mean.copy_(torch.where(any_ok, m1, mean))
        std.copy_(torch.where(any_ok, var.sqrt() + 1e-6 * (hi - lo), std))
        cbest = torch.where(any_ok, c[0], inf)
        history.index_copy_(0, it, cbest.reshape(1))
        it.add_(1)
        improved = cbest < best_cot
        first = cand.index_select(0, idx[:1].clamp(min=0))[0]   # (indexing with a 0-dim tensor would read it back on the host)
        best_cand.copy_(torch.where(improved, first, best_cand))
        best_cot.copy_(torch.where(improved, cbest, best_cot))

    use_graph = (graph is True) or (graph == "auto" and ev.world == 1 and iters >= 64)
    if use_graph and ev.world > 1:
        raise ValueError("cem_search(graph=True) needs a single rank")
    done = 0
    if use_graph:
        iteration()                                       # warm-up: grows every workspace the capture must not allocate
        done = 1
        cg = torch.cuda.CUDAGraph()
        cg.register_generator_state(g)
        with torch.cuda.graph(cg):
            iteration()
        for _ in range(iters - done):                     # capture does not execute: every remaining iteration is a replay
            cg.replay()
    else:
        for _ in range(iters):
            iteration()
    hist = [float(v) for v in history[:iters].cpu()] if iters else []   # the one read-back
    hist = [v for v in hist if np.isfinite(v)]
    bc = float(best_cot)
    ev.close()
    return dict(best_cot=bc, best_candidate=best_cand if np.isfinite(bc) else None, history=hist, mean=mean, std=std)
