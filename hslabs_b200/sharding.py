"""Multi-GPU host logic: shard candidates over ranks, evaluate locally, all-gather the costs for selection.

Candidates are independent and never span GPUs.  `torch.distributed` is the plumbing: NCCL over
NVLink on the GPU box (costs stay on the device), gloo in the CPU test tier.  The only collective is one all-gather
of 8 B per candidate; selection (argmin / ranking) runs on every rank.
"""
import numpy as np


def shard_bounds(n_cand, world, rank):
    """Contiguous blocks of ceil(C/G) candidates per rank (the last ranks may get fewer or none)."""
    per = -(-n_cand // world)
    lo = min(rank * per, n_cand)
    return lo, min(lo + per, n_cand), per


def evaluate_sharded(eval_fn, params, world, rank, all_gather_fn):
    """eval_fn(local_params) -> cost array for the local shard (NaN = failed candidate).
    all_gather_fn(padded_local_costs) -> array [world * per].  Returns (costs[C], best index or -1, ranking of the valid
    candidates)."""
    params = np.asarray(params, np.float64)
    params = params.reshape(-1, params.shape[-1])
    n = params.shape[0]
    lo, hi, per = shard_bounds(n, world, rank)
    local = np.full(per, np.nan)
    if hi > lo:
        local[:hi - lo] = np.asarray(eval_fn(params[lo:hi]), np.float64)
    gathered = np.asarray(all_gather_fn(local), np.float64).reshape(world, per)
    costs = np.concatenate([gathered[r][:max(0, min(per, n - r * per))] for r in range(world)])
    # same rule as hsl_select_best / hsl_select_topk on the device: failed candidates (NaN) are never ranked, ties go to
    # the lowest index, best = -1 when nothing is valid
    valid = np.flatnonzero(~np.isnan(costs))
    order = valid[np.argsort(costs[valid], kind="stable")]
    return costs, (int(order[0]) if order.size else -1), order


def torch_all_gather(dist, device=None):
    """all_gather_fn backed by torch.distributed (NCCL when `device` is a CUDA device, gloo on CPU)."""
    import torch

    def fn(local):
        t = torch.from_numpy(np.ascontiguousarray(local))
        if device is not None:
            t = t.to(device)
        out = torch.empty(dist.get_world_size() * t.numel(), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, t)
        return out.cpu().numpy()
    return fn


def fall_sweep_sharded(sweep_fn, kick_step, kick_dv, world, rank, all_gather_fn):
    """The fall / perturbation sweep over several ranks (BASELINE configs[4]: 1e6 perturbed worlds over 8 GPUs).  Worlds are
    independent: rank r integrates its contiguous block of ceil(W/G) worlds with sweep_fn(kick_step, kick_dv) -> dict with
    `fell`, `t_end`, `final_z` (Model.fall_sweep), and the three per-world results (17 B a world) are all-gathered.
    all_gather_fn as in evaluate_sharded (float64 arrays).  Returns dict(fell uint8 [W], t_end [W], final_z [W])."""
    ks = np.ascontiguousarray(kick_step, np.int32)
    kv = np.ascontiguousarray(kick_dv, np.float64).reshape(-1, 3)
    w = ks.shape[0]
    lo, hi, per = shard_bounds(w, world, rank)
    local = np.full((3, per), np.nan)
    if hi > lo:
        r = sweep_fn(ks[lo:hi], kv[lo:hi])
        local[0, :hi - lo] = r["fell"]
        local[1, :hi - lo] = r["t_end"]
        local[2, :hi - lo] = r["final_z"]
    out = []
    for k in range(3):
        g = np.asarray(all_gather_fn(local[k]), np.float64).reshape(world, per)
        out.append(np.concatenate([g[q][:max(0, min(per, w - q * per))] for q in range(world)]))
    return dict(fell=out[0].astype(np.uint8), t_end=out[1], final_z=out[2])
