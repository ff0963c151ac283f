"""World-size-2 gloo test of the multi-GPU host logic (sharding + all-gather + selection), CPU only.
The local evaluator here is the CPU oracle standing in for the per-rank GPU evaluation; the `-m gpu` tier and
bench.py run the same logic over NCCL with the CUDA path."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, model_xml


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, params, n_t, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hslabs_b200.sharding import evaluate_sharded, torch_all_gather
    from oracle import orc
    m = orc.Model(model_xml("hexapod"))

    def local_eval(p):
        r = m.eval_batch(p, n_t, 1)
        return np.where(r["status"] == 0, r["cot"], np.nan)
    costs, best, order = evaluate_sharded(local_eval, params, world, rank, torch_all_gather(dist))
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), costs=costs, best=best, order=order)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_bounds():
    from hslabs_b200.sharding import shard_bounds
    for n in (1, 7, 8, 9, 65536):
        for w in (1, 2, 4, 8):
            cover = []
            for r in range(w):
                lo, hi, per = shard_bounds(n, w, r)
                assert 0 <= lo <= hi <= n and hi - lo <= per
                cover += list(range(lo, hi))
            assert cover == list(range(n))


@pytest.mark.timeout(300)
def test_two_rank_gather_matches_single(orc, tmp_path):
    rng = np.random.default_rng(5)
    n, n_t = 7, 12  # odd count: ragged last shard
    p = np.zeros((n, 13))
    p[:, 2] = rng.uniform(-0.15, -0.05, n); p[:, 6] = rng.uniform(0, 1, n); p[:, 7] = rng.uniform(1, 6, n)
    p[:, 8] = rng.uniform(0.1, 0.5, n); p[:, 9] = rng.uniform(0.02, 0.12, n); p[:, 11] = -1
    p[3, 2] = 0.4  # unreachable candidate -> NaN cost, must rank last
    port = _free_port()
    mp.spawn(_worker, args=(2, port, p, n_t, str(tmp_path)), nprocs=2, join=True)
    ref = orc.Model(model_xml("hexapod")).eval_batch(p, n_t, 1)
    ref_cost = np.where(ref["status"] == 0, ref["cot"], np.nan)
    a = np.load(tmp_path / "rank0.npz")
    b = np.load(tmp_path / "rank1.npz")
    assert np.array_equal(a["costs"], b["costs"], equal_nan=True) and a["best"] == b["best"]
    assert np.array_equal(a["costs"], ref_cost, equal_nan=True)
    assert np.isnan(a["costs"][3]) and 3 not in a["order"] and len(a["order"]) == n - 1  # failed candidates are not ranked
    assert a["best"] == int(np.nanargmin(ref_cost))


def test_selection_rule_matches_the_device_one():
    """no valid candidate -> best index -1 (like hsl_select_best), NaN never ranked, list input accepted."""
    from hslabs_b200.sharding import evaluate_sharded
    costs, best, order = evaluate_sharded(lambda p: np.full(len(p), np.nan), [[0.0] * 13] * 3, 1, 0, lambda x: x)
    assert best == -1 and order.size == 0 and np.isnan(costs).all()
    vals = np.array([2.0, np.nan, 1.0, 1.0])
    costs, best, order = evaluate_sharded(lambda p: vals, np.zeros((4, 13)), 1, 0, lambda x: x)
    assert best == 2 and list(order) == [2, 3, 0]


def test_fall_sweep_sharding_logic():
    """fall_sweep_sharded: contiguous blocks of worlds per rank, ragged and empty shards, results in world order."""
    from hslabs_b200.sharding import fall_sweep_sharded, shard_bounds

    def sweep(ks, kv):   # stands in for Model.fall_sweep on a rank's block
        return dict(fell=(np.abs(kv[:, 1]) > 8).astype(np.uint8), t_end=ks * 0.02, final_z=kv[:, 0] + 0.5)
    rng = np.random.default_rng(4)
    for w, world in ((7, 2), (1, 4), (64, 8), (9, 4)):
        ks = rng.integers(0, 100, w).astype(np.int32)
        kv = rng.uniform(-16, 16, (w, 3))
        want = sweep(ks, kv)
        for rank in range(world):
            def gather(local, _k=[0]):
                k = _k[0] % 3
                _k[0] += 1
                per = local.shape[0]
                rows = []
                for q in range(world):
                    lo, hi, _ = shard_bounds(w, world, q)
                    row = np.full(per, np.nan)
                    if hi > lo:
                        r = sweep(ks[lo:hi], kv[lo:hi])
                        row[:hi - lo] = (r["fell"], r["t_end"], r["final_z"])[k]
                    rows.append(row)
                assert np.array_equal(rows[rank], local, equal_nan=True)   # what this rank contributes is its own block
                return np.concatenate(rows)
            got = fall_sweep_sharded(sweep, ks, kv, world, rank, gather)
            for key in ("fell", "t_end", "final_z"):
                assert np.array_equal(got[key], want[key]), (w, world, rank, key)
