"""Worker of tests/test_gpu_multirank.py: one process per GPU (torch.distributed.run), NCCL.
Every rank checks that the sharded DeviceEvaluator / search results equal its own single-GPU evaluation of the whole
batch and writes rank<r>.ok into the directory given on the command line."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(out_dir):
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import hslabs_b200 as hsl
    from hslabs_b200 import search
    model = hsl.Model(hsl.model_path("hexapod"))
    rng = np.random.default_rng(3)
    n, n_t = 1001, 32                      # odd count: ragged last shard
    p = np.zeros((n, 13))
    p[:, 2] = rng.uniform(-0.15, -0.05, n); p[:, 6] = rng.uniform(0, 1, n); p[:, 7] = rng.uniform(1, 6, n)
    p[:, 8] = rng.uniform(0.1, 0.5, n); p[:, 9] = rng.uniform(0.02, 0.12, n); p[:, 11] = -1
    p[17, 2] = 0.4                         # unreachable: NaN cost on every rank
    single = model.eval_gaits(p, n_t)
    cand = torch.from_numpy(p).to(dev)
    ev = search.DeviceEvaluator(model, n_t)               # sharded over the ranks; costs gathered over NVLink peer memory
    cot, st = ev(cand)
    got = cot.cpu().numpy()
    assert np.array_equal(got, single["cot"], equal_nan=True), "sharded costs differ from the single-GPU evaluation"
    assert np.array_equal(st.cpu().numpy(), single["status"])
    cot_n, st_n = search.DeviceEvaluator(model, n_t, gather="nccl")(cand)   # the same through torch.distributed's all-gather
    assert torch.equal(cot_n.nan_to_num(nan=-1.0), cot.nan_to_num(nan=-1.0)) and torch.equal(st_n, st)
    # run-ahead: 40 back-to-back gathers of changing batches, one rank delayed every few calls; every result is checked
    # after the loop, so a rank that overwrote a buffer still in use would show
    outs = []
    for it in range(40):
        if it % 7 == rank % 7:
            torch.cuda._sleep(20_000_000)
        sub = cand[(it * 13) % 200:(it * 13) % 200 + 257 + it]
        outs.append((sub, ev(sub)[0]))
    for sub, c in outs:
        want = model.eval_gaits(sub.cpu().numpy(), n_t)["cot"]
        assert np.array_equal(c.cpu().numpy(), want, equal_nan=True), "gathered costs differ after back-to-back calls"
    # the fused form the bench uses: scatter (no extra launch) + an argmin kernel that waits for the flags itself
    from hslabs_b200 import api
    per = -(-n // world)
    lo, hi = min(rank * per, n), min(rank * per + per, n)
    g = api.torch_gather(dist, per)
    best = torch.empty(1, dtype=torch.int64, device=dev)
    bval = torch.empty(1, dtype=torch.float64, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    key = np.where(np.isnan(single["cot"]), np.inf, single["cot"])
    for it in range(10):
        if it % 3 == rank % 3:
            torch.cuda._sleep(10_000_000)
        model.eval_gaits_scatter(g, hi - lo, n_t, cand[lo:hi].contiguous().data_ptr(), stream=s)
        g.select_best(best.data_ptr(), bval.data_ptr(), s)
        assert int(best.item()) == int(np.argmin(key)) and float(bval.item()) == float(key.min())
    torch.cuda.synchronize()
    dist.barrier()
    g.free()
    # gather="auto" when a rank cannot map the buffers: every rank falls back to NCCL, same results

    class _FailingLib:
        def __getattr__(self, name):
            if name == "hsl_gather_connect" and rank == world - 1:
                return lambda *a: -4
            return getattr(real_lib, name)
    real_lib = api._load()
    api._lib = _FailingLib()
    try:
        assert api.torch_gather(dist, 64, fallback=True) is None
        ev_auto = search.DeviceEvaluator(model, n_t, gather="auto")
        cot_a, st_a = ev_auto(cand)
        assert ev_auto._gathers and all(v is None for v in ev_auto._gathers.values())
        assert torch.equal(cot_a.nan_to_num(nan=-1.0), cot.nan_to_num(nan=-1.0)) and torch.equal(st_a, st)
    finally:
        api._lib = real_lib
    # an empty shard (fewer candidates than ranks) and a batch of one
    for k in (1, world - 1 if world > 1 else 1):
        c1, _ = ev(cand[:k])
        assert np.array_equal(c1.cpu().numpy(), single["cot"][:k], equal_nan=True)
    idx, val = search.top_k(cot, 5)
    key = np.where(np.isnan(single["cot"]), np.inf, single["cot"])
    assert list(idx.cpu().numpy()) == list(np.argsort(key, kind="stable")[:5])
    res = search.cem_search(model, p[0], {"period": (1.0, 6.0), "step_length": (0.1, 0.5)}, n_t, pop=256, elite=32, iters=3, seed=1)
    allbest = [None] * world
    dist.all_gather_object(allbest, float(res["best_cot"]))
    assert all(b == allbest[0] for b in allbest), "ranks disagree on the search result"
    rs = search.random_search(model, p[0], {"period": (1.0, 6.0), "step_length": (0.1, 0.5)}, 300, n_t, seed=4, k=3)   # sharded, gather freed after
    one = model.eval_gaits(rs["candidates"].cpu().numpy(), n_t)["cot"]
    assert np.array_equal(rs["cot"].cpu().numpy(), one, equal_nan=True)
    # the C-ABI collective (hsl_allgather_costs over an ncclComm_t created from a broadcast unique id)
    if hasattr(hsl.api, "nccl_allgather_selftest"):
        hsl.api.nccl_allgather_selftest(rank, world, dist, dev)
    # the fall sweep sharded over the ranks: the same outcomes as all worlds on one GPU (worlds are independent)
    from hslabs_b200.sharding import fall_sweep_sharded, torch_all_gather
    fparams, _ = hsl.load_preset(os.path.join(ROOT, "hslabs_b200", "models", "pgs_presets.txt"), 8)
    frng = np.random.default_rng(9)
    fw = 67
    fmag, fth = frng.uniform(0, 16, fw), frng.uniform(0, 2 * np.pi, fw)
    fkv = np.stack([fmag * np.cos(fth), fmag * np.sin(fth), np.zeros(fw)], axis=1)
    fks = frng.integers(5, 100, fw).astype(np.int32)
    sh = fall_sweep_sharded(lambda a, b: model.fall_sweep(fparams, 200, a, b), fks, fkv, world, rank, torch_all_gather(dist, dev))
    one = model.fall_sweep(fparams, 200, fks, fkv)
    assert np.array_equal(sh["fell"], one["fell"]) and np.array_equal(sh["t_end"], one["t_end"]) and np.array_equal(sh["final_z"], one["final_z"])
    assert one["fell"].any() and not one["fell"].all()
    ev.close()
    dist.barrier()
    open(os.path.join(out_dir, "rank%d.ok" % rank), "w").write("ok\n")
    dist.destroy_process_group()


if __name__ == "__main__":
    main(sys.argv[1])
