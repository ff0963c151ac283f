"""One-off check of the bounded wait of the peer-memory gather: rank 1 skips a scatter, rank 0's selection gives up after the
time-out and hsl_gather_check reports it.  Needs a library built with a short time-out:
    HSL_NVCC_EXTRA=-DHSL_GATHER_TIMEOUT_NS=300000000ull python hslabs_b200/build.py --force ; cp hslabs_b200/lib/libhsl_b200.so tools/_lib_to.so
    HSL_B200_LIB=tools/_lib_to.so python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 tools/gather_timeout_probe.py"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.getcwd())
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
import hslabs_b200 as hsl
from hslabs_b200 import api

m = hsl.Model(hsl.model_path("hexapod"))
params, _ = hsl.load_preset(os.path.join("hslabs_b200", "models", "pgs_presets.txt"), 8)
d = torch.from_numpy(np.tile(params, (64, 1))).cuda()
g = api.torch_gather(dist, 64)
best = torch.empty(1, dtype=torch.int64, device="cuda")
m.eval_gaits_scatter(g, 64, 20, d.data_ptr())
g.select_best(best.data_ptr())
g.check()                                   # a complete call: fine
dist.barrier()
if rank == 0:
    t0 = time.perf_counter()
    m.eval_gaits_scatter(g, 64, 20, d.data_ptr())
    g.select_best(best.data_ptr())          # rank 1 never scatters: the wait gives up
    try:
        g.check()
        print("rank 0: NO time-out reported")
    except hsl.HslError as e:
        print("rank 0: reported after %.2f s: %s" % (time.perf_counter() - t0, e))
dist.barrier()
dist.destroy_process_group()
