"""Workload for an ncu capture of the selection kernels and of the finish kernel with the fused gather (one rank):
    ncu --set full --clock-control none --import-source on -k regex:"topk_prune|argmin|finish" -c 8 -o gpurun_out/select python tools/profile_select.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.getcwd())
import hslabs_b200 as hsl
from hslabs_b200 import api

n, k = 65536, 4096
t = torch.rand(n, dtype=torch.float64, device="cuda")
idx = torch.empty(k, dtype=torch.int64, device="cuda")
val = torch.empty(k, dtype=torch.float64, device="cuda")
for _ in range(2):
    api.select_topk_device(t.data_ptr(), n, k, idx.data_ptr(), val.data_ptr())
m = hsl.Model(hsl.model_path("hexapod"))
params, _ = hsl.load_preset(os.path.join("hslabs_b200", "models", "pgs_presets.txt"), 8)
rng = np.random.default_rng(0)
batch = np.tile(params, (4096, 1))
batch[:, 7] = rng.uniform(2.0, 5.0, 4096)
d = torch.from_numpy(batch).cuda()
g = api.Gather(0, 1, 4096, lambda h: [h])
best = torch.empty(1, dtype=torch.int64, device="cuda")
for _ in range(2):
    m.eval_gaits_scatter(g, 4096, 256, d.data_ptr())
    g.select_best(best.data_ptr())
torch.cuda.synchronize()
print("ok", int(best.item()))
