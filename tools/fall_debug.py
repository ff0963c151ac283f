import os, sys, numpy as np
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests")
import hostlib, hslabs_b200 as hsl
xml = hsl.model_path("hexapod")
p, _ = hsl.load_preset(os.path.join("hslabs_b200", "models", "pgs_presets.txt"), 8)
ks = np.array([-1], np.int32); kv = np.zeros((1, 3))
os.environ["HSL_FALL_DEBUG"] = "gpurun_out/fall_dbg_host.bin"
h = hostlib.fall_sweep(xml, p, 20, ks, kv, want_traj=True)
os.environ["HSL_FALL_DEBUG"] = "gpurun_out/fall_dbg_gpu.bin"
g = hsl.Model(xml).fall_sweep(p, 20, ks, kv, want_traj=True)
a = np.fromfile("gpurun_out/fall_dbg_host.bin", np.float64, 150 * 54 + 96 + 128)
b = np.fromfile("gpurun_out/fall_dbg_gpu.bin", np.float64, 150 * 54 + 96 + 128)
print("ctrl diff", np.abs(a[:8100] - b[:8100]).max(), "pos0 diff", np.abs(a[8100:8196] - b[8100:8196]).max(), "quat0 diff", np.abs(a[8196:] - b[8196:]).max())
ra = open("gpurun_out/fall_dbg_host.bin", "rb").read()[(8100 + 224) * 8:]
rb = open("gpurun_out/fall_dbg_gpu.bin", "rb").read()[(8100 + 224) * 8:]
print("simpod equal", ra == rb, len(ra))
print("traj diff per step", np.abs(h["traj"][0] - g["traj"][0]).max(axis=1))
c = a[:8100].reshape(150, 3, 18); d = b[:8100].reshape(150, 3, 18)
print("ctrl diff by part", np.abs(c - d).max(axis=(0, 2)))
