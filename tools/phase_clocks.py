#!/usr/bin/env python3
"""Per-role, per-phase cycle counts from inside the kernels (profiling aid).
Build the instrumented library next to the normal one and point the ctypes layer at it:
    HSL_NVCC_EXTRA=-DHSL_PHASE_CLOCKS python hslabs_b200/build.py --force
    cp hslabs_b200/lib/libhsl_b200.so hslabs_b200/lib/libhsl_b200_clk.so
    python hslabs_b200/build.py --force
    HSL_B200_LIB=$PWD/hslabs_b200/lib/libhsl_b200_clk.so python tools/phase_clocks.py      # on the GPU box
The library prints the averages to stderr (hsl_capi.cu, #ifdef HSL_PHASE_CLOCKS)."""
import os
import sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hslabs_b200 as hsl
from bench import synth_candidates
m = hsl.Model(hsl.model_path('hexapod'))
p = synth_candidates(8192, 20261018)
st = m.eval_gaits(p, 20)['status']
p = np.ascontiguousarray(p[st == 0][:4096])
for fb, mb in ((64, 128), (64, 1)):
    m.set_tuning(fb, mb)
    print('fb', fb, 'maxreg', mb, file=sys.stderr)
    m.eval_gaits(p, 256)
    m.eval_gaits(p, 256)
