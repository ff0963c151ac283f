import sys, numpy as np
sys.path.insert(0, '.')
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
