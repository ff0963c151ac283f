"""Worker of tests/test_gpu_fall.py: the CPU side of the fall sweep -- the reference's own position-control loop
(player.cpp) on the ODE shim's stepper (oracle/shim/ode_step.cpp) -- in a process of its own, because the reference keeps
ONE static ODE world per process (visualization.cpp:165) and the simulation needs it to hold this model only.

    python tests/fall_cpu_worker.py <in.npz> <out.npz>
in: xml, params[13], n_steps, kick_step[W], kick_dv[W][3], play_dt, t0, hc, tmin, nprocs, n_traj (first n_traj worlds also return
their torso trajectory)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(inp, outp):
    from oracle import ref
    a = np.load(inp, allow_pickle=False)
    m = ref.Model(str(a["xml"]))
    n_steps = int(a["n_steps"])
    kw = dict(play_dt=float(a["play_dt"]), t0=float(a["t0"]), hc=float(a["hc"]), tmin=float(a["tmin"]))
    res = m.fall_batch(a["params"], n_steps, a["kick_step"], a["kick_dv"], nprocs=int(a["nprocs"]), **kw)
    trajs = []
    for i in range(int(a["n_traj"])):
        r = m.fall_run(a["params"], n_steps, kick_step=int(a["kick_step"][i]), kick_dv=a["kick_dv"][i], want_traj=True, **kw)
        t = np.full((n_steps, 3), np.nan)
        t[:r["steps"]] = r["traj"]
        trajs.append(t)
    np.savez(outp, fell=res["fell"], t_end=res["t_end"], final_z=res["final_z"], traj=np.array(trajs))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
