"""Non-finite samples of the full-size test batch: how many, and when (GPU box only)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np, torch, contextlib, io
from conftest import planner_inputs
from manipulator_mujoco_b200 import cem_planner
B, T = 4096, 100
Q0 = np.array([1.5, -1.8, 1.75, -1.25, -1.6, 0.0]); TP = np.array([-0.3, -0.3, 0.5]); TR = np.array([0.0, 1.0, 0.0, 0.0])
with contextlib.redirect_stdout(io.StringIO()):
    pl = cem_planner(num_dof=6, num_batch=B, num_steps=T, timestep=0.05, maxiter_cem=1, num_elite=0.05, w_pos=20.0, w_rot=3.0, w_col=80.0, maxiter_projection=10)
pr, z, xi, st, xif, td = planner_inputs(T, B, seed=21)
xi_f, thetadot = pl._project(xi, st, True)
theta, cost4, ep, er, col = pl._rollout(thetadot, Q0, np.zeros(6), TP, TR, True)
c = cost4.cpu().numpy(); th = theta.cpu().numpy().reshape(B, 6, T)
bad_c = ~np.isfinite(c).all(axis=1); bad_t = ~np.isfinite(th).all(axis=(1, 2))
print("non-finite cost:", int(bad_c.sum()), " non-finite theta:", int(bad_t.sum()), " theta-only:", int((bad_t & ~bad_c).sum()))
for s in np.where(bad_t)[0][:12]:
    first = int(np.argmax(~np.isfinite(th[s]).all(axis=0)))
    ncol = (col[s].cpu().numpy() < 0).sum(axis=1)
    print(f"  sample {s}: first non-finite theta at step {first}, robot contacts around then {ncol[max(0, first - 3):first + 1]}, cost finite {not bad_c[s]}")
