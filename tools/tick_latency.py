"""Planning-tick latency of the closed-loop configuration (run_mpc_planner.py: B=1000, T=16, 3 CEM iterations)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from manipulator_mujoco_b200 import cem_planner
import contextlib, io
with contextlib.redirect_stdout(io.StringIO()):
    pl = cem_planner(num_dof=6, num_batch=1000, num_steps=16, timestep=0.05, maxiter_cem=3, num_elite=0.05, w_pos=20.0, w_rot=3.0,
                     w_col=80.0, maxiter_projection=10)
q0 = np.array([1.5, -1.8, 1.75, -1.25, -1.6, 0.0]); tp = np.array([-0.3, -0.3, 0.5]); tr = np.array([0.0, 1.0, 0.0, 0.0])
xm = np.zeros(66)
for _ in range(5):
    out = pl.compute_cem(xm, q0, np.zeros(6), np.zeros(6), tp, tr)
t = []
for _ in range(50):
    t0 = time.perf_counter(); out = pl.compute_cem(xm, q0, np.zeros(6), np.zeros(6), tp, tr); t.append(1e3 * (time.perf_counter() - t0))
print(f"closed-loop planning tick (B=1000, T=16, 3 iterations): median {np.median(t):.2f} ms, min {np.min(t):.2f} ms")
