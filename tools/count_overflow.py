"""How many samples of a planner-distributed batch exceed the 48-contact total capacity (flag bit 0; GPU box only).
Samples between 21 and 48 contacts use the spill area and are not flagged."""
import contextlib, io, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from manipulator_mujoco_b200 import cem_planner  # noqa: E402

B, T = int(os.environ.get("B", 4096)), int(os.environ.get("T", 100))
with contextlib.redirect_stdout(io.StringIO()):
    pl = cem_planner(num_dof=6, num_batch=B, num_steps=T, timestep=0.05, maxiter_cem=4, num_elite=0.05, w_pos=20.0, w_rot=3.0,
                     w_col=80.0, maxiter_projection=10)
q0 = np.array([1.5, -1.8, 1.75, -1.25, -1.6, 0.0])
for tp in ([-0.3, -0.3, 0.5], [0.3, 0.3, 0.44], [0.0, 0.0, 0.44]):
    real_iter = pl.cem_iter
    counts = []

    def wrapped(carry, x):
        out = real_iter(carry, x)
        counts.append(int((pl._buf("flags", (B,), torch.int32) & 1).sum()))
        return out

    pl.cem_iter = wrapped
    pl.compute_cem(np.zeros(66), q0, np.zeros(6), np.zeros(6), np.array(tp), np.array([0.0, 1.0, 0.0, 0.0]))
    pl.cem_iter = real_iter
    print(f"target {tp}: samples flagged (more than 48 contacts) per CEM iteration: {counts}")
