"""Time cemk_project (projection filter + Bernstein evaluation) alone.  GPU box only."""
import contextlib, io, os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from manipulator_mujoco_b200 import cem_planner  # noqa: E402
B, T = int(os.environ.get("B", 4096)), int(os.environ.get("T", 100))
dev = torch.device("cuda:0")
with contextlib.redirect_stdout(io.StringIO()):
    pl = cem_planner(num_dof=6, num_batch=B, num_steps=T, timestep=0.05, maxiter_cem=1, num_elite=0.05, w_pos=20.0, w_rot=3.0,
                     w_col=80.0, maxiter_projection=10, device=dev)
z6 = torch.zeros(6, device=dev)
q0 = torch.tensor([1.5, -1.8, 1.75, -1.25, -1.6, 0.0], device=dev)
st = torch.cat([q0, z6, z6, z6, z6]).unsqueeze(0).expand(B, 30).contiguous()
xi, _ = pl.compute_xi_samples(pl.key, torch.zeros(pl.nvar, device=dev), 10 * torch.eye(pl.nvar, device=dev))
for _ in range(3):
    pl._project(xi, st, True)
ms = []
for _ in range(10):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); pl._project(xi, st, True); b.record(); torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
print(f"project B={B} T={T}: {np.mean(ms) * 1e3:.1f} us (min {np.min(ms) * 1e3:.1f})")
