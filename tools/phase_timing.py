"""Per-phase share of k_rollout's time (SM clocks of one lane per lane group, summed; a warp carries two groups).
The clock read that closes a barrier phase issues before BAR.SYNC completes, so the waiting time of a barrier
shows up in the phase that follows it (cross-checked with ncu: the stall samples sit on the instruction behind BAR.SYNC).
Builds a debug library with -DCEMK_PHASE_TIMING next to the normal one, runs one 4096 x 100 rollout.
    python tools/phase_timing.py            (on a GPU box)"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from manipulator_mujoco_b200 import _lib  # noqa: E402

dbg = os.path.join(ROOT, "manipulator_mujoco_b200", "libcemk_phase.so")
cmd = ["nvcc"] + [f for f in _lib.NVCC_FLAGS] + os.environ.get("EXTRA", "").split() + ["-DCEMK_PHASE_TIMING", "-o", dbg, _lib.SRC[0]]
subprocess.run(cmd, check=True)
_lib.LIB_PATH = dbg
import torch  # noqa: E402
from manipulator_mujoco_b200 import cem_planner  # noqa: E402

B, T = int(os.environ.get("B", 4096)), int(os.environ.get("T", 100))
pl = cem_planner(num_dof=6, num_batch=B, num_steps=T, timestep=0.05, maxiter_cem=1, num_elite=0.05, w_pos=20.0, w_rot=3.0,
                 w_col=80.0, maxiter_projection=10)
q0 = np.array([1.5, -1.8, 1.75, -1.25, -1.6, 0.0])
tp, tr = np.array([-0.3, -0.3, 0.5]), np.array([0.0, 1.0, 0.0, 0.0])
lib = C.CDLL(dbg)
buf = (C.c_ulonglong * 24)()
for _ in range(2):
    pl.compute_cem(np.zeros(66), q0, np.zeros(6), np.zeros(6), tp, tr)
torch.cuda.synchronize()
pl._lib.cemk_debug_phase_clocks(buf)
pl._lib.cemk_debug_phase_cond((C.c_ulonglong * 25)())        # (reading clears the tables)
pl._lib.cemk_debug_events((C.c_ulonglong * 24)())
pl.compute_cem(np.zeros(66), q0, np.zeros(6), np.zeros(6), tp, tr)
torch.cuda.synchronize()
pl._lib.cemk_debug_phase_clocks(buf)
names = ["step barrier wait (sampled behind BAR.SYNC) + command load", "P1 FK chain", "P2-P5 dynamics", "P6 qacc_smooth", "N1 robot narrow phase + cost", "N1 free-box pairs", "N2 emit contacts",
         "phase barrier wait (sampled behind BAR.SYNC) + C1 limit rows", "S1 warm/smooth", "S3 grad + H", "S4 Cholesky + solve", "S5a line-search setup", "obs + euler", "step barrier issue",
         "prologue", "epilogue", "C2 contact Jacobians", "C3 row parameters", "(issue of phase barrier)", "S5b line-search trips", "N1b ballots + near list", "N1c near capsule-box pass", "N1d deferred accounting", "N1e park"]
v = np.array(list(buf), dtype=np.float64)
print(f"k_rollout phase shares, B={B} T={T} (clock64 per warp, summed)")
for n, x in sorted(zip(names, v), key=lambda t: -t[1]):
    print(f"  {100 * x / v.sum():6.2f}%  {x / (B * T):9.0f} clk/env-step  {n}")
print(f"  total {v.sum() / (B * T):.0f} clk per env-step per warp")
pc = (C.c_ulonglong * 25)()
pl._lib.cemk_debug_phase_cond(pc)
c = np.array(list(pc)[:24], dtype=np.float64)
nflag = max(float(pc[24]), 1.0)
nall = B * T                       # one lane per lane group counts: group-steps
rest_n = max(nall - nflag, 1.0)
print(f"conditional profile: clk per step of a warp whose step has robot contacts ({int(nflag)} group-steps) vs the other steps")
for n, xc, xa in sorted(zip(names, c, v), key=lambda t: -t[1]):
    print(f"  {xc / nflag:9.0f} vs {(xa - xc) / rest_n:9.0f}   {n}")
print(f"  {c.sum() / nflag:9.0f} vs {(v.sum() - c.sum()) / rest_n:9.0f}   total")
ev = (C.c_ulonglong * 24)()
pl._lib.cemk_debug_events(ev)
e = [float(x) for x in ev]
n = max(e[0], 1.0)
print("events per sample-step (lane group), B x T =", B * T, "counted", int(e[0]))
for label, val in [("has near capsule-box pairs", e[1] / n), ("near pairs (mean)", e[2] / n), ("warp runs the near pass (either sample)", e[11] / n),
                   ("active contacts (mean)", e[3] / n), ("has robot contacts", e[4] / n), ("warp emits robot contacts (either sample)", e[12] / n),
                   ("no active row", e[8] / n), ("has active limit rows", e[13] / n), ("coupled 12x12 solve", e[7] / n), ("spill instantiation", e[9] / n),
                   ("line-search trips executed by the warp (mean)", e[6] / n), ("line-search trips the sample needed (mean)", e[5] / n),
                   ("free-box pairs walked by the warp (mean)", e[10] / n),
                   ("clk per near-pass trip: pair evaluation block", e[14] / max(e[11], 1.0)), ("clk per near-pass trip: scan + contact records", e[15] / max(e[11], 1.0)),
                   ("clk of capbox_local + capsule_box_near (edges included) inside the evaluation block, lane 0", e[16] / max(e[17], 1.0))]:
    print(f"  {val:8.3f}  {label}")
os.remove(dbg)
