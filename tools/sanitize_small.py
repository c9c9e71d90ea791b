"""Smallest end-to-end case with robot contacts, for `compute-sanitizer --tool memcheck python tools/sanitize_small.py`
(one tool per gpurun call).  Shared-memory hazards between the lanes of a sample are NOT checked this way -- racecheck does
not follow named barriers over 16-lane groups; that check is the race-detecting emulation build of the same source,
tests/test_emu_race.py (csrc/warp_dsl.h, CEMK_EMU_RACE)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manipulator_mujoco_b200 import cem_planner

pl = cem_planner(num_dof=6, num_batch=40, num_steps=12, timestep=0.05, maxiter_cem=2, num_elite=0.1, w_pos=20.0, w_rot=3.0, w_col=80.0,
                 maxiter_projection=3)
pl.use_cuda_graph = False
q0 = np.array([1.5, -1.8, 1.75, -1.25, -1.6, 0.0])
out = pl.compute_cem(np.zeros(66), q0, np.zeros(6), np.zeros(6), np.array([-0.3, -0.3, 0.5]), np.array([0.0, 1.0, 0.0, 0.0]))
# a state with robot contacts: start low over the table so the narrow phase emits contacts
out2 = pl.compute_cem(np.zeros(66), np.array([1.5, -0.9, 2.2, -1.25, -1.6, 0.0]), np.zeros(6), np.zeros(6), np.array([-0.3, -0.3, 0.5]),
                      np.array([0.0, 1.0, 0.0, 0.0]))
print("costs", out[0], out2[0])
