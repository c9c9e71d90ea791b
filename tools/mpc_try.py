"""Closed-loop experiment: the planner scene as shipped vs with the contact excludes that scene.xml
carries commented out (scene.xml:27-45).  Prints target switches and the distance history."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from manipulator_mujoco_b200.mpc_planner import run_cem_planner
robot = ["base_1", "shoulder_link_1", "upper_arm_link_1", "forearm_link_1", "wrist_1_link_1", "wrist_2_link_1", "wrist_3_link_1", "hande"]
excl = [(b, t) for b in robot for t in ("target_0", "target_1")]
for name, ex in (("scene.xml as shipped", None), ("with the commented-out <exclude> block enabled", excl)):
    res = run_cem_planner(num_dof=6, num_batch=1000, num_steps=16, num_elite=0.05, timestep=0.05, maxiter_cem=3, maxiter_projection=10,
        w_pos=20.0, w_rot=3.0, w_col=80.0, show_viewer=False, show_contact_points=False, initial_qpos=[1.5,-1.8,1.75,-1.25,-1.6,0],
        target_names=["target_0","target_1","target_2","home"], cam_distance=4, position_threshold=0.05, rotation_threshold=0.1,
        save_data=False, data_dir='x', stop_at_final_target=True, max_ticks=2500, verbose=False, contact_exclude=ex)
    g = np.array(res['cost_g']) / 16
    print(name, '| ticks', len(g), '| final target', res['final_target'], '| reached final', res['reached_final'],
          '| median plan ms %.2f' % np.median(res['tick_ms']), '| mean horizon distance every 150 ticks:', np.round(g[::150], 3))
