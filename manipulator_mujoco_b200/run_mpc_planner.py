"""The reference's canonical closed-loop configuration (sampling_based_planner/run_mpc_planner.py:7-44),
headless: `python -m manipulator_mujoco_b200.run_mpc_planner [max_ticks]`."""
import sys

from .mpc_planner import run_cem_planner


def main(max_ticks=300):
    return run_cem_planner(
        num_dof=6, num_batch=1000, num_steps=16, num_elite=0.05, timestep=0.05, maxiter_cem=3, maxiter_projection=10,
        w_pos=20.0, w_rot=3.0, w_col=80.0, show_viewer=False, show_contact_points=True,
        initial_qpos=[1.5, -1.8, 1.75, -1.25, -1.6, 0], target_names=["target_0", "target_1", "target_2", "home"],
        cam_distance=4, position_threshold=0.05, rotation_threshold=0.1, save_data=True, data_dir='custom_data',
        stop_at_final_target=True, max_ticks=max_ticks)


if __name__ == "__main__":
    res = main(int(sys.argv[1]) if len(sys.argv) > 1 else 300)
    import numpy as np
    print(f"ticks {len(res['theta'])}  median planning latency {np.median(res['tick_ms']):.2f} ms  final target {res['final_target']}  reached_final {res['reached_final']}")
