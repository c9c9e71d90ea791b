"""Closed-loop run at the reference's canonical parameters (run_mpc_planner.py:7-44) next to the reference's own
recording (tests/golden/closed_loop_kat.npz = data/cost_c.csv, cost_g.csv of 897 ticks): target switching and the
distribution of the best sample's cost_c.  `python tools/closed_loop_report.py [max_ticks] [out.json]`"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from manipulator_mujoco_b200.mpc_planner import run_cem_planner


def main(max_ticks=1200, out=None):
    res = run_cem_planner(num_dof=6, num_batch=1000, num_steps=16, num_elite=0.05, timestep=0.05, maxiter_cem=3, maxiter_projection=10,
                          w_pos=20.0, w_rot=3.0, w_col=80.0, show_viewer=False, show_contact_points=False,
                          initial_qpos=[1.5, -1.8, 1.75, -1.25, -1.6, 0], target_names=["target_0", "target_1", "target_2", "home"],
                          cam_distance=4, position_threshold=0.05, rotation_threshold=0.1, save_data=False, data_dir='x',
                          stop_at_final_target=True, max_ticks=max_ticks, verbose=False)
    c = np.array(res['cost_c'])
    rec = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "closed_loop_kat.npz"))
    q = lambda a: [float(x) for x in np.percentile(a, [50, 90, 99, 100])]
    rep = {"ticks": len(c), "final_target": res['final_target'], "reached_final": bool(res['reached_final']),
           "switch_ticks": res.get('switch_ticks'), "median_plan_ms": float(np.median(res['tick_ms'])),
           "cost_c_p50_p90_p99_max": q(c), "cost_c_nonzero_frac": float((c > 0).mean()),
           "recorded_cost_c_p50_p90_p99_max": q(rec['cost_c']), "recorded_cost_c_nonzero_frac": float((rec['cost_c'] > 0).mean()),
           "recorded_ticks": int(rec['cost_c'].size),
           "dist_every_50_ticks": [float(x) for x in np.array(res['dist'])[::50]]}
    print(json.dumps(rep, indent=1))
    if out:
        with open(out, "w") as f:
            json.dump(rep, f, indent=1)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 1200, sys.argv[2] if len(sys.argv) > 2 else None)
