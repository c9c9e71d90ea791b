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
    # which robot slots are in contact at the final plant state (the t = 0 collision of every rollout)
    cem = res.get('cem')
    if cem is not None:
        mc = cem._mc
        names = []
        for (g1, g2), a, n in zip(mc.pair_geom, mc.pair_slotadr, mc.pair_nslot):
            if g1 in cem.geom_ids or g2 in cem.geom_ids:
                names += [f"{mc.geom_names[g1]}|{mc.geom_names[g2]}#{k}" for k in range(n)]
        import torch
        q = np.array(res['theta'][-1])
        _, _, _, col = cem.compute_rollout_batch(torch.zeros(1, 6 * cem.num, device=cem.device), q, np.zeros(6))
        c0 = col[0, 0].cpu().numpy()
        rep["final_state_contacts"] = {names[i]: float(c0[i]) for i in np.where(c0 < 0)[0]}
        rep["final_state_near"] = {names[i]: float(c0[i]) for i in np.where((c0 >= 0) & (c0 < 0.03))[0]}
        # where does the best plan's cost_c at the final state come from?  (slot, step, term)
        L = res['last']
        plan = cem.compute_cem(L['xi_mean'], L['qpos'][:6], L['qvel'][:6], L['qacc'][:6], L['target_pos'], L['target_rot'])
        td = torch.as_tensor(plan[4].T.reshape(1, -1).copy(), dtype=torch.float32, device=cem.device)
        _, _, _, col = cem.compute_rollout_batch(td, L['qpos'][:6], L['qvel'][:6])
        cc = col[0].cpu().numpy()
        terms = []
        for t in range(cc.shape[0]):
            for i in range(cc.shape[1]):
                if cc[t, i] < 0:
                    terms.append((t, names[i], "count", float(cc[t, i])))
                if t > 0 and 0.995 * cc[t - 1, i] - cc[t, i] > 0.05:
                    terms.append((t, names[i], "approach", float(cc[t - 1, i]), float(cc[t, i])))
        rep["final_best_plan_cost_c"] = float(plan[3])
        rep["final_best_plan_terms"] = terms[:40]
    c_ts = np.array(res['cost_c'])
    rep["cost_c_first_100_ticks_p50_max"] = [float(np.median(c_ts[:100])), float(c_ts[:100].max())]
    rep["cost_c_by_100_ticks_median"] = [float(np.median(c_ts[i:i + 100])) for i in range(0, len(c_ts), 100)]
    print(json.dumps(rep, indent=1))
    if out:
        with open(out, "w") as f:
            json.dump(rep, f, indent=1)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 1200, sys.argv[2] if len(sys.argv) > 2 else None)
