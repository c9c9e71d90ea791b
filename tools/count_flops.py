"""Algorithmic work per env-step (F_A of the roofline), counted instead of estimated.

Runs the op-counting build of the CPU restatement (oracle/mjstep.c with -DORACLE_COUNT: `real` is a double that tallies
add / sub / mul / div / sqrt / sin / cos / pow = 1 each, so a multiply-add = 2; comparisons, min/max, abs, copies = 0;
exact shortcuts for far capsule-box pairs, disjoint box-box pairs and the frames of inactive slots -- see count_real.h)
on the rollout inputs of bench.py's workload: the planner's own first-iteration samples (PRNGKey(0) -> xi ~ N(0, 10 I)
-> projection filter -> thetadot), scene A, T = 100.  A subset of the 4096 samples is enough: the count per env-step
is an average over samples and steps (default 512 samples = 51200 env-steps, a minute on 8 cores).

    python tools/count_flops.py [n_samples] [out.json]     -> profiles/r2_flop_count.json
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402
from manipulator_mujoco_b200.mjcf import load_model  # noqa: E402


def main(n=512, out=os.path.join(ROOT, "profiles", "r2_flop_count.json")):
    cem = bench.CpuCem(bench.B_PER_GPU, bench.T, os.cpu_count() or 1)
    f32 = np.float32
    pr = cem.pr
    z = cem.normal(bench.B_PER_GPU * pr.nvar).reshape(bench.B_PER_GPU, pr.nvar)
    xi = pr.compute_xi_samples(z, np.zeros(pr.nvar, f32), 10 * np.identity(pr.nvar, dtype=f32)).astype(f32)
    st = pr.state_term(bench.Q0, np.zeros(6), np.zeros(6), bench.B_PER_GPU).astype(f32)
    td = (pr.compute_projection_filter(xi, st) @ pr.A_thetadot.T).astype(np.float64)
    sel = np.linspace(0, bench.B_PER_GPU - 1, n).astype(int)             # evenly spread over the batch
    ora = Oracle(load_model(), bench.DT, dtype="count")
    warm = ora.initial_warmstart()
    ora.read_counts()                                                    # discard the warm-start forward
    ora.rollout(td[sel], bench.Q0, np.zeros(6), warm=warm, want_collision=True)
    cnt = ora.read_counts()
    steps = n * bench.T
    per = {k: v / steps for k, v in cnt.items()}
    uncounted = per.pop(Oracle.STAGES[7])
    # planner-side cost accumulation fused into the rollout kernel (mjx_planner.py:277-296), not part of mjx.step:
    # per robot slot one multiply-subtract + one accumulate (187 slots), the goal / orientation terms ~30 flop
    cost_acc = 187 * 3 + 30
    total = sum(per.values()) + cost_acc
    rec = {"flop_per_env_step": total, "per_stage": per, "cost_accumulation_analytic": cost_acc,
           "excluded_dense_unobservable_per_env_step": uncounted,
           "counting_rule": "add/sub/mul/div/sqrt/sin/cos/pow = 1 (multiply-add = 2); comparisons, min/max, abs, negation, copies, integer work = 0",
           "workload": f"scene A, T={bench.T}, {n} of the {bench.B_PER_GPU} first-iteration planner samples of bench.py (PRNGKey(0), N(0,10 I), 10 projection iterations)",
           "env_steps_counted": steps, "tool": "tools/count_flops.py (oracle/mjstep.c -DORACLE_COUNT, float64 arithmetic)",
           "previous_estimate": 8.0e4}
    print(json.dumps(rec, indent=1))
    with open(out, "w") as f:
        json.dump(rec, f, indent=1)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 512, *(sys.argv[2:3]))
