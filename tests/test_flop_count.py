"""The op-counting build of the oracle (oracle/count_real.h, tools/count_flops.py) that defines the roofline's F_A:
its exact shortcuts must not change any result, and the number frozen in profiles/ must be what the counter says."""
import json
import os

import numpy as np

from conftest import Q0, ROOT, planner_inputs


def test_counting_build_reproduces_the_float64_oracle_bit_for_bit(mc, oracle64):
    from oracle.oracle import Oracle
    _, _, _, _, _, td = planner_inputs(24, 48, seed=5)
    cnt = Oracle(mc, 0.05, dtype="count")
    a = oracle64.rollout(td, Q0, np.zeros(6), want_state=True)
    b = cnt.rollout(td, Q0, np.zeros(6), want_state=True)
    for x, y, name in zip(a, b, ("theta", "eef_pos", "eef_rot", "collision", "qpos", "qacc")):
        assert np.array_equal(x, y), name            # far-pair / disjoint-box / inactive-frame shortcuts are exact


def test_frozen_flop_count_is_what_the_counter_measures(mc):
    from oracle.oracle import Oracle
    with open(os.path.join(ROOT, "profiles", "r2_flop_count.json")) as f:
        rec = json.load(f)
    assert abs(sum(rec["per_stage"].values()) + rec["cost_accumulation_analytic"] - rec["flop_per_env_step"]) < 1e-6
    _, _, _, _, _, td = planner_inputs(100, 32, seed=0)
    cnt = Oracle(mc, 0.05, dtype="count")
    warm = cnt.initial_warmstart()
    cnt.read_counts()
    cnt.rollout(td, Q0, np.zeros(6), warm=warm)
    c = cnt.read_counts()
    per_step = sum(v for k, v in c.items() if not k.startswith("not counted")) / (32 * 100) + rec["cost_accumulation_analytic"]
    assert abs(per_step - rec["flop_per_env_step"]) < 0.1 * rec["flop_per_env_step"], (per_step, rec["flop_per_env_step"])
    # the smooth-dynamics stages do a fixed amount of work per step
    assert c["kinematics"] % (32 * 100) == 0 and c["com_pos + CRBA + factor"] % (32 * 100) == 0
