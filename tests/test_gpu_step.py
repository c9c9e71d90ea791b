"""GPU parity, step level: single teacher-forced steps from states where the robot is in contact
(limits, capsule-box, capsule-capsule), through the C ABI.  See test_gpu_rollout.py for the tolerance
rationale: narrow-phase distances must agree to rounding; the solver output is compared wherever the
reference algorithm itself is precision-stable (float32 and float64 oracle builds agree)."""
import ctypes as C

import numpy as np
import pytest
import torch

from conftest import Q0, TARGET_POS, TARGET_ROT, planner_inputs

pytestmark = pytest.mark.gpu


def test_contact_states_teacher_forced(oracle64, oracle32):
    from manipulator_mujoco_b200 import _lib, cem_planner
    T, B = 100, 64
    pl = cem_planner(num_dof=6, num_batch=B, num_steps=T, timestep=0.05, maxiter_cem=1, num_elite=0.05,
                     w_pos=20.0, w_rot=3.0, w_col=80.0, maxiter_projection=10)
    pr, z, xi, st, xif, td = planner_inputs(T, B, seed=1)
    oth, oep, oer, ocol, oqp, oqa = oracle64.rollout(td, Q0, np.zeros(6), want_state=True)
    has = np.where((ocol < 0).any(axis=(1, 2)))[0]
    assert len(has) >= 8
    warm0 = oracle64.initial_warmstart()
    lib, dev = pl._lib, pl.device
    km = pl.mjx_model
    saved = (list(km.qpos0), list(km.warm0), list(km.qvel0))
    n = stable = 0
    worst_col = 0.0
    devs = []
    try:
        for s in has[:12]:
            qvbox = np.zeros(6)
            for t in range(T):
                qpos = np.concatenate([Q0, oracle64.mc.qpos0[6:]]) if t == 0 else oqp[s, t - 1]
                warm = warm0 if t == 0 else oqa[s, t - 1]
                qvel = np.zeros(12); qvel[6:] = qvbox; qvel[:6] = td[s].reshape(6, T)[:, t]
                qvbox = qvbox + 0.05 * oqa[s, t, 6:]
                if not (ocol[s, t] < 0).any() or t % 4:
                    continue
                for i in range(13):
                    km.qpos0[i] = qpos[i]
                for i in range(12):
                    km.warm0[i], km.qvel0[i] = warm[i], qvel[i]
                _lib.check(lib.cemk_set_model(pl._h, C.byref(km), C.sizeof(km)), lib)
                f = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32, device=pl.device).contiguous()
                tdd, q0, v0, tp, tr = f(qvel[:6].reshape(1, 6)), f(qpos[:6]), f(qvel[:6]), f(TARGET_POS), f(TARGET_ROT)
                theta, cost4 = torch.empty(1, 6, device=dev), torch.empty(1, 4, device=dev)
                col, qacc = torch.empty(1, 1, 187, device=dev), torch.empty(1, 1, 12, device=dev)
                p = lambda x: C.c_void_p(x.data_ptr())
                _lib.check(lib.cemk_rollout_cost(pl._h, 1, 1, p(tdd), p(q0), p(v0), p(tp), p(tr), 20.0, 3.0, 80.0, p(theta), p(cost4),
                                                 None, None, p(col), p(qacc), None, pl._stream()), lib)
                r64, r32 = oracle64.forward(qpos, qvel, warm), oracle32.forward(qpos, qvel, warm)
                worst_col = max(worst_col, np.abs(col[0, 0].cpu().numpy() - r64["con_dist"][oracle64.mask]).max())
                scale = max(1.0, np.abs(r64["qacc"]).max())
                n += 1
                if np.abs(r32["qacc"] - r64["qacc"]).max() < 1e-3 * scale:
                    stable += 1
                    devs.append(np.abs(qacc[0, 0].cpu().numpy() - r64["qacc"]).max() / scale)
    finally:
        for i in range(13):
            km.qpos0[i] = saved[0][i]
        for i in range(12):
            km.warm0[i], km.qvel0[i] = saved[1][i], saved[2][i]
        _lib.check(lib.cemk_set_model(pl._h, C.byref(km), C.sizeof(km)), lib)
    assert n > 30 and stable > 0.3 * n
    assert worst_col < 1e-5
    # MJX's line search returns one of its two bracket ends and accepts candidates on rounding-level comparisons
    # (DESIGN.md section 3): a state on which the oracle's float32 and float64 builds agree can still take the other
    # end under the GPU's FMA-contracted arithmetic, which moves qacc by a few 1e-2 of its scale.  Every state stays
    # inside that jump; the bulk matches to 1e-3.
    dev = np.array(devs)
    print("teacher-forced contact states:", n, "stable:", stable, "dev max %.3g p90 %.3g median %.3g" % (dev.max(), np.percentile(dev, 90), np.median(dev)))
    assert (dev < 5e-2).all(), np.sort(dev)[-5:]
    assert (dev < 1e-3).mean() >= 0.85, np.sort(dev)[-10:]


def test_long_horizon_costs_track_oracle():
    """BASELINE config 2 shape (T=100): contact-rich samples diverge sample-wise (chaotic solver), but
    the cost distribution and the ranking of the low-cost (elite-relevant) samples must agree."""
    from manipulator_mujoco_b200 import cem_planner
    from manipulator_mujoco_b200.mjcf import load_model
    from oracle.oracle import Oracle
    T, B = 100, 256
    pl = cem_planner(num_dof=6, num_batch=B, num_steps=T, timestep=0.05, maxiter_cem=1, num_elite=0.05,
                     w_pos=20.0, w_rot=3.0, w_col=80.0, maxiter_projection=10)
    pr, z, xi, st, xif, td = planner_inputs(T, B, seed=4)
    theta, cost4, ep, er, col = pl._rollout(td, Q0, np.zeros(6), TARGET_POS, TARGET_ROT, True)
    ora = Oracle(load_model(), 0.05)
    oth, oep, oer, ocol = ora.rollout(td, Q0, np.zeros(6))
    oc = pr.compute_cost_batch(oep, oer, ocol, TARGET_POS, TARGET_ROT)
    c = cost4.cpu().numpy()
    free = ~(ocol < 0).any(axis=(1, 2))
    assert free.sum() > 50
    np.testing.assert_allclose(c[free, 0], oc[0][free], rtol=2e-3)
    np.testing.assert_allclose(theta.cpu().numpy()[free], oth[free], atol=1e-3)
    # the k best samples by the oracle are (nearly) the k best by the kernel
    k = 12
    assert len(set(np.argsort(c[:, 0])[:k]) & set(np.argsort(oc[0])[:k + 2])) >= k - 1
    assert np.isfinite(c).all()
