"""GPU parity of the fused rollout + cost kernel against the CPU oracle (through the C ABI).

Tolerances (float32 kernel vs float64 oracle), frozen from measurements in DESIGN.md "Parity":
  * before the free box lands (t < 4) nothing in the reference algorithm is ill-conditioned:
    theta / eef within 2e-6;
  * from the landing on, MJX's single-iteration Newton + 5-step bracketed line search returns
    whichever bracket end rounding noise favours; the float32 and float64 builds of
    the *oracle itself* then differ by up to ~5e-4 rad on contact-free samples, so that is the
    trajectory tolerance there; samples with robot contacts are compared per step (teacher forced)
    in test_gpu_step.py instead.
"""
import numpy as np
import pytest

from conftest import Q0, TARGET_POS, TARGET_ROT, planner_inputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def planner16():
    from manipulator_mujoco_b200 import cem_planner
    return cem_planner(num_dof=6, num_batch=100, num_steps=16, timestep=0.05, maxiter_cem=1, num_elite=0.05,
                       w_pos=20.0, w_rot=3.0, w_col=80.0, maxiter_projection=10)


def test_warm0_matches_oracle(planner16, oracle64):
    np.testing.assert_allclose(planner16.mjx_data["qacc"], oracle64.initial_warmstart(), atol=2e-5)


def test_rollout_c1_config(planner16, oracle64):
    """BASELINE config 1: B=100, T=16, order-10 Bernstein, planner-distributed samples."""
    pr, z, xi, st, xif, td = planner_inputs(16, 100)
    theta, eef_pos, eef_rot, col = [a.cpu().numpy() for a in planner16.compute_rollout_batch(td, Q0, np.zeros(6))]
    oth, oep, oer, ocol = oracle64.rollout(td, Q0, np.zeros(6))
    T = 16
    th = theta.reshape(100, 6, T); ot = oth.reshape(100, 6, T)
    assert np.abs(th[:, :, :4] - ot[:, :, :4]).max() < 2e-6
    assert np.abs(eef_pos[:, :5] - oep[:, :5]).max() < 2e-6
    assert np.abs(eef_rot[:, :5] - oer[:, :5]).max() < 2e-6
    free = ~(ocol < 0).any(axis=(1, 2))
    assert free.sum() >= 90
    assert np.abs(theta[free] - oth[free]).max() < 5e-4
    assert np.abs(eef_pos[free] - oep[free]).max() < 5e-4
    assert np.median(np.abs(theta[free] - oth[free]).max(axis=1)) < 5e-5
    # collision distances: identical branch decisions except within rounding of a switch
    dc = np.abs(col[free] - ocol[free])
    assert (dc > 1e-3).mean() < 1e-4
    # fused cost == reference cost formula on the oracle's trajectories
    cost, cg, cr, cc = [a.cpu().numpy() for a in planner16.compute_cost_batch(td, eef_pos, eef_rot, col,
                                                                             np.tile(TARGET_POS, (100, 1)), np.tile(TARGET_ROT, (100, 1)))]
    oc = pr.compute_cost_batch(oep, oer, ocol, TARGET_POS, TARGET_ROT)
    np.testing.assert_allclose(cg[free], oc[1][free], rtol=2e-4)
    np.testing.assert_allclose(cr[free], oc[2][free], rtol=2e-4, atol=1e-4)
    np.testing.assert_allclose(cc[free], oc[3][free], rtol=0, atol=5e-3)


def test_fused_cost_equals_standalone_cost(planner16):
    pr, z, xi, st, xif, td = planner_inputs(16, 100, seed=3)
    theta, cost4, ep, er, col = planner16._rollout(td, Q0, np.zeros(6), TARGET_POS, TARGET_ROT, True)
    c = planner16.compute_cost_batch(td, ep, er, col, np.tile(TARGET_POS, (100, 1)), np.tile(TARGET_ROT, (100, 1)))
    for k in range(4):
        np.testing.assert_allclose(cost4[:, k].cpu().numpy(), c[k].cpu().numpy(), rtol=1e-5, atol=1e-5)


def test_results_do_not_depend_on_batch_size_or_cta_variant():
    """Samples are independent: the same sample must give bit-identical results whether it runs in a
    batch of 100, 1000 or 2300 (different kernel instantiations, CTA shares and a partly filled last CTA)."""
    from manipulator_mujoco_b200 import cem_planner
    T = 24
    pr, z, xi, st, xif, td = planner_inputs(T, 2300, seed=9)
    outs = {}
    for B in (100, 1000, 2300):
        pl = cem_planner(num_dof=6, num_batch=B, num_steps=T, timestep=0.05, maxiter_cem=1, num_elite=0.05,
                         w_pos=20.0, w_rot=3.0, w_col=80.0, maxiter_projection=10)
        theta, cost4, ep, er, col = pl._rollout(td[:B], Q0, np.zeros(6), TARGET_POS, TARGET_ROT, True)
        outs[B] = (theta.cpu().numpy(), cost4.cpu().numpy(), col.cpu().numpy())
    for B in (1000, 2300):
        for a, b in zip(outs[100], outs[B]):
            np.testing.assert_array_equal(a, b[:100])
    np.testing.assert_array_equal(outs[1000][1], outs[2300][1][:1000])


def test_results_do_not_depend_on_the_partner_sample_or_odd_batch_sizes():
    """Two samples share a warp (16-lane groups).  A sample's result must not depend on which sample it is
    paired with, on which half of the warp it runs, or on partly filled warps / CTAs (odd batch sizes)."""
    from manipulator_mujoco_b200 import cem_planner
    T, B = 40, 301
    pr, z, xi, st, xif, td = planner_inputs(T, B, seed=21)
    td = np.ascontiguousarray(td, dtype=np.float32)

    def run(rows):
        pl = cem_planner(num_dof=6, num_batch=len(rows), num_steps=T, timestep=0.05, maxiter_cem=1, num_elite=0.05,
                         w_pos=20.0, w_rot=3.0, w_col=80.0, maxiter_projection=10)
        theta, cost4, ep, er, col = pl._rollout(td[rows], Q0, np.zeros(6), TARGET_POS, TARGET_ROT, True)
        return theta.cpu().numpy(), cost4.cpu().numpy(), col.cpu().numpy()

    base = run(np.arange(B))
    rev = run(np.arange(B)[::-1].copy())                       # other partner, other half of the warp
    for a, b in zip(base, rev):
        np.testing.assert_array_equal(a, b[::-1])
    rng = np.random.default_rng(0)
    perm = rng.permutation(B)
    shuf = run(perm)
    for a, b in zip(base, shuf):
        np.testing.assert_array_equal(a[perm], b)
    for n in (1, 2, 3, 29, 57):                                # single group, odd counts, one full CTA + 1
        sub = run(np.arange(n))
        for a, b in zip(base, sub):
            np.testing.assert_array_equal(a[:n], b)


def test_multi_wave_cta_shares_cover_every_sample_once():
    """More samples than one wave of CTAs holds (148 SMs x 28): the per-SM share is split over waves with two
    different CTA shares (cemk_rollout_cost).  Every sample must be rolled out exactly once, with the same
    result as in a small batch."""
    from manipulator_mujoco_b200 import cem_planner
    T, B = 10, 9000
    pr, z, xi, st, xif, td = planner_inputs(T, 600, seed=33)
    td = np.ascontiguousarray(np.tile(td, (B // 600, 1)), dtype=np.float32)        # 15 copies of 600 distinct samples
    pl = cem_planner(num_dof=6, num_batch=B, num_steps=T, timestep=0.05, maxiter_cem=1, num_elite=0.05,
                     w_pos=20.0, w_rot=3.0, w_col=80.0, maxiter_projection=10)
    theta, cost4, _, _, _ = pl._rollout(td, Q0, np.zeros(6), TARGET_POS, TARGET_ROT, False)
    theta, cost4 = theta.cpu().numpy(), cost4.cpu().numpy()
    assert np.isfinite(cost4).all()
    for rep in range(1, B // 600):                                                 # all copies agree bit for bit
        np.testing.assert_array_equal(theta[:600], theta[rep * 600:(rep + 1) * 600])
        np.testing.assert_array_equal(cost4[:600], cost4[rep * 600:(rep + 1) * 600])
    small = cem_planner(num_dof=6, num_batch=600, num_steps=T, timestep=0.05, maxiter_cem=1, num_elite=0.05,
                        w_pos=20.0, w_rot=3.0, w_col=80.0, maxiter_projection=10)
    th2, c2, _, _, _ = small._rollout(td[:600], Q0, np.zeros(6), TARGET_POS, TARGET_ROT, False)
    np.testing.assert_array_equal(theta[:600], th2.cpu().numpy())
    np.testing.assert_array_equal(cost4[:600], c2.cpu().numpy())


def test_big_capacity_kernel_agrees_with_fast_kernel():
    """`force_rerun` recomputes every sample with the 48-contact all-in-shared-memory instantiation; the fast
    kernel (20 contacts in shared memory, the rest in the global spill area) must agree bit for bit -- also
    from start configurations that lie deep inside the table (more than 20 simultaneous contacts)."""
    from manipulator_mujoco_b200 import _lib, cem_planner
    T, B = 60, 256
    pl = cem_planner(num_dof=6, num_batch=B, num_steps=T, timestep=0.05, maxiter_cem=1, num_elite=0.05,
                     w_pos=20.0, w_rot=3.0, w_col=80.0, maxiter_projection=10)
    pr, z, xi, st, xif, td = planner_inputs(T, B, seed=11)
    a = pl._rollout(td, Q0, np.zeros(6), TARGET_POS, TARGET_ROT, True)
    assert int(pl._buf("flags", (B,), __import__("torch").int32).max()) == 0      # nothing overflowed the fast kernel's contact capacity
    _lib.check(pl._lib.cemk_set_option(pl._h, b"force_rerun", 1), pl._lib)
    b = pl._rollout(td, Q0, np.zeros(6), TARGET_POS, TARGET_ROT, True)
    _lib.check(pl._lib.cemk_set_option(pl._h, b"force_rerun", 0), pl._lib)
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x.cpu().numpy(), y.cpu().numpy())
    for q_deep in ([1.181, 0.789, -0.949, -0.051, 2.391, -0.535], [1.168, 0.758, -0.921, -0.796, 0.168, 2.145]):
        a = pl._rollout(td * 0.2, np.array(q_deep), np.zeros(6), TARGET_POS, TARGET_ROT, True)
        ncol = int((a[4][:, 0].cpu().numpy() < 0).sum(axis=1).max())
        assert ncol > 16                                           # robot slots alone; the box adds 4 after it lands
        _lib.check(pl._lib.cemk_set_option(pl._h, b"force_rerun", 1), pl._lib)
        b = pl._rollout(td * 0.2, np.array(q_deep), np.zeros(6), TARGET_POS, TARGET_ROT, True)
        _lib.check(pl._lib.cemk_set_option(pl._h, b"force_rerun", 0), pl._lib)
        for x, y in zip(a, b):
            np.testing.assert_array_equal(x.cpu().numpy().view(np.int32), y.cpu().numpy().view(np.int32))
