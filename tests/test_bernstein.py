"""Basis matrices vs golden values produced by the reference's own module (tools/make_golden.py)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from manipulator_mujoco_b200.bernstein import bernstein_coeff_ordern_new
from oracle.planner_ref import bernstein_coeff_ordern

G = np.load(os.path.join(GOLDEN, "bernstein.npz"))


@pytest.mark.parametrize("T,dt", [(16, 0.05), (100, 0.05), (10, 0.04), (50, 0.05)])
def test_matches_reference_golden(T, dt):
    tt = np.linspace(0, T * dt, T).reshape(T, 1)               # mjx_planner.py:36-38
    for fn in (bernstein_coeff_ordern_new, bernstein_coeff_ordern):
        P, Pd, Pdd = fn(10, tt[0], tt[-1], tt)
        for name, a in (("P", P), ("Pdot", Pd), ("Pddot", Pdd)):
            ref = G[f"{name}_{T}_{dt}"]
            assert a.shape == ref.shape == (T, 11)
            np.testing.assert_allclose(a, ref, rtol=0, atol=1e-12 * max(1.0, np.abs(ref).max()))
            # the hand-expanded order-10 file of the reference agrees with its order-n file
            np.testing.assert_allclose(G[f"{name}10_{T}_{dt}"], ref, rtol=0, atol=2e-13 * max(1.0, np.abs(ref).max()))


def test_known_answers():
    """SURVEY.md section 4 KATs: partition of unity, end-point interpolation, first rows at T=16."""
    T, dt = 16, 0.05
    tt = np.linspace(0, T * dt, T).reshape(T, 1)
    P, Pd, Pdd = bernstein_coeff_ordern_new(10, tt[0], tt[-1], tt)
    np.testing.assert_allclose(P.sum(1), 1.0, atol=1e-14)
    np.testing.assert_allclose(Pd.sum(1), 0.0, atol=1e-11)
    np.testing.assert_allclose(P[0], np.eye(11)[0], atol=1e-15)
    np.testing.assert_allclose(P[-1], np.eye(11)[10], atol=1e-15)
    np.testing.assert_allclose(Pd[0, :3], [-12.5, 12.5, 0.0], atol=1e-12)
    np.testing.assert_allclose(Pdd[0, :3], [140.625, -281.25, 140.625], atol=1e-10)
    # the grid spacing is num*t/(num-1), not t (mjx_planner.py:36)
    assert abs((tt[1] - tt[0])[0] - 16 * 0.05 / 15) < 1e-15


@pytest.mark.parametrize("n", [5, 8, 12, 15])
def test_other_orders_match_reference_golden(n):
    """bernstein_coeff_ordern_new(n, ...) of the reference for the orders of the order-n extension (cem_planner(bernstein_order=n))."""
    T, dt = 16, 0.05
    tt = np.linspace(0, T * dt, T).reshape(T, 1)
    for fn in (bernstein_coeff_ordern_new, bernstein_coeff_ordern):
        P, Pd, Pdd = fn(n, tt[0], tt[-1], tt)
        for name, a in (("P", P), ("Pdot", Pd), ("Pddot", Pdd)):
            ref = G[f"{name}_n{n}"]
            assert a.shape == ref.shape == (T, n + 1)
            np.testing.assert_allclose(a, ref, rtol=0, atol=1e-12 * max(1.0, np.abs(ref).max()))
