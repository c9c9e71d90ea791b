"""Host-side algebra: the structured per-DOF projection used by k_project equals the reference's
dense iteration (oracle/planner_ref.py restates mjx_planner.py:181-249 literally)."""
import numpy as np

from conftest import Q0
from oracle.planner_ref import PlannerRef


def structured_projection(pr, xi, state_term, iters):
    """numpy mirror of csrc/cemk.cu:k_project (float64)."""
    n1, nd, nv = pr.nvar_single, pr.num_dof, pr.nvar
    Kpp, Kpe = pr.Q_inv[:n1, :n1], pr.Q_inv[:n1, nv:nv + 5]
    G = [pr.Pdot, pr.Pddot, pr.P]
    bnd = [pr.v_max, pr.a_max, pr.p_max]
    B = xi.shape[0]
    out = np.zeros_like(xi)
    for d in range(nd):
        xs = xi[:, d * n1:(d + 1) * n1]
        beq = state_term.reshape(B, 5, nd)[:, :, d]
        lam = np.zeros((B, n1)); rh = np.zeros((B, n1))
        for _ in range(iters):
            x = (lam + xs + rh) @ Kpp.T + beq @ Kpe.T
            u = [x @ g.T for g in G]
            cl = [np.clip(ui, -b, b) for ui, b in zip(u, bnd)]
            rh = sum((ui + ci) @ g for ui, ci, g in zip(u, cl, G))
            lam = lam - sum((ui - ci) @ g for ui, ci, g in zip(u, cl, G))
        out[:, d * n1:(d + 1) * n1] = x
    return out


def test_q_inv_is_block_diagonal_per_dof():
    pr = PlannerRef(6, 8, 16, 0.05, 0.05, 20, 3, 80, 10)
    n1, nv = 11, 66
    scale = np.abs(pr.Q_inv).max()
    for d in range(6):
        blk = pr.Q_inv[d * n1:(d + 1) * n1]
        assert np.abs(blk[:, d * n1:(d + 1) * n1] - pr.Q_inv[:n1, :n1]).max() < 1e-7 * scale
        mask = np.ones(96, bool); mask[d * n1:(d + 1) * n1] = False; mask[nv + 5 * d:nv + 5 * d + 5] = False
        assert np.abs(blk[:, mask]).max() < 1e-7 * scale


def test_structured_projection_equals_dense_reference():
    for T in (16, 50):
        pr = PlannerRef(6, 32, T, 0.05, 0.05, 20, 3, 80, 10)
        rng = np.random.default_rng(0)
        xi = rng.normal(size=(32, 66)) * np.sqrt(10)
        st = pr.state_term(Q0, rng.normal(size=6) * 0.1, rng.normal(size=6) * 0.1, 32)
        dense = pr.compute_projection_filter(xi, st)
        fast = structured_projection(pr, xi, st, 10)
        np.testing.assert_allclose(fast, dense, rtol=0, atol=2e-8 * max(1.0, np.abs(dense).max()))
        # the filter enforces the boundary conditions exactly and shrinks bound violations
        np.testing.assert_allclose(dense @ pr.A_eq.T, pr.boundary_vec(st), atol=1e-6)
        v_raw = np.abs(xi @ pr.A_thetadot.T).max()
        v_fil = np.abs(dense @ pr.A_thetadot.T).max()
        assert v_fil < 0.5 * v_raw


def test_mean_cov_and_elites_reference_properties():
    pr = PlannerRef(6, 100, 16, 0.05, 0.05, 20, 3, 80, 10)
    rng = np.random.default_rng(1)
    cost = rng.uniform(1, 5, 100); cost[7] = cost[3]
    xi = rng.normal(size=(100, 66))
    xe, idx, ce = pr.compute_ellite_samples(cost, xi)
    assert len(ce) == 5 and np.all(np.diff(ce) >= 0)
    assert list(idx).index(3) < list(idx).index(7)               # stable: ties keep the lower index first
    mean, cov = pr.compute_mean_cov(ce, np.zeros(66), 10 * np.eye(66), xe)
    np.testing.assert_allclose(cov, cov.T, atol=1e-12)
    assert np.linalg.eigvalsh(cov).min() > 0


def test_vectorised_cost_and_float32_variant_match_the_literal_restatement():
    """bench.py's CPU arm times compute_cost_batch_vec and the float32 copy of the constants; both must be the same
    arithmetic as the line-by-line float64 restatement."""
    rng = np.random.default_rng(3)
    B, T, ns = 12, 16, 187
    pr = PlannerRef(6, B, T, 0.05, 0.25, 20, 3, 80, 10)
    ep, er = rng.normal(size=(B, T, 3)), rng.normal(size=(B, T, 4))
    col = rng.uniform(-0.05, 1.0, size=(B, T, ns))
    col[rng.uniform(size=col.shape) < 0.5] = 1.0                     # sentinel slots
    tp, tr = np.array([-0.3, -0.3, 0.5]), np.array([0.0, 1.0, 0.0, 0.0])
    a = pr.compute_cost_batch(ep, er, col, tp, tr)
    b = pr.compute_cost_batch_vec(ep, er, col, tp, tr)
    for x, y in zip(a, b):
        np.testing.assert_allclose(x, y, rtol=1e-12, atol=1e-12)
    pr32 = pr.astype(np.float32)
    assert pr32.Q_inv.dtype == np.float32 and pr.Q_inv.dtype == np.float64
    xi = rng.normal(size=(B, 66)) * 3
    st = pr.state_term(Q0, np.zeros(6), np.zeros(6), B)
    x64 = pr.compute_projection_filter(xi, st)
    x32 = pr32.compute_projection_filter(xi.astype(np.float32), st.astype(np.float32))
    assert x32.dtype == np.float32
    np.testing.assert_allclose(x32, x64, atol=5e-5)
