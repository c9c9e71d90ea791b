import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
Q0 = np.array([1.5, -1.8, 1.75, -1.25, -1.6, 0.0])        # mjx_planner.py:367 / run_mpc_planner.py:26
TARGET_POS = np.array([-0.3, -0.3, 0.5])                   # objects.xml:6 (target_0)
TARGET_ROT = np.array([0.0, 1.0, 0.0, 0.0])


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def mc():
    from manipulator_mujoco_b200.mjcf import load_model
    return load_model()


@pytest.fixture(scope="session")
def oracle64(mc):
    from oracle.oracle import Oracle
    return Oracle(mc, 0.05, dtype="f64")


@pytest.fixture(scope="session")
def oracle32(mc):
    from oracle.oracle import Oracle
    return Oracle(mc, 0.05, dtype="f32")


def planner_inputs(T, B, seed=0, maxiter_projection=10):
    """Planner-distributed rollout inputs: xi ~ N(0, 10 I) -> projection filter -> thetadot (float32-exact)."""
    from oracle.planner_ref import PlannerRef
    pr = PlannerRef(6, B, T, 0.05, 0.05, 20.0, 3.0, 80.0, maxiter_projection)
    rng = np.random.default_rng(seed)
    z = rng.normal(size=(B, 66)).astype(np.float32).astype(np.float64)
    xi = pr.compute_xi_samples(z, np.zeros(66), 10 * np.identity(66)).astype(np.float32).astype(np.float64)
    st = pr.state_term(Q0, np.zeros(6), np.zeros(6), B)
    xif = pr.compute_projection_filter(xi, st)
    td = (xif @ pr.A_thetadot.T).astype(np.float32).astype(np.float64)
    return pr, z, xi, st, xif, td
