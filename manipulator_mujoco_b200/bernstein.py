"""Bernstein basis on an arbitrary interval (host-side constant generation, float64).

Same quantities as the reference's ``bernstein_coeff_ordern_new``
(``sampling_based_planner/bernstein_coeff_ordern_arbitinterval.py:4-28``), called from
``cem_planner.__init__`` at ``mjx_planner.py:40``: P, Pdot, Pddot of shape [len(t), n+1].
Written with the derivative identity  B'_{i,n} = n (B_{i-1,n-1} - B_{i,n-1})  applied once and
twice instead of the reference's expanded power expressions; tests/test_bernstein.py pins the
values against golden matrices produced by the reference module itself.
"""
from __future__ import annotations

from math import comb

import numpy as np


def _basis(n, s):
    """B_{i,n}(s) for i = 0..n as columns; empty-degree guard for n < 0."""
    if n < 0:
        return np.zeros((s.size, 0))
    i = np.arange(n + 1)
    binom = np.array([comb(n, k) for k in i], dtype=np.float64)
    return binom * np.power(1.0 - s[:, None], n - i) * np.power(s[:, None], i)


def _pad(b):
    """[0 | b | 0] so that column i-1 / i lookups of the lower-degree basis are total."""
    z = np.zeros((b.shape[0], 1))
    return np.hstack((z, b, z))


def bernstein_coeff_ordern_new(n, tmin, tmax, t_actual):
    t_actual = np.asarray(t_actual, dtype=np.float64).reshape(-1)
    l = float(np.asarray(tmax).reshape(-1)[0] - np.asarray(tmin).reshape(-1)[0])
    s = (t_actual - float(np.asarray(tmin).reshape(-1)[0])) / l
    P = _basis(n, s)
    b1 = _pad(_basis(n - 1, s))                       # b1[:, i] = B_{i-1,n-1}
    Pdot = n * (b1[:, :-1] - b1[:, 1:]) / l
    b2 = _pad(_pad(_basis(n - 2, s)))                 # b2[:, i] = B_{i-2,n-2}
    Pddot = n * (n - 1) * (b2[:, :-2] - 2.0 * b2[:, 1:-1] + b2[:, 2:]) / (l * l)
    return P, Pdot, Pddot
