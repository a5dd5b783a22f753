"""oracle/production_mixed.py (DG2 x RT3 x CG3 of advection_diffusion_mixed.py): no recorded output exists for it
(parity unpinned), so the checks are internal: the element Jacobian is the derivative of the residual, the
discrete spaces have the expected dimension, and the total bottom flux approaches the analytic reaction rate
4 pi D1 sigma c/(gamma + c) that the reference prints next to it (:264-267)."""
import numpy as np

from oracle import fem
from oracle.production_mixed import ProductionMixed

SIGMA, GAMMA, D1, D2 = 0.1, 1.0, 2.0, 5.5
V1 = 4 * np.pi * (5.0 - SIGMA) ** 3 / 3


def _problem(n, c, dt=None):
    m = fem.mapped_rectangle_mesh(n, n, 5.0, 5.0, SIGMA, GAMMA)
    fc = m.facets()
    a, b = fc.verts[:, 0].astype(np.int64), fc.verts[:, 1].astype(np.int64)
    ax, ay, bx, by = a % (n + 1), a // (n + 1), b % (n + 1), b // (n + 1)
    ob = fc.cell1 < 0
    mk = np.zeros(fc.verts.shape[0], dtype=np.int64)
    mk[ob & (ax == n) & (bx == n)] = 21
    mk[ob & (ay == n) & (by == n)] = 22
    mk[ob & (ax == 0) & (bx == 0)] = 23
    mk[ob & (ay == 0) & (by == 0)] = 24
    return ProductionMixed(m, mk, c, SIGMA, GAMMA, V1, D1, D2, dt)


def test_dimension_and_jacobian_is_derivative_of_residual():
    S = _problem(4, 2.0, dt=0.1)
    assert S.N == 6 * S.C + (3 * S.E + 6 * S.C) + (S.V + 2 * S.E + S.C)          # DG2 + RT3 + CG3
    rng = np.random.default_rng(0)
    # a state on the nonlinear branch of Gstar: |p/(4 c pi) - q| > 1e-3
    u = np.concatenate([0.5 * (1 + rng.random(S.np_)), 0.1 * rng.standard_normal(S.ns), 0.2 * (1 + rng.random(S.nq_))])
    po = 0.3 * (1 + rng.random(S.N))
    J, b = S.assemble(u, True, po)
    d = rng.standard_normal(S.N)
    eps = 1e-6
    _, bp = S.assemble(u + eps * d, False, po)
    _, bm = S.assemble(u - eps * d, False, po)
    fd = (bp - bm) / (2 * eps)
    assert np.linalg.norm(J @ d - fd) < 1e-7 * np.linalg.norm(fd)


def test_total_flux_approaches_the_analytic_rate():
    S = _problem(24, 1.0)
    u = S.solve()
    assert S.converged
    r2, r1, tot = S.bottom_flux(S.flux(u))
    approx = 4 * np.pi * D1 * SIGMA * 1.0 / (GAMMA + 1.0)
    assert abs(tot - 1.26643) < 2e-4 and abs(tot / approx - 1) < 0.05 and abs(r2 + r1 - tot) < 1e-12
