"""The RT2 x DG1 x CG2 restatements of the MixedPrimal script family (oracle/paper_m1.py) have no recorded table in
the reference (parity unpinned), so they are checked for internal consistency here: the assembled Jacobian is the
derivative of the assembled residual (central differences), the discrete problem is consistent with the manufactured
solution (p converges with order 2 on every variant) and each variant shows the order in q its stabilisation gives."""
import numpy as np
import pytest

from oracle import paper_m1

VARIANTS = {
    "galerkin": lambda n: paper_m1.MixedPrimalM1(n),
    "supg": lambda n: paper_m1.MixedPrimalM1(n, supg=True),
    "c0ip": lambda n: paper_m1.MixedPrimalC0IPM1(n),
    "c0ip_other": lambda n: paper_m1.MixedPrimalC0IPOtherM1(n),
    "c0ip_modified": lambda n: paper_m1.MixedPrimalC0IPModifiedM1(n),
}


@pytest.mark.parametrize("name", sorted(VARIANTS))
def test_jacobian_is_the_derivative_of_the_residual(name):
    m = VARIANTS[name](3)
    rng = np.random.default_rng(5)
    u = np.concatenate([0.3 * rng.standard_normal(2 * m.E + 2 * m.C), 0.5 + rng.random(3 * m.C),
                        1.0 + rng.random(m.V + m.E)])
    J, _ = m.assemble(u, True)
    d = rng.standard_normal(m.N)
    eps = 1e-6
    _, bp = m.assemble(u + eps * d, False)
    bp = bp.copy()
    _, bm = m.assemble(u - eps * d, False)
    fd = (bp - bm) / (2 * eps)
    assert np.abs(J @ d - fd).max() < 1e-7 * np.abs(fd).max()


@pytest.mark.parametrize("name,q_order", [("galerkin", (0.8, 1.3)), ("supg", (1.8, 2.6)), ("c0ip", (1.0, 1.8)),
                                          ("c0ip_other", (1.5, 1.9))])
def test_orders_of_convergence(name, q_order):
    errs, hs = [], []
    for n in (4, 8, 16):
        m = VARIANTS[name](n)
        errs.append(m.errors(m.solve()))
        hs.append(m.mesh.hmax())
    rate = lambda k: np.log(errs[2][k] / errs[1][k]) / np.log(hs[2] / hs[1])
    assert 1.8 < rate(1) < 2.1                     # e_0(p)
    # e_div(s): 1.5 where the flux is resolved; the C0IP script's strong coupling G = r2^2 p^2/q (gcoef = 1 against
    # D2 4 pi = 0.019 in the Galerkin / SUPG scripts) lets the q error dominate the flux error, whose order decays
    assert (0.4 if name == "c0ip" else 0.9) < rate(0) < 1.8
    assert q_order[0] < rate(2) < q_order[1]       # e_1(q)
