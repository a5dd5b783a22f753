"""The oracle against the reference's recorded outputs: README.md:109-115."""
import numpy as np

from oracle import fem, pq, scalar

import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _golden(name):
    return [ln.rstrip("\n") for ln in open(os.path.join(GOLDEN, name)) if ln.strip()]


# README.md:109-115 of the reference, verbatim rows (DoF, h, e_1(u), r_1(u), e_0(u), r_0(u));
# tests/golden/readme_table.txt holds the same rows as a file
README_TABLE = [
    "     9  0.7071 1.81e+00  0.000  1.74e-01  0.000 ",
    "    25  0.3536 9.23e-01  0.968  3.52e-02  2.305 ",
    "    81  0.1768 4.45e-01  1.050  7.51e-03  2.228 ",
    "   289  0.0884 2.20e-01  1.021  1.75e-03  2.102 ",
    "  1089  0.0442 1.09e-01  1.007  4.29e-04  2.030 ",
    "  4225  0.0221 5.46e-02  1.001  1.09e-04  1.971 ",
    " 16641  0.0110 2.73e-02  0.998  2.95e-05  1.891 ",
]


def test_golden_files_match_embedded_tables():
    assert [r.split() for r in _golden("readme_table.txt")] == [r.split() for r in README_TABLE]
    assert [r.split() for r in _golden("paper_k0_table.txt")] == [r.split() for r in PAPER_K0]


def test_readme_table_digit_for_digit():
    rows = scalar.convergence_table(scalar.solve_fb, 7)
    assert [r.split() for r in rows] == [r.split() for r in README_TABLE]


def test_quadrature_rules_integrate_exactly():
    # degree-d rule integrates x^a y^b, a+b<=d, on the reference triangle: a! b!/(a+b+2)!
    from math import factorial
    for deg in (1, 2, 4, 6, 8, 12):
        pts, w = fem.triangle_rule(deg)
        assert abs(w.sum() - 0.5) < 1e-14
        for a in range(deg + 1):
            for b in range(deg + 1 - a):
                ex = factorial(a) * factorial(b) / factorial(a + b + 2)
                assert abs((w * pts[:, 0] ** a * pts[:, 1] ** b).sum() - ex) < 2e-14, (deg, a, b)


def test_rectangle_mesh_dolfin_numbering():
    m = fem.rectangle_mesh(0.0, 0.0, 1.0, 1.0, 2, 2)
    assert m.cells[:4].tolist() == [[0, 1, 4], [0, 3, 4], [1, 2, 5], [1, 4, 5]]
    assert abs(m.hmax() - np.sqrt(2) / 2) < 1e-15
    fc = m.facets()
    assert fc.verts.shape[0] == 16 and fc.exterior.size == 8
    rp, col = fem.p1_node_pattern(m)
    assert rp[-1] == 9 + 2 * 16


def test_laplace_rates():
    # config 1 (SphericalLaplace_convergence.py) has no recorded table: expected orders only
    r = [scalar.solve_fa(n) for n in (16, 32, 64)]
    r1 = np.log(r[2]["e1"] / r[1]["e1"]) / np.log(0.5)
    r0 = np.log(r[2]["e0"] / r[1]["e0"]) / np.log(0.5)
    assert 0.95 < r1 < 1.05 and 1.8 < r0 < 2.2


def _fd_jac(S, u, **kw):
    import scipy.sparse as sp
    vals, _ = S.assemble(u, True, **kw)
    J = sp.csr_matrix((vals, S.col, S.rowptr), shape=(S.N, S.N)).toarray()
    Jfd = np.zeros_like(J)
    for k in range(S.N):
        up, um = u.copy(), u.copy()
        up[k] += 1e-6
        um[k] -= 1e-6
        Jfd[:, k] = (S.assemble(up, False, **kw)[1] - S.assemble(um, False, **kw)[1]) / 2e-6
    return np.abs(J - Jfd).max() / np.abs(J).max()


def test_pq_jacobian_matches_finite_differences():
    mesh = fem.rectangle_mesh(0, 0, 1, 1, 3, 4)
    rng = np.random.default_rng(0)
    V = mesh.num_vertices
    u = np.empty(2 * V)
    u[0::2] = 0.3 + rng.random(V)
    u[1::2] = 1.0 + rng.random(V)
    fc = mesh.facets()
    variants = [
        (pq.PQParams(D1=1e-3, D2=1.5e-3, gkind=pq.G_POWER, gcoef=1.5e-3 * 4 * np.pi, q_react=1.0,
                     inv_dt=10.0), False),
        (pq.PQParams(**pq.gstar_conv(1.0)), False),
        (pq.PQParams(D1=1e-3, D2=1.5e-3, gkind=pq.G_POWER, gcoef=1.5e-3 * 4 * np.pi, q_react=1.0,
                     q_x2p=0, q_x2q=1, supg_stab=1.0), False),
        (pq.PQParams(D1=1, D2=1.5, gkind=pq.G_POWER, gcoef=1.0, q_adv=0, q_divr=1, q_ibp=1,
                     c0ip_gamma=0.05, c0ip_h=1, pen_stab2=1.0, degree=6), True),
    ]
    for prm, ext in variants:
        S = pq.PQSystem(mesh, prm, fc.exterior if ext else None,
                        np.full(fc.exterior.size, 7) if ext else None)
        err = _fd_jac(S, u, p_old=rng.random(V), q_data=pq.q_exact)
        assert err < 1e-8, (prm, err)


def test_pq_manufactured_rates():
    from oracle.newton import merge_bcs, newton_solve
    prm = pq.PQParams(D1=1e-3, D2=1.5e-3, gkind=pq.G_POWER, gcoef=1.5e-3 * 4 * np.pi, q_react=1.0)
    forcing = pq.manufactured_forcing(prm)
    res = []
    for n in (16, 32):
        mesh = fem.rectangle_mesh(0, 0, 1, 1, n, n)
        S = pq.PQSystem(mesh, prm)
        fc = mesh.facets()
        bv = np.unique(fc.verts[fc.exterior])
        X = mesh.coords
        qv = np.nonzero(fem.near(X[:, 0], 1.0))[0]
        dofs, vals = merge_bcs([(2 * bv, pq.p_exact(X[bv, 0], X[bv, 1])),
                                (2 * qv + 1, pq.q_exact(X[qv, 0], X[qv, 1]))])
        x, it, conv = newton_solve(lambda x, wj: S.assemble(x, wj, None, forcing), np.zeros(S.N),
                                   S.rowptr, S.col, dofs, vals, atol=1e-7, rtol=1e-7)
        assert conv and it <= 4
        res.append(pq.errors_pq(mesh, x))
    rates = [np.log(a / b) / np.log(0.5) for a, b in zip(res[1], res[0])]
    assert rates[0] > 1.9 and 0.95 < rates[1] < 1.05 and 0.95 < rates[2] < 1.1


def test_c_port_matches_numpy_oracle():
    from oracle import cport
    mesh = fem.rectangle_mesh(0, 0.1, 5, 5, 12, 9)
    rng = np.random.default_rng(5)
    V = mesh.num_vertices
    u = np.empty(2 * V)
    u[0::2] = 0.3 + rng.random(V)
    u[1::2] = 1.0 + rng.random(V)
    pold = rng.random(V)
    for prm in (pq.PQParams(**pq.gstar_conv(1.0)),
                pq.PQParams(D1=1e-3, D2=1.5e-3, gkind=pq.G_POWER, gcoef=0.02, q_react=1.0, q_x2p=0,
                            q_x2q=1, supg_stab=1.0, inv_dt=10.0),
                pq.PQParams(D1=1, D2=1.5, gkind=pq.G_POWER, gcoef=1.0, q_adv=0, q_divr=1, q_ibp=1)):
        S = pq.PQSystem(mesh, prm)
        ovals, ob = S.assemble(u, True, pold)
        vals, b = cport.assemble_cells(prm, mesh.cells, mesh.coords, u, S.cell_slots, S.col.size, pold)
        assert np.max(np.abs(vals - ovals)) / np.max(np.abs(ovals)) < 1e-13
        assert np.max(np.abs(b - ob)) / np.max(np.abs(ob)) < 1e-13


# convergence/test_for_paper_convergence.py:200-207, verbatim (k=0 table)
PAPER_K0 = [
    "    33 & 0.7071 & 2.93e+01 & 0.000 & 1.03e+00 & 0.000 & 4.92e+00 & 0.000 ",
    "   113 & 0.3536 & 1.73e+01 & 0.763 & 5.56e-01 & 0.888 & 7.08e+00 & -0.526 ",
    "   417 & 0.1768 & 9.96e+00 & 0.793 & 2.84e-01 & 0.969 & 1.84e+00 & 1.943 ",
    "  1601 & 0.0884 & 5.93e+00 & 0.747 & 1.43e-01 & 0.992 & 7.10e-01 & 1.375 ",
    "  6273 & 0.0442 & 3.72e+00 & 0.673 & 7.16e-02 & 0.997 & 3.35e-01 & 1.085 ",
    " 24833 & 0.0221 & 2.45e+00 & 0.604 & 3.59e-02 & 0.995 & 1.63e-01 & 1.036 ",
    " 98817 & 0.0110 & 1.67e+00 & 0.551 & 1.82e-02 & 0.981 & 8.13e-02 & 1.004 ",
]


def test_paper_k0_table_pins_facet_terms():
    """RT1xDG0xCG1 system whose q-block facet tensors come from oracle.pq's facet functions: rows from
    the third level on match the recorded table digit for digit; the two coarsest differ in the last
    digit only (analytic data instead of FEniCS's degree-5/6 interpolants)."""
    from oracle import paper_m0
    rows = paper_m0.table(6)
    got = [[float(t) for t in r.replace("&", " ").split()] for r in rows]
    ref = [[float(t) for t in r.replace("&", " ").split()] for r in PAPER_K0[:6]]
    for k in range(2, 6):
        assert rows[k].split() == PAPER_K0[k].split(), (rows[k], PAPER_K0[k])
    for k in range(2):
        assert got[k][0] == ref[k][0] and got[k][1] == ref[k][1]
        for a, b in zip(got[k][2::2], ref[k][2::2]):          # error columns: <= 1 in the 3rd digit
            assert abs(a - b) <= 0.011 * 10 ** np.floor(np.log10(b)), (a, b)


# convergence/test_for_paper_convergence.py:211-217, verbatim (first k=1 table: RT2 x DG1 x CG2)
PAPER_K1 = _golden("paper_k1_table.txt")


def test_paper_k1_table_rt2_dg1_cg2():
    """oracle/paper_m1.py (RT2 x DG1 x CG2, stab = 0.1 as the comment at :70 says, stab2 = 2) against the first
    recorded k=1 table: rows from the third level on digit for digit, the two coarsest to the last printed digit
    of e_1(q) (analytic data instead of degree-5/6 interpolants).  DoF column = 3E + 5C + V.
    The second recorded k=1 table (:219-225) has the same e_div(s) / e_0(p) columns and a different e_1(q): its
    stabilisation parameters are not recorded in the script."""
    from oracle import paper_m1
    levels = 5
    rows = paper_m1.table(levels)
    got = [[float(t) for t in r.replace("&", " ").split()] for r in rows]
    ref = [[float(t) for t in r.replace("&", " ").split()] for r in PAPER_K1[:levels]]
    for k in range(2, levels):
        assert rows[k].split() == PAPER_K1[k].split(), (rows[k], PAPER_K1[k])
    for k in range(2):
        assert got[k][:6] == ref[k][:6], (rows[k], PAPER_K1[k])          # DoF, h, e_div(s), r, e_0(p), r exactly
        assert abs(got[k][6] - ref[k][6]) <= 0.011 * 10 ** np.floor(np.log10(ref[k][6])), (got[k], ref[k])
