"""Closed-form checks of the oracle's flux post-processing (oracle/functionals.py)."""
import numpy as np

from oracle import fem, functionals as ofn


def _mesh():
    return fem.rectangle_mesh(0.0, 0.1, 5.0, 5.0, 12, 9)


def test_dg0_flux_of_affine_state_with_constant_ratio():
    """p = a + b x + c y, q = p/k  =>  flux = (D2 (b + k mean(x^2 p)), D1 c): the quotient term is a cubic,
    integrated exactly by the degree-5 rule, so the cell mean equals the exact one."""
    m = _mesh()
    X = m.coords
    a, b, c, k, D1, D2 = 0.7, 0.2, -0.1, 3.0, 2.0, 1.5
    p = a + b * X[:, 0] + c * X[:, 1]
    u = np.empty(2 * m.num_vertices)
    u[0::2], u[1::2] = p, p / k
    F = ofn.project_flux_dg0(m, u, D1, D2)
    # exact cell mean of k x^2 p with a high-order rule
    pts, wts = fem.triangle_rule(8)
    phi = np.stack([1 - pts[:, 0] - pts[:, 1], pts[:, 0], pts[:, 1]], 1)
    Xc = X[m.cells]
    xq = np.einsum("qi,ci->cq", phi, Xc[:, :, 0])
    yq = np.einsum("qi,ci->cq", phi, Xc[:, :, 1])
    mean = (k * xq ** 2 * (a + b * xq + c * yq) * wts).sum(1) / wts.sum()
    assert np.allclose(F[:, 0], D2 * (b + mean), rtol=1e-12, atol=1e-13)
    assert np.allclose(F[:, 1], D1 * c, rtol=1e-12, atol=1e-13)


def test_boundary_flux_of_constant_field_on_flat_bottom():
    """flux = (fx, fy) constant, bottom r1 = sigma, r2 in [0,5]: int f.n 4 pi r1^2 ds = -fy 4 pi sigma^2 5."""
    m = _mesh()
    markers = fem.mark_facets(m, [(lambda x: fem.near(x[..., 1], 0.1), 24), (lambda x: fem.near(x[..., 0], 5.0), 21)])
    F = np.tile(np.array([[0.3, -0.8]]), (m.num_cells, 1))
    r2, r1, tot = ofn.boundary_flux(m, F, markers, 24)
    assert abs(r2) < 1e-15 and abs(r1 - 0.8 * 4 * np.pi * 0.01 * 5.0) < 1e-13 and tot == r2 + r1
    # right boundary x = 5, r1 in [0.1,5]: n = (1,0): fx 4 pi (5^3 - 0.1^3)/3
    r2, r1, tot = ofn.boundary_flux(m, F, markers, 21)
    assert abs(r1) < 1e-15 and abs(r2 - 0.3 * 4 * np.pi * (125 - 1e-3) / 3) < 1e-11
