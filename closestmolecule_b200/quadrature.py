"""FIAT "default" quadrature tables (15-digit constants as FIAT stores them) --
the host-side copy of csrc/cm_common.cuh, used to evaluate forcing data at the kernel's points."""
import numpy as np

_A1, _B1 = 0.816847572980459, 0.091576213509771
_A2, _B2 = 0.108103018168070, 0.445948490915965
_TRI = {
    3: (np.array([[1 / 6, 1 / 6], [1 / 6, 2 / 3], [2 / 3, 1 / 6]]), np.array([1 / 6] * 3)),
    6: (np.array([[_A1, _B1], [_B1, _A1], [_B1, _B1], [_A2, _B2], [_B2, _A2], [_B2, _B2]]),
        np.array([0.109951743655322 / 2] * 3 + [0.223381589678011 / 2] * 3)),
    12: (np.array([[0.873821971016996, 0.063089014491502], [0.063089014491502, 0.873821971016996],
                   [0.063089014491502, 0.063089014491502], [0.501426509658179, 0.249286745170910],
                   [0.249286745170910, 0.501426509658179], [0.249286745170910, 0.249286745170910],
                   [0.636502499121399, 0.310352451033785], [0.636502499121399, 0.053145049844816],
                   [0.310352451033785, 0.636502499121399], [0.310352451033785, 0.053145049844816],
                   [0.053145049844816, 0.636502499121399], [0.053145049844816, 0.310352451033785]]),
         np.array([0.050844906370207 / 2] * 3 + [0.116786275726379 / 2] * 3
                  + [0.082851075618374 / 2] * 6)),
}


def nq_of_degree(degree):
    return 3 if degree <= 2 else (6 if degree <= 4 else 12)


def triangle_rule(degree):
    """(points [nq,2] in (xi,eta), weights) of the rule the kernels use for `degree`."""
    if degree > 6:
        raise NotImplementedError("kernels support quadrature degree <= 6")
    return _TRI[nq_of_degree(degree)]


def facet_rule(degree):
    """Gauss-Legendre on [0,1]: 3 points up to degree 4, else 4 (matches cm::facet_rule)."""
    m = 3 if degree <= 4 else 4
    p, w = np.polynomial.legendre.leggauss(m)
    return 0.5 * (p + 1.0), 0.5 * w


def collapsed_triangle_rule(degree):
    """Collapsed Gauss-Jacobi rule on the UFC triangle, m = (degree+2)//2 points per direction (exact for
    polynomials of that degree): what FIAT's default scheme uses above degree 6.  Points [m*m,2]
    in (xi, eta), weights summing to 1/2."""
    from scipy.special import roots_jacobi
    m = (degree + 2) // 2
    px, wx = roots_jacobi(m, 0.0, 0.0)
    py, wy = roots_jacobi(m, 1.0, 0.0)
    pts, wts = [], []
    for i in range(m):
        for j in range(m):
            eta = 0.5 * (1.0 + py[j])
            xi = 0.5 * (1.0 + px[i]) * (1.0 - eta)
            pts.append([xi, eta])
            wts.append(0.125 * wx[i] * wy[j])
    return np.array(pts), np.array(wts)
