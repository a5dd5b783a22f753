"""CPU restatement (test infrastructure) of the scalar functionals and flux post-processing the
reference drivers compute after the Newton solve:

* generic_error_norms: assemble((u_ex-u_h)**2*[W]*dx), assemble((grad u_ex - grad u_h)**2*[W]*dx)
  (SphericalAdvectionDiffusionReaction_convergence.py:117-118, …MixedPrimal_convergence.py:148-150);
* project_flux_dg0: project(dMatrix*grad(p_h) + D2*r2*r2*p_h/q_h*p_h*r2_vec, VectorFunctionSpace(mesh,"DG",0))
  (convergence/advection_diffusion_convergence.py:222-223): for DG0 the L2 projection is the cell mean,
  evaluated with the degree-5 rule UFL estimates for the quotient term [ext];
* boundary_flux: assemble(dot(flux,n)*4*pi*r1**2*ds(marker)) and its r1 / r2 parts
  (advection_diffusion_mixed.py:261-263); degree-2 facet rule (2 Gauss points), exact here.

Parity unpinned by reference outputs (no recorded flux values ship); pinned by closed forms in
tests/test_oracle_functionals.py."""
import numpy as np

from . import fem


def generic_error_norms(mesh, uh, exact, grad_exact, degree=4, weighted=False, x_derivative_only=False):
    det, grads, X = fem.cell_geometry(mesh)
    pts, wts = fem.triangle_rule(degree)
    phi = np.stack([1.0 - pts[:, 0] - pts[:, 1], pts[:, 0], pts[:, 1]], 1)
    P = np.einsum("qi,cid->cqd", phi, X)
    x, y = P[:, :, 0], P[:, :, 1]
    a = wts[None, :] * np.abs(det)[:, None]
    if weighted:
        a = a * (4.0 * np.pi * x * y) ** 2
    ue = uh[mesh.cells]
    uq = ue @ phi.T
    ux = np.einsum("ci,ci->c", ue, grads[:, :, 0])[:, None]
    uy = np.einsum("ci,ci->c", ue, grads[:, :, 1])[:, None]
    gx, gy = grad_exact(x, y)
    e0 = np.sqrt((a * (exact(x, y) - uq) ** 2).sum())
    d = (gx - ux) ** 2 if x_derivative_only else (gx - ux) ** 2 + (gy - uy) ** 2
    return float(e0), float(np.sqrt((a * d).sum()))


def project_flux_dg0(mesh, u, D1, D2, degree=5):
    det, grads, X = fem.cell_geometry(mesh)
    pts, wts = fem.triangle_rule(degree)
    phi = np.stack([1.0 - pts[:, 0] - pts[:, 1], pts[:, 0], pts[:, 1]], 1)
    x = np.einsum("qi,ci->cq", phi, X[:, :, 0])
    pe, qe = u[0::2][mesh.cells], u[1::2][mesh.cells]
    p, q = pe @ phi.T, qe @ phi.T
    px = np.einsum("ci,ci->c", pe, grads[:, :, 0])
    py = np.einsum("ci,ci->c", pe, grads[:, :, 1])
    mean = ((x * x * p / q * p) * wts[None, :]).sum(1) / wts.sum()
    return np.stack([D2 * (px + mean), D1 * py], 1)


def boundary_flux(mesh, flux, markers, marker):
    fc = mesh.facets()
    fidx = np.nonzero((fc.cell1 < 0) & (markers == marker))[0]
    L, nrm, A, B = fem.facet_geometry(mesh, fidx)
    s, w = fem.facet_rule(2)
    y = A[:, 1][:, None] + s[None, :] * (B[:, 1] - A[:, 1])[:, None]
    wgt = 4.0 * np.pi * L * (w[None, :] * y * y).sum(1)
    F = flux[fc.cell0[fidx]]
    r2 = float((F[:, 0] * nrm[:, 0] * wgt).sum())
    r1 = float((F[:, 1] * nrm[:, 1] * wgt).sum())
    return r2, r1, r2 + r1
