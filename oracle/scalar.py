"""Oracle for the scalar P1 forms (test infrastructure).

F-B: SphericalAdvectionDiffusionReaction_convergence.py:92-110 (nonlinear, Newton, degree-4 rule) --
     the script whose output is recorded in README.md:109-115 (golden vector).
F-A: SphericalLaplace_convergence.py:128-135 (linear; UFL-estimated degrees: bilinear 2, rhs 8).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import fem
from .newton import newton_solve, apply_bc_matrix_vec


def _points(X, pts):
    """physical points [C,nq,2] and basis values [nq,3]."""
    phi = np.stack([1.0 - pts[:, 0] - pts[:, 1], pts[:, 0], pts[:, 1]], 1)
    return np.einsum("qi,cid->cqd", phi, X), phi


# manufactured data u_ex = sin(pi x) sin(pi y) (both scripts, u_str)
def u_ex(x, y):
    return np.sin(np.pi * x) * np.sin(np.pi * y)


def grad_u_ex(x, y):
    return (np.pi * np.cos(np.pi * x) * np.sin(np.pi * y), np.pi * np.sin(np.pi * x) * np.cos(np.pi * y))


def f_fb(x, y, D1, D2):
    """f_ex = u_ex - Lap_r2r1(u_ex) - div_r2(4*pi*r2**2*u_ex**2*r2vec) with the D-weighted operators
    of SphericalAdvectionDiffusionReaction_convergence.py:28-35,85 (density = 1)."""
    u = u_ex(x, y)
    ux, uy = grad_u_ex(x, y)
    uxx = -np.pi ** 2 * u
    uyy = -np.pi ** 2 * u
    lap = D2 * uxx + D2 * 2.0 / x * ux + D1 * uyy + D1 * 2.0 / y * uy
    g = 4.0 * np.pi * x ** 2 * u ** 2
    gx = 4.0 * np.pi * (2.0 * x * u ** 2 + x ** 2 * 2.0 * u * ux)
    divg = D2 * gx + D2 * 2.0 / x * g
    return u - lap - divg


def element_fb(mesh, u, D1=1e-3, D2=1e-4, degree=4, want_jac=True):
    """Per-cell F [C,3] and J [C,3,3] of form F-B."""
    det, grads, X = fem.cell_geometry(mesh)
    pts, wts = fem.triangle_rule(degree)
    P, phi = _points(X, pts)
    ue = u[mesh.cells]                                    # [C,3]
    uq = ue @ phi.T                                       # [C,nq]
    ux = np.einsum("ci,ci->c", ue, grads[:, :, 0])
    uy = np.einsum("ci,ci->c", ue, grads[:, :, 1])
    x, y = P[:, :, 0], P[:, :, 1]
    w = wts[None, :] * np.abs(det)[:, None]
    g = D2 * 4.0 * np.pi * x ** 2 * uq ** 2
    gp = D2 * 8.0 * np.pi * x ** 2 * uq
    f = f_fb(x, y, D1, D2)
    Gx, Gy = grads[:, :, 0], grads[:, :, 1]               # [C,3]
    F = np.zeros((mesh.num_cells, 3))
    for i in range(3):
        F[:, i] = (w * (uq * phi[None, :, i] + (D2 * ux * Gx[:, i] + D1 * uy * Gy[:, i])[:, None]
                        - (2.0 * D2 / x) * ux[:, None] * phi[None, :, i]
                        - (2.0 * D1 / y) * uy[:, None] * phi[None, :, i]
                        + g * Gx[:, i][:, None] - (2.0 / x) * g * phi[None, :, i]
                        - f * phi[None, :, i])).sum(1)
    if not want_jac:
        return F, None
    J = np.zeros((mesh.num_cells, 3, 3))
    for i in range(3):
        for j in range(3):
            J[:, i, j] = (w * (phi[None, :, j] * phi[None, :, i]
                               + (D2 * Gx[:, j] * Gx[:, i] + D1 * Gy[:, j] * Gy[:, i])[:, None]
                               - (2.0 * D2 / x) * Gx[:, j][:, None] * phi[None, :, i]
                               - (2.0 * D1 / y) * Gy[:, j][:, None] * phi[None, :, i]
                               + gp * phi[None, :, j] * Gx[:, i][:, None]
                               - (2.0 / x) * gp * phi[None, :, j] * phi[None, :, i])).sum(1)
    return F, J


def f_fa(x, y, D):
    """f_ex = u_ex - D*Lap_r2r1(u_ex) (SphericalLaplace_convergence.py:76-83,121)."""
    u = u_ex(x, y)
    ux, uy = grad_u_ex(x, y)
    return u - D * (-2.0 * np.pi ** 2 * u + 2.0 / x * ux + 2.0 / y * uy)


def element_fa(mesh, D=1e-4, deg_a=2, deg_rhs=8):
    """Per-cell A [C,3,3] (degree-2 rule) and b [C,3] (degree-8 rule) of form F-A."""
    det, grads, X = fem.cell_geometry(mesh)
    Gx, Gy = grads[:, :, 0], grads[:, :, 1]
    pts, wts = fem.triangle_rule(deg_a)
    P, phi = _points(X, pts)
    x, y = P[:, :, 0], P[:, :, 1]
    w = wts[None, :] * np.abs(det)[:, None]
    A = np.zeros((mesh.num_cells, 3, 3))
    for i in range(3):
        for j in range(3):
            A[:, i, j] = (w * (phi[None, :, j] * phi[None, :, i]
                               + D * (Gx[:, j] * Gx[:, i] + Gy[:, j] * Gy[:, i])[:, None]
                               - (2.0 * D / x) * Gx[:, j][:, None] * phi[None, :, i]
                               - (2.0 * D / y) * Gy[:, j][:, None] * phi[None, :, i])).sum(1)
    pts, wts = fem.triangle_rule(deg_rhs)
    P, phi = _points(X, pts)
    w = wts[None, :] * np.abs(det)[:, None]
    f = f_fa(P[:, :, 0], P[:, :, 1], D)
    b = np.einsum("cq,qi->ci", w * f, phi)
    return A, b


class ScalarP1System:
    """CSR pattern + slots for a scalar P1 space on a mesh (dof = vertex, B.3)."""

    def __init__(self, mesh):
        self.mesh = mesh
        self.rowptr, self.col = fem.p1_node_pattern(mesh)
        c = mesh.cells.astype(np.int64)
        self.slots = fem.csr_slots(self.rowptr, self.col, np.repeat(c, 3, axis=1), np.tile(c, (1, 3)))
        self.slots = self.slots.reshape(-1, 3, 3)

    def assemble(self, Fe, Je):
        V = self.mesh.num_vertices
        b = np.zeros(V)
        fem.add_serial(b, self.mesh.cells.astype(np.int64), Fe)
        vals = None
        if Je is not None:
            vals = np.zeros(self.col.size)
            fem.add_serial(vals, self.slots, Je)
        return vals, b

    def boundary_dofs(self):
        fc = self.mesh.facets()
        return np.unique(fc.verts[fc.exterior])


def errors(mesh, u, degree, exact=u_ex, grad_exact=grad_u_ex):
    """sqrt(assemble((grad(u_ex)-grad(u_h))**2*dx)), sqrt(assemble((u_ex-u_h)**2*dx))
    (SphericalAdvectionDiffusionReaction_convergence.py:117-118)."""
    det, grads, X = fem.cell_geometry(mesh)
    pts, wts = fem.triangle_rule(degree)
    P, phi = _points(X, pts)
    ue = u[mesh.cells]
    uq = ue @ phi.T
    ux = np.einsum("ci,ci->c", ue, grads[:, :, 0])[:, None]
    uy = np.einsum("ci,ci->c", ue, grads[:, :, 1])[:, None]
    w = wts[None, :] * np.abs(det)[:, None]
    ex = exact(P[:, :, 0], P[:, :, 1])
    gx, gy = grad_exact(P[:, :, 0], P[:, :, 1])
    return (float(np.sqrt((w * ((gx - ux) ** 2 + (gy - uy) ** 2)).sum())),
            float(np.sqrt((w * (ex - uq) ** 2).sum())))


def solve_fb(n, D1=1e-3, D2=1e-4, atol=1e-7, rtol=1e-7):
    """One level of SphericalAdvectionDiffusionReaction_convergence.py:62-118."""
    mesh = fem.rectangle_mesh(0.0, 0.0, 1.0, 1.0, n, n)
    S = ScalarP1System(mesh)
    bd = S.boundary_dofs()
    g = u_ex(mesh.coords[bd, 0], mesh.coords[bd, 1])

    def asm(x, want_jac):
        Fe, Je = element_fb(mesh, x, D1, D2, 4, want_jac)
        return S.assemble(Fe, Je)

    x, it, conv = newton_solve(asm, np.zeros(mesh.num_vertices), S.rowptr, S.col, bd, g,
                               atol=atol, rtol=rtol)
    e1, e0 = errors(mesh, x, 4)
    return dict(dof=mesh.num_vertices, h=mesh.hmax(), e1=e1, e0=e0, newton_its=it, converged=conv,
                u=x, mesh=mesh)


def solve_fa(n, D=1e-4):
    """One level of SphericalLaplace_convergence.py:100-143 (linear solve; BC rows replaced --
    DOLFIN's symmetric assemble_system gives the same solution)."""
    mesh = fem.rectangle_mesh(0.0, 0.0, 1.0, 1.0, n, n)
    S = ScalarP1System(mesh)
    Ae, be = element_fa(mesh, D)
    vals, b = S.assemble(be, Ae)
    bd = S.boundary_dofs()
    apply_bc_matrix_vec(S.rowptr, S.col, vals, bd)
    b[bd] = u_ex(mesh.coords[bd, 0], mesh.coords[bd, 1])
    A = sp.csr_matrix((vals, S.col, S.rowptr), shape=(b.size, b.size))
    u = spla.splu(A.tocsc()).solve(b)
    e1, _ = errors(mesh, u, 12)
    _, e0 = errors(mesh, u, 14)
    return dict(dof=mesh.num_vertices, h=mesh.hmax(), e1=e1, e0=e0, u=u, mesh=mesh)


def convergence_table(solve, levels):
    """Rows (DoF, h, e1, r1, e0, r0) formatted like the scripts' printout."""
    rows = []
    prev = None
    for nk in range(levels):
        r = solve(2 ** (nk + 1))
        r1 = r0 = 0.0
        if prev is not None:
            r1 = np.log(r["e1"] / prev["e1"]) / np.log(r["h"] / prev["h"])
            r0 = np.log(r["e0"] / prev["e0"]) / np.log(r["h"] / prev["h"])
        rows.append('{:6d}  {:.4f} {:6.2e}  {:.3f}  {:6.2e}  {:.3f} '.format(
            r["dof"], r["h"], r["e1"], r1, r["e0"], r0))
        prev = r
    return rows
