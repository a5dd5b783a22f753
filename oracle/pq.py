"""Oracle for the P-Q primal P1 (CG1 x CG1) Newton system (test infrastructure).

Follows, term by term:
* advection_diffusion_convergence.py:133-135,196-215   weight, form F-C (+ transient term :198),
  Gstar :21-27, BC list :174-183, Newton parameters :207-215;
* advection_diffusion_mixed.py:17-23                    production Gstar (threshold 1e-3, c in the condition);
* …MixedPrimal_convergence.py:43-44,111-113,120-123    manufactured G, f2_ex/f3_ex, q-equation with reaction;
* …MixedPrimal_convergence_SUPG.py:112,118-129         SUPG test function and residual (x^2*q as written);
* …MixedPrimal_convergence_C0IP.py:141-147, test_for_paper_convergence.py:140-153
                                                         C0IP cell / exterior-facet / interior-facet q-terms.
Cell-local dof order [p@v0,v1,v2, q@v0,v1,v2]; global dof 2*vertex+component (SURVEY B.3 with
reorder_dofs_serial=False).  The arithmetic is written the way FFC generates it (a plain sum over
quadrature points of the integrand), NOT the way the CUDA kernels reorganise it.
"""
from __future__ import annotations

from dataclasses import dataclass, field, replace

import numpy as np

from . import fem
from .fem import DOLFIN_EPS

G_NONE, G_POWER, G_GSTAR = 0, 1, 2
H_CELL_DIAMETER, H_FACET_AREA = 0, 1
# exterior-facet kind bits
EF_ADV, EF_DATA, EF_PENALTY = 1, 2, 4


@dataclass
class PQParams:
    D1: float = 2.0
    D2: float = 1.5
    inv_dt: float = 0.0            # 1/dt, 0 => steady state (advection_diffusion_convergence.py:196-204)
    gkind: int = G_GSTAR
    gcoef: float = 1.0             # power: gcoef*x^2*p^2/(q+geps)
    geps: float = DOLFIN_EPS
    gs_thr: float = 1e-6           # gstar: |gs_cond_scale*p - q| > gs_thr ? (p/q)*p*x^2 : gs_lin*p*x^2
    gs_cond_scale: float = 1.0 / (4.0 * np.pi)
    gs_lin: float = 4.0 * np.pi
    q_react: float = 0.0           # + q*w
    q_adv: float = 1.0             # + dq/dx * w
    q_x2p: float = 1.0             # + x^2 p w
    q_x2q: float = 0.0             # + x^2 q w   (SUPG script quirk, :119)
    q_divr: float = 0.0            # - (2/x) q w (C0IP)
    q_ibp: float = 0.0             # - q dw/dx   (C0IP)
    supg_stab: float = 0.0         # test function w + supg_stab*hK/2*dw/dx
    degree: int = 4                # quadrature degree
    # C0IP facet terms
    c0ip_gamma: float = 0.0        # sign*stab/(k_q)^3.5*beta (0 => no dS integral, narrow pattern)
    c0ip_h: int = H_CELL_DIAMETER
    pen_stab2: float = 0.0         # stab2 of test_for_paper_convergence.py:145
    r2vec: tuple = (1.0, 0.0)


def gstar_conv(c):
    """advection_diffusion_convergence.py:21-27."""
    return dict(gkind=G_GSTAR, gs_thr=1e-6, gs_cond_scale=1.0 / (4.0 * np.pi), gs_lin=4.0 * np.pi * c)


def gstar_prod(c):
    """advection_diffusion_mixed.py:17-23."""
    return dict(gkind=G_GSTAR, gs_thr=1e-3, gs_cond_scale=1.0 / (4.0 * c * np.pi), gs_lin=4.0 * np.pi * c)


def eval_G(prm, x, p, q):
    """G, dG/dp, dG/dq at points (UFL differentiates the branches of a conditional, not the
    condition; both branches are evaluated eagerly -- SURVEY B.6)."""
    if prm.gkind == G_NONE:
        z = np.zeros_like(p)
        return z, z, z
    x2 = x * x
    if prm.gkind == G_POWER:
        qe = q + prm.geps
        G = prm.gcoef * x2 * p * p / qe
        return G, 2.0 * prm.gcoef * x2 * p / qe, -prm.gcoef * x2 * p * p / (qe * qe)
    with np.errstate(all="ignore"):
        cond = np.abs(prm.gs_cond_scale * p - q) > prm.gs_thr
        G = np.where(cond, (p / q) * p * x2, prm.gs_lin * p * x2)
        Gp = np.where(cond, 2.0 * p * x2 / q, prm.gs_lin * x2)
        Gq = np.where(cond, -(p * p) * x2 / (q * q), 0.0)
    return G, Gp, Gq


def element_cells(mesh, u, prm, p_old=None, forcing=None, want_jac=True):
    """Fe [C,6], Je [C,6,6] of the cell integrals."""
    det, grads, X = fem.cell_geometry(mesh)
    Gx, Gy = grads[:, :, 0], grads[:, :, 1]
    pts, wts = fem.triangle_rule(prm.degree)
    phi = np.stack([1.0 - pts[:, 0] - pts[:, 1], pts[:, 0], pts[:, 1]], 1)     # [nq,3]
    P = np.einsum("qi,cid->cqd", phi, X)
    x, y = P[:, :, 0], P[:, :, 1]
    W = (4.0 * np.pi * y * x) ** 2
    a = wts[None, :] * np.abs(det)[:, None] * W                                # [C,nq]
    pe = u[0::2][mesh.cells]
    qe = u[1::2][mesh.cells]
    pq_ = pe @ phi.T
    qq_ = qe @ phi.T
    px = np.einsum("ci,ci->c", pe, Gx)[:, None]
    py = np.einsum("ci,ci->c", pe, Gy)[:, None]
    qx = np.einsum("ci,ci->c", qe, Gx)[:, None]
    G, Gp, Gq = eval_G(prm, x, pq_, qq_)
    x2 = x * x
    tau = prm.supg_stab * fem.cell_diameter(mesh) / 2.0 if prm.supg_stab != 0.0 else np.zeros(mesh.num_cells)
    f2 = f3 = 0.0
    if forcing is not None:
        f2, f3 = forcing(x, y)
    dpt = 0.0
    if prm.inv_dt != 0.0:
        dpt = (pq_ - p_old[mesh.cells] @ phi.T) * prm.inv_dt
    resid = (prm.q_react * qq_ + prm.q_adv * qx + x2 * (prm.q_x2p * pq_ + prm.q_x2q * qq_)
             - prm.q_divr * (2.0 / x) * qq_ - f3)
    C = mesh.num_cells
    Fe = np.zeros((C, 6))
    for i in range(3):
        ph = phi[None, :, i]
        gxi = Gx[:, i][:, None]
        gyi = Gy[:, i][:, None]
        Fe[:, i] = (a * (dpt * ph + prm.D2 * (px + G) * gxi + prm.D1 * py * gyi - f2 * ph)).sum(1)
        wt = ph + tau[:, None] * gxi
        Fe[:, 3 + i] = (a * (resid * wt - prm.q_ibp * qq_ * gxi)).sum(1)
    if not want_jac:
        return Fe, None
    Je = np.zeros((C, 6, 6))
    for i in range(3):
        ph = phi[None, :, i]
        gxi = Gx[:, i][:, None]
        gyi = Gy[:, i][:, None]
        wt = ph + tau[:, None] * gxi
        for j in range(3):
            pj = phi[None, :, j]
            gxj = Gx[:, j][:, None]
            gyj = Gy[:, j][:, None]
            Je[:, i, j] = (a * (prm.inv_dt * pj * ph + prm.D2 * (gxj + Gp * pj) * gxi
                                + prm.D1 * gyj * gyi)).sum(1)
            Je[:, i, 3 + j] = (a * (prm.D2 * Gq * pj * gxi)).sum(1)
            Je[:, 3 + i, j] = (a * (x2 * prm.q_x2p * pj * wt)).sum(1)
            Je[:, 3 + i, 3 + j] = (a * ((prm.q_react * pj + prm.q_adv * gxj + x2 * prm.q_x2q * pj
                                         - prm.q_divr * (2.0 / x) * pj) * wt
                                        - prm.q_ibp * pj * gxi)).sum(1)
    return Fe, Je


def element_exterior_facets(mesh, u, prm, fidx, kinds, q_data=None):
    """q-row exterior-facet terms on facets fidx (global facet ids) with per-facet kind bits:
    EF_ADV     + (r2vec.n) q w W ds                     (…C0IP.py:146, test_for_paper:142-143)
    EF_DATA    + (r2vec.n) q_ex w W ds                  (test_for_paper:152, moved to the lhs of FF)
    EF_PENALTY + stab2 (hK neg(r2vec.n)^2 + 8/hK)(q - q_ex) w ds, unweighted, hK=FacetArea (:145,153)
    q_data(x,y) -> q_ex.  Returns cell ids, Fe [n,3] and Je [n,3,3] on the attached cell's q dofs."""
    fc = mesh.facets()
    L, nrm, A, B = fem.facet_geometry(mesh, fidx)
    s, w = fem.facet_rule(prm.degree)
    cell = fc.cell0[fidx]
    loc = fc.loc0[fidx]
    cv = mesh.cells[cell]                                     # [n,3]
    Xc = mesh.coords[cv]                                      # [n,3,2]
    P = A[:, None, :] + s[None, :, None] * (B - A)[:, None, :]
    # basis values of the attached cell at the facet points: barycentric via the two facet vertices
    va = fc.verts[fidx, 0][:, None] == cv                     # which local vertex is A
    vb = fc.verts[fidx, 1][:, None] == cv
    phi = va[:, None, :] * (1.0 - s)[None, :, None] + vb[:, None, :] * s[None, :, None]   # [n,nq,3]
    x, y = P[:, :, 0], P[:, :, 1]
    W = (4.0 * np.pi * y * x) ** 2
    bn = nrm[:, 0] * prm.r2vec[0] + nrm[:, 1] * prm.r2vec[1]
    qe = u[1::2][cv]
    qq_ = np.einsum("nqi,ni->nq", phi, qe)
    qx_ = q_data(x, y) if q_data is not None else np.zeros_like(x)
    ds = w[None, :] * L[:, None]
    neg = 0.5 * (np.abs(bn) - bn)
    pen = prm.pen_stab2 * (L * neg ** 2 + 8.0 / L)
    k_adv = ((kinds & EF_ADV) != 0).astype(float)
    k_dat = ((kinds & EF_DATA) != 0).astype(float)
    k_pen = ((kinds & EF_PENALTY) != 0).astype(float)
    cq = (k_adv * bn)[:, None] * W + (k_pen * pen)[:, None]             # multiplies q w
    cd = (k_dat * bn)[:, None] * W - (k_pen * pen)[:, None]             # multiplies q_ex w
    Fe = np.einsum("nq,nqi->ni", ds * (cq * qq_ + cd * qx_), phi)
    Je = np.einsum("nq,nqi,nqj->nij", ds * cq, phi, phi)
    return cell, Fe, Je


def element_interior_facets(mesh, u, prm):
    """C0IP jump penalty gamma*avg(hK)^2*(jump(grad q).n+)(jump(grad w).n+)*W dS over all interior
    facets (…C0IP.py:147; test_for_paper_convergence.py:144).  Returns macro vertex ids [Fi,6]
    (3 of cell0 then 3 of cell1), Fe [Fi,6], Je [Fi,6,6] on the q dofs."""
    fc = mesh.facets()
    fi = fc.interior
    L, nrm, A, B = fem.facet_geometry(mesh, fi)
    det, grads, X = fem.cell_geometry(mesh)
    c0, c1 = fc.cell0[fi], fc.cell1[fi]
    d = np.concatenate([np.einsum("nid,nd->ni", grads[c0], nrm),
                        -np.einsum("nid,nd->ni", grads[c1], nrm)], 1)   # [Fi,6]
    if prm.c0ip_h == H_FACET_AREA:
        hbar = L
    else:
        hK = fem.cell_diameter(mesh)
        hbar = 0.5 * (hK[c0] + hK[c1])
    s, w = fem.facet_rule(prm.degree)
    P = A[:, None, :] + s[None, :, None] * (B - A)[:, None, :]
    W = (4.0 * np.pi * P[:, :, 1] * P[:, :, 0]) ** 2
    intW = (w[None, :] * W).sum(1) * L
    coef = prm.c0ip_gamma * hbar ** 2 * intW
    Je = coef[:, None, None] * d[:, :, None] * d[:, None, :]
    mv = np.concatenate([mesh.cells[c0], mesh.cells[c1]], 1)
    qm = u[1::2][mv]
    Fe = np.einsum("nij,nj->ni", Je, qm)
    return mv, Fe, Je


class PQSystem:
    """DOLFIN-pattern CSR of the CG1xCG1 space and serial-order assembly of cells, then exterior
    facets, then interior facets."""

    def __init__(self, mesh, prm, ext_facets=None, ext_kinds=None):
        self.mesh, self.prm = mesh, prm
        self.with_dS = prm.c0ip_gamma != 0.0
        self.nrowptr, self.ncol = fem.p1_node_pattern(mesh, self.with_dS)
        self.rowptr, self.col = fem.blocked_pattern(self.nrowptr, self.ncol, 2)
        c = mesh.cells.astype(np.int64)
        ld = np.concatenate([2 * c, 2 * c + 1], 1)                        # [C,6]
        self.cell_dofs = ld
        self.cell_slots = fem.csr_slots(self.rowptr, self.col, np.repeat(ld, 6, axis=1),
                                        np.tile(ld, (1, 6))).reshape(-1, 6, 6)
        self.ext_facets = np.zeros(0, dtype=np.int64) if ext_facets is None else np.asarray(ext_facets)
        self.ext_kinds = np.zeros(0, dtype=np.int64) if ext_kinds is None else np.asarray(ext_kinds)
        self.N = 2 * mesh.num_vertices

    def assemble(self, u, want_jac=True, p_old=None, forcing=None, q_data=None):
        prm, mesh = self.prm, self.mesh
        Fe, Je = element_cells(mesh, u, prm, p_old, forcing, want_jac)
        b = np.zeros(self.N)
        fem.add_serial(b, self.cell_dofs, Fe)
        vals = None
        if want_jac:
            vals = np.zeros(self.col.size)
            fem.add_serial(vals, self.cell_slots, Je)
        if self.ext_facets.size:
            cell, Fx, Jx = element_exterior_facets(mesh, u, prm, self.ext_facets, self.ext_kinds, q_data)
            qd = 2 * mesh.cells[cell].astype(np.int64) + 1
            fem.add_serial(b, qd, Fx)
            if want_jac:
                sl = fem.csr_slots(self.rowptr, self.col, np.repeat(qd, 3, axis=1), np.tile(qd, (1, 3)))
                fem.add_serial(vals, sl, Jx)
        if self.with_dS:
            mv, Fi, Ji = element_interior_facets(mesh, u, prm)
            qd = 2 * mv.astype(np.int64) + 1
            fem.add_serial(b, qd, Fi)
            if want_jac:
                sl = fem.csr_slots(self.rowptr, self.col, np.repeat(qd, 6, axis=1), np.tile(qd, (1, 6)))
                fem.add_serial(vals, sl, Ji)
        return vals, b


# ----------------------------------------------------------------------------------------------
# Dirichlet data


def facet_vertices_with_marker(mesh, markers, mid):
    fc = mesh.facets()
    return np.unique(fc.verts[markers == mid])


def p_initial(x, y, c, V1, sigma_b, right=False):
    """PInitial.eval (advection_diffusion_convergence.py:40-44; advection_diffusion_mixed.py:41-45);
    sigma_b = flat_reaction_boundary() or the curved boundary value at x."""
    base = c / V1 * np.exp(-4.0 / 3.0 * np.pi * c * np.power(x, 3))
    with np.errstate(divide="ignore", invalid="ignore"):
        full = base * (1.0 - sigma_b / y)
    if right:
        return np.where(y > 0, full, base)
    return full


def q_initial(x, c, V1):
    """QInitial.eval (advection_diffusion_convergence.py:79-80)."""
    return 1.0 / (4.0 * np.pi * V1) * np.exp(-4.0 / 3.0 * np.pi * c * np.power(x, 3))


# ----------------------------------------------------------------------------------------------
# manufactured data of the MixedPrimal scripts (p_ex = sin(pi x) sin(pi y), q_ex = 1.5 + cos(pi x y))


def p_exact(x, y):
    return np.sin(np.pi * x) * np.sin(np.pi * y)


def q_exact(x, y):
    return 1.5 + np.cos(np.pi * x * y)


def manufactured_forcing(prm, x2_on_q=False, react=None):
    """(f2_ex, f3_ex) of …MixedPrimal_convergence.py:111-113 for the G selected in prm (power kind):
    f2 = div_rad(D*sig_ex), sig_ex = -grad(p_ex) - G(p_ex,q_ex) r2vec;
    f3 = react*q_ex + dq_ex/dx + x^2*(p_ex or q_ex)   (q_ex in the SUPG script :112)."""
    import sympy as sm
    xs, ys = sm.symbols("x y", positive=True)
    pe = sm.sin(sm.pi * xs) * sm.sin(sm.pi * ys)
    qe = sm.Rational(3, 2) + sm.cos(sm.pi * xs * ys)
    if prm.gkind == G_POWER:
        G = prm.gcoef * xs ** 2 * pe ** 2 / (qe + prm.geps)
    elif prm.gkind == G_NONE:
        G = 0
    else:
        raise ValueError("manufactured forcing is defined for the power-law G only")
    sx = -sm.diff(pe, xs) - G
    sy = -sm.diff(pe, ys)
    f2 = (sm.diff(xs ** 2 * prm.D2 * sx, xs) / xs ** 2 + sm.diff(ys ** 2 * prm.D1 * sy, ys) / ys ** 2)
    r = prm.q_react if react is None else react
    f3 = r * qe + sm.diff(qe, xs) + xs ** 2 * (qe if x2_on_q else pe)
    f2n = sm.lambdify((xs, ys), f2, "numpy")
    f3n = sm.lambdify((xs, ys), f3, "numpy")
    return lambda x, y: (f2n(x, y), f3n(x, y))


def errors_pq(mesh, u, degree=4):
    """Weighted error norms in the style of …MixedPrimal_convergence.py:148-150 for the P1 primal
    restatement: e0(p) = ||p_ex-p_h||_W, e1(p) = |p_ex-p_h|_{1,W}, e1(q) = ||q-q_h||_W + |.|_{1,W}."""
    det, grads, X = fem.cell_geometry(mesh)
    pts, wts = fem.triangle_rule(degree)
    phi = np.stack([1.0 - pts[:, 0] - pts[:, 1], pts[:, 0], pts[:, 1]], 1)
    P = np.einsum("qi,cid->cqd", phi, X)
    x, y = P[:, :, 0], P[:, :, 1]
    a = wts[None, :] * np.abs(det)[:, None] * (4.0 * np.pi * x * y) ** 2
    pe = u[0::2][mesh.cells]
    qe = u[1::2][mesh.cells]
    ph, qh = pe @ phi.T, qe @ phi.T
    gx = lambda f: np.einsum("ci,ci->c", f, grads[:, :, 0])[:, None]
    gy = lambda f: np.einsum("ci,ci->c", f, grads[:, :, 1])[:, None]
    pex, qex = p_exact(x, y), q_exact(x, y)
    pex_x = np.pi * np.cos(np.pi * x) * np.sin(np.pi * y)
    pex_y = np.pi * np.sin(np.pi * x) * np.cos(np.pi * y)
    qex_x = -np.pi * y * np.sin(np.pi * x * y)
    qex_y = -np.pi * x * np.sin(np.pi * x * y)
    e0p = np.sqrt((a * (pex - ph) ** 2).sum())
    e1p = np.sqrt((a * ((pex_x - gx(pe)) ** 2 + (pex_y - gy(pe)) ** 2)).sum())
    e1q = np.sqrt((a * ((qex - qh) ** 2 + (qex_x - gx(qe)) ** 2 + (qex_y - gy(qe)) ** 2)).sum())
    return float(e0p), float(e1p), float(e1q)
