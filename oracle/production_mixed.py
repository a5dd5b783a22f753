"""Oracle for the production mixed formulation DG2 x RT3 x CG3 of advection_diffusion_mixed.py:94-269
(solve_problem_instance; deg = 2 at :135,147-150) -- test infrastructure.  PARITY UNPINNED: the reference
ships no output of this script (.gitignore:5-10); the only check against the reference is the analytic
reaction rate 4 pi D1 sigma c/(gamma + c) it prints next to its own total flux (:264-267).

Weak form (:196-216), unknowns p in DG2, s in RT3, q in CG3, W = (4 pi r1 r2)^2, x = r2, y = r1:
  [(p - p_old)/dt] v W + div_rad(D s) v W + (s + Gstar(p,q) e_x).tau W - p div_rad(tau) W
  + p_right (tau.n) W ds(right) + p_top (tau.n) W ds(top) + (dq/dx + x^2 p) w W = 0,
essential: s = 0 on the left facets (:181-182), q = q_initial on the right facets (:187).
Gstar is the production conditional (:17-23).  Post-processing (:238-263): flux = project(D 4 pi r2^2 s_h, RT3),
total_flux = V1 * assemble(dot(flux, n) 4 pi r1^2 ds(bottom)) and its r1 / r2 parts.

Bases (any basis of the same spaces gives the same discrete solution):
  DG2: lambda_i (2 lambda_i - 1), 4 lambda_j lambda_k.
  RT3 (= P2^2 + x P2~, 15 per cell): per facet i (end points j < k) lambda_j^2 N_i, lambda_j lambda_k N_i,
       lambda_k^2 N_i with N_i = sigma_i (x - X_i)/(2|K|); interior lambda_i lambda_m N_i (unsigned), i = 1,2, m = 0,1,2.
  CG3: Lagrange P3 (vertices; two nodes per facet at 1/3, 2/3 from the lower vertex; the cell bubble).
Global dofs [6 per cell | 3 per edge, 6 per cell | vertices, 2 per edge, 1 per cell].
UFL would estimate its own (high) quadrature degrees for these integrals [ext]; here degree 10 on cells and 6 Gauss
points on facets.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import fem

OTHER = ((1, 2), (0, 2), (0, 1))
CELL_DEGREE = 10
FACET_POINTS = 6


def p2_values(lam):
    return np.concatenate([lam * (2 * lam - 1), np.stack([4 * lam[..., j] * lam[..., k] for j, k in OTHER], -1)], -1)


def p3_basis(lam, g):
    """Lagrange P3: values [..., 10], x-derivatives [..., 10]; order: 3 vertices, per facet i (end points j < k) the
    node nearer j then the node nearer k, then the bubble.  lam [C, nq, 3], g [C, 3, 2]."""
    gx = g[:, None, :, 0]
    val, dx = [], []
    for i in range(3):
        l = lam[..., i]
        val.append(0.5 * l * (3 * l - 1) * (3 * l - 2))
        dx.append(0.5 * (27 * l * l - 18 * l + 2) * gx[..., i])
    for j, k in OTHER:
        lj, lk = lam[..., j], lam[..., k]
        for a, b, ga, gb in ((lj, lk, gx[..., j], gx[..., k]), (lk, lj, gx[..., k], gx[..., j])):
            val.append(4.5 * a * b * (3 * a - 1))
            dx.append(4.5 * ((ga * b + a * gb) * (3 * a - 1) + a * b * 3 * ga))
    l0, l1, l2 = lam[..., 0], lam[..., 1], lam[..., 2]
    val.append(27 * l0 * l1 * l2)
    dx.append(27 * (gx[..., 0] * l1 * l2 + l0 * gx[..., 1] * l2 + l0 * l1 * gx[..., 2]))
    return np.stack(val, -1), np.stack(dx, -1)


def gstar(p, q, x2, c):
    """(G, dG/dp, dG/dq) of the production conditional (:17-23); both branches guarded like a C ternary."""
    cond = np.abs(p / (4 * c * np.pi) - q) > 1e-3
    qs = np.where(cond, q, 1.0)
    G = np.where(cond, p / qs * p * x2, 4 * np.pi * c * p * x2)
    Gp = np.where(cond, 2 * p / qs * x2, 4 * np.pi * c * x2)
    Gq = np.where(cond, -p * p / (qs * qs) * x2, 0.0)
    return G, Gp, Gq


class ProductionMixed:
    RIGHT, TOP, LEFT, BOTTOM = 21, 22, 23, 24

    def __init__(self, mesh, markers, concentration, sigma, gamma, V1, D1, D2, dt=None):
        self.mesh, self.markers = mesh, markers
        self.c, self.sigma, self.gamma, self.V1, self.D1, self.D2 = float(concentration), sigma, gamma, V1, D1, D2
        self.inv_dt = 0.0 if dt is None else 1.0 / dt
        self.fc = fc = mesh.facets()
        self.E, self.C, self.V = fc.verts.shape[0], mesh.num_cells, mesh.num_vertices
        E, C, V = self.E, self.C, self.V
        self.np_, self.ns = 6 * C, 3 * E + 6 * C
        self.nq_ = V + 2 * E + C
        self.N = self.np_ + self.ns + self.nq_
        X = mesh.coords
        cells = mesh.cells.astype(np.int64)
        fa = np.stack([cells[:, 1], cells[:, 0], cells[:, 0]], 1)
        fb = np.stack([cells[:, 2], cells[:, 2], cells[:, 1]], 1)
        key = np.minimum(fa, fb) * V + np.maximum(fa, fb)
        allkey = fc.verts[:, 0].astype(np.int64) * V + fc.verts[:, 1]
        self.cell_edges = ce = np.searchsorted(allkey, key)
        A, B = X[fc.verts[:, 0]], X[fc.verts[:, 1]]
        t = B - A
        self.elen = np.sqrt((t ** 2).sum(1))
        self.enorm = np.stack([t[:, 1], -t[:, 0]], 1) / self.elen[:, None]
        det, grads, Xc = fem.cell_geometry(mesh)
        self.area, self.Xc, self.grads = 0.5 * np.abs(det), Xc, grads
        self.sign = np.sign(((0.5 * (A + B)[ce] - Xc) * self.enorm[ce]).sum(-1))
        self.pts, self.wts = fem.triangle_rule(CELL_DEGREE)
        self.lam = np.stack([1 - self.pts[:, 0] - self.pts[:, 1], self.pts[:, 0], self.pts[:, 1]], 1)
        self.P = np.einsum("qi,cid->cqd", self.lam, Xc)
        cid = np.arange(C)[:, None]
        self.pd = 6 * cid + np.arange(6)[None, :]
        s0 = self.np_
        self.sd = np.concatenate([s0 + 3 * ce[:, i, None] + np.arange(3)[None, :] for i in range(3)]
                                 + [s0 + 3 * E + 6 * cid + np.arange(6)[None, :]], 1)          # [C,15]
        q0 = self.q0 = self.np_ + self.ns
        self.qd = np.concatenate([q0 + cells]
                                 + [q0 + V + 2 * ce[:, i, None] + np.arange(2)[None, :] for i in range(3)]
                                 + [q0 + V + 2 * E + cid], 1)                                  # [C,10]
        self.cell_dofs = np.concatenate([self.pd, self.sd, self.qd], 1)                      # [C,31]

    # -- data ---------------------------------------------------------------------------------------------
    def p_initial(self, x, y):
        sb = self.sigma * np.exp(-4 / 3 * np.pi * self.gamma * x ** 3)
        base = self.c / self.V1 * np.exp(-4 / 3 * np.pi * self.c * x ** 3)
        return np.where(y > 0, base * (1 - sb / np.where(y > 0, y, 1.0)), base)

    def q_initial(self, x):
        return 1 / (4 * np.pi * self.V1) * np.exp(-4 / 3 * np.pi * self.c * x ** 3)

    # -- bases ----------------------------------------------------------------------------------------------
    def rt_basis(self, lam, P, cells=None):
        """RT3 values [n,nq,15,2], d/dx of the x component and d/dy of the y component [n,nq,15]."""
        idx = slice(None) if cells is None else cells
        Xc, g, area, sign = self.Xc[idx], self.grads[idx], self.area[idx], self.sign[idx]
        if lam.ndim == 2:
            lam = np.broadcast_to(lam[None], (Xc.shape[0],) + lam.shape)
        funcs = []
        for i in range(3):
            j, k = OTHER[i]
            funcs += [(i, j, j, sign[:, i]), (i, j, k, sign[:, i]), (i, k, k, sign[:, i])]
        for i in (1, 2):
            funcs += [(i, i, m, None) for m in range(3)]
        R, dxx, dyy = [], [], []
        for i, a, b, sg in funcs:
            c = (1.0 if sg is None else sg) / (2 * area)
            d = P - Xc[:, None, i, :]
            qv = lam[..., a] * lam[..., b]
            qg = lam[..., a, None] * g[:, None, b, :] + lam[..., b, None] * g[:, None, a, :]
            R.append(qv[..., None] * d * c[:, None, None])
            dxx.append(c[:, None] * (qg[..., 0] * d[..., 0] + qv))
            dyy.append(c[:, None] * (qg[..., 1] * d[..., 1] + qv))
        return np.stack(R, 2), np.stack(dxx, 2), np.stack(dyy, 2)

    def assemble(self, u, want_jac=True, p_old=None):
        C, D1, D2 = self.C, self.D1, self.D2
        lam, P = self.lam, self.P
        x, y = P[:, :, 0], P[:, :, 1]
        a = self.wts[None, :] * (2 * self.area)[:, None] * (4 * np.pi * x * y) ** 2
        R, dxx, dyy = self.rt_basis(lam, P)
        Rx, Ry = R[..., 0], R[..., 1]
        dr = dxx + dyy + 2 * Rx / x[..., None] + 2 * Ry / y[..., None]
        drD = D2 * (dxx + 2 * Rx / x[..., None]) + D1 * (dyy + 2 * Ry / y[..., None])
        lamc = np.broadcast_to(lam[None], (C,) + lam.shape)
        chi = p2_values(lamc)
        phi, phix = p3_basis(lamc, self.grads)
        pe, se, qe = u[self.pd], u[self.sd], u[self.qd]
        p = np.einsum("cqb,cb->cq", chi, pe)
        s = np.einsum("cqad,ca->cqd", R, se)
        q = np.einsum("cqk,ck->cq", phi, qe)
        qx = np.einsum("cqk,ck->cq", phix, qe)
        x2 = x * x
        G, Gp, Gq = gstar(p, q, x2, self.c)
        Fe = np.zeros((C, 31))
        rp = np.einsum("cqa,ca->cq", drD, se)
        if self.inv_dt:
            rp = rp + self.inv_dt * (p - np.einsum("cqb,cb->cq", chi, p_old[self.pd]))
        Fe[:, :6] = np.einsum("cq,cqb->cb", a * rp, chi)
        Fe[:, 6:21] = np.einsum("cq,cqa->ca", a, np.einsum("cqd,cqad->cqa", s, R) + G[..., None] * Rx - p[..., None] * dr)
        Fe[:, 21:] = np.einsum("cq,cqk->ck", a * (qx + x2 * p), phi)
        b = np.zeros(self.N)
        np.add.at(b, self.cell_dofs.ravel(), Fe.ravel())
        b += self._natural()
        J = None
        if want_jac:
            Je = np.zeros((C, 31, 31))
            if self.inv_dt:
                Je[:, :6, :6] = self.inv_dt * np.einsum("cq,cqb,cqd->cbd", a, chi, chi)
            Je[:, :6, 6:21] = np.einsum("cq,cqb,cqa->cba", a, chi, drD)
            Je[:, 6:21, :6] = np.einsum("cq,cqa,cqb->cab", a, Gp[..., None] * Rx - dr, chi)
            Je[:, 6:21, 6:21] = np.einsum("cq,cqad,cqbd->cab", a, R, R)
            Je[:, 6:21, 21:] = np.einsum("cq,cqa,cqk->cak", a * Gq, Rx, phi)
            Je[:, 21:, :6] = np.einsum("cq,cqk,cqb->ckb", a * x2, phi, chi)
            Je[:, 21:, 21:] = np.einsum("cq,cqk,cql->ckl", a, phi, phix)
            J = sp.coo_matrix((Je.ravel(), (np.repeat(self.cell_dofs, 31, axis=1).ravel(),
                                            np.tile(self.cell_dofs, (1, 31)).ravel())), shape=(self.N, self.N)).tocsr()
        return J, b

    def _edge_trace(self, s):
        """normal-trace shape functions of the three dofs of a facet at the points s in [0,1] (from the lower vertex)."""
        return np.stack([(1 - s) ** 2, (1 - s) * s, s ** 2], 0)

    def _natural(self):
        """+ p_right (tau.n) W ds(right) + p_top (tau.n) W ds(top): only the facet's own three functions have a
        normal trace there, q_m(s) sign_out / |F|."""
        if hasattr(self, "_nat"):
            return self._nat
        mesh, fc = self.mesh, self.fc
        sel = np.nonzero((fc.cell1 < 0) & ((self.markers == self.RIGHT) | (self.markers == self.TOP)))[0]
        L, nrm, A, B = fem.facet_geometry(mesh, sel)
        gp, gw = np.polynomial.legendre.leggauss(FACET_POINTS)
        s, w = 0.5 * (gp + 1), 0.5 * gw
        Pq = A[:, None, :] + s[None, :, None] * (B - A)[:, None, :]
        x, y = Pq[..., 0], Pq[..., 1]
        W = (4 * np.pi * x * y) ** 2
        sgn = np.sign((nrm * self.enorm[sel]).sum(1))
        pw = w[None, :] * W * self.p_initial(x, y)                       # ds / |F| = w
        tr = self._edge_trace(s)
        out = np.zeros(self.N)
        for m in range(3):
            np.add.at(out, self.np_ + 3 * sel + m, (pw * tr[m][None, :]).sum(1) * sgn)
        self._nat = out
        return out

    def bcs(self):
        fc, mesh = self.fc, self.mesh
        left = np.nonzero((fc.cell1 < 0) & (self.markers == self.LEFT))[0]
        sd = (self.np_ + 3 * left[:, None] + np.arange(3)[None, :]).ravel()
        right = np.nonzero((fc.cell1 < 0) & (self.markers == self.RIGHT))[0]
        va, vb = fc.verts[right, 0], fc.verts[right, 1]
        XA, XB = mesh.coords[va], mesh.coords[vb]
        dofs = [self.q0 + va, self.q0 + vb, self.q0 + self.V + 2 * right, self.q0 + self.V + 2 * right + 1]
        vals = [self.q_initial(XA[:, 0]), self.q_initial(XB[:, 0]),
                self.q_initial((2 * XA[:, 0] + XB[:, 0]) / 3), self.q_initial((XA[:, 0] + 2 * XB[:, 0]) / 3)]
        qd, qi = np.unique(np.concatenate(dofs), return_index=True)
        qv = np.concatenate(vals)[qi]
        return np.concatenate([sd, qd]), np.concatenate([np.zeros(sd.size), qv])

    def solve(self, tol=1e-8, maxit=10, u0=None, p_old=None):
        dofs, g = self.bcs()
        u = np.zeros(self.N) if u0 is None else u0.copy()
        _, b = self.assemble(u, False, p_old)
        b[dofs] = u[dofs] - g
        r0 = r = np.linalg.norm(b)
        it = 0
        while not (r < tol or r / r0 < tol) and it < maxit:
            J, _ = self.assemble(u, True, p_old)
            J = J.tolil()
            J[dofs, :] = 0.0
            J[dofs, dofs] = 1.0
            u -= spla.splu(J.tocsc()).solve(b)
            it += 1
            _, b = self.assemble(u, False, p_old)
            b[dofs] = u[dofs] - g
            r = np.linalg.norm(b)
        self.newton_its, self.converged = it, bool(r < tol or r / r0 < tol)
        return u

    def flux(self, u):
        """project(D 4 pi r2^2 s_h, RT3) (:243): unweighted L2 projection; returns the RT3 coefficient vector."""
        R, _, _ = self.rt_basis(self.lam, self.P)
        a = self.wts[None, :] * (2 * self.area)[:, None]
        x = self.P[:, :, 0]
        s = np.einsum("cqad,ca->cqd", R, u[self.sd])
        f = 4 * np.pi * (x * x)[..., None] * s * np.array([self.D2, self.D1])[None, None, :]
        Me = np.einsum("cq,cqad,cqbd->cab", a, R, R)
        be = np.einsum("cq,cqad,cqd->ca", a, R, f)
        sl = self.sd - self.np_
        M = sp.coo_matrix((Me.ravel(), (np.repeat(sl, 15, axis=1).ravel(), np.tile(sl, (1, 15)).ravel())),
                          shape=(self.ns, self.ns)).tocsc()
        rhs = np.zeros(self.ns)
        np.add.at(rhs, sl.ravel(), be.ravel())
        return spla.splu(M).solve(rhs)

    def bottom_flux(self, fl):
        """V1 * assemble(dot(flux, n) 4 pi r1^2 ds(bottom)) (:261) and its (r2, r1) parts for an RT3 vector fl.
        On a facet only its own three functions contribute to flux.n; the r1 / r2 parts need the full vector."""
        mesh, fc = self.mesh, self.fc
        sel = np.nonzero((fc.cell1 < 0) & (self.markers == self.BOTTOM))[0]
        L, nrm, A, B = fem.facet_geometry(mesh, sel)
        gp, gw = np.polynomial.legendre.leggauss(FACET_POINTS)
        s, w = 0.5 * (gp + 1), 0.5 * gw
        cell = fc.cell0[sel]
        cv = mesh.cells[cell]
        va = fc.verts[sel, 0][:, None] == cv
        vb = fc.verts[sel, 1][:, None] == cv
        lam = va[:, None, :] * (1.0 - s)[None, :, None] + vb[:, None, :] * s[None, :, None]
        Pq = A[:, None, :] + s[None, :, None] * (B - A)[:, None, :]
        R, _, _ = self.rt_basis(lam, Pq, cell)
        f = np.einsum("nqad,na->nqd", R, fl[self.sd[cell] - self.np_])
        wgt = w[None, :] * L[:, None] * 4 * np.pi * Pq[..., 1] ** 2
        r2 = self.V1 * float((wgt * f[..., 0] * nrm[:, None, 0]).sum())
        r1 = self.V1 * float((wgt * f[..., 1] * nrm[:, None, 1]).sum())
        return r2, r1, r2 + r1
