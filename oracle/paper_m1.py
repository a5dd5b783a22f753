"""Oracle for the k=1 "paper" system RT2 x DG1 x CG2 of convergence/test_for_paper_convergence.py (deg=1),
whose two recorded error tables are at :209-225 (test infrastructure; SURVEY section 8f row 1, second half).

Same weak form as oracle/paper_m0.py (:137-153) on the next element family:
  s in RT2 (FEniCS numbering of the degree: P1^2 + x P1~, 8 per cell), p in DG1 (3 per cell), q in CG2 (6 per cell).
Bases (any basis of the same spaces gives the same discrete solution; FIAT's nodal bases are not needed):
  RT2: for the facet i opposite local vertex i with end points j < k: lambda_j N_i, lambda_k N_i with the RT0
       function N_i = sigma_i (x - X_i) / (2|K|) (normal trace lambda/|F| on its own facet, zero on the others:
       H(div)-conforming with one dof per (facet, end point)); interior: lambda_1 N_1, lambda_2 N_2 (unsigned).
  DG1: the barycentric coordinates.   CG2: lambda_i (2 lambda_i - 1) at the vertices, 4 lambda_j lambda_k on the edges.
Global dofs [2 per edge | 2 per cell | 3 per cell | vertices | edges]: 3E + 5C + V, the DoF column of the table.
Quadrature degree 6 everywhere (:38), stab2 = 1.0*(deg+1) = 2 (:71), stab = 0.1 as the comment at :70 says
("0.05 for k=0 // 0.1 for k=1"; the first recorded k=1 table is reproduced with it), hK = FacetArea (:97).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import fem
from .fem import DOLFIN_EPS
from .paper_m0 import D1, D2, _sym, p_ex, q_ex

DEG = 1
OTHER = ((1, 2), (0, 2), (0, 1))          # end points (local vertices, ascending) of the facet opposite vertex i


def p2_basis(lam, g):
    """lam [..., 3] barycentric coordinates, g [C, 3, 2] their gradients -> values [..., 6], gradients [..., 6, 2]
    (leading dims of lam must broadcast against [C, nq])."""
    val = np.concatenate([lam * (2 * lam - 1),
                          np.stack([4 * lam[..., j] * lam[..., k] for j, k in OTHER], -1)], -1)
    gv = (4 * lam - 1)[..., None] * g[:, None, :, :]
    ge = np.stack([4 * (lam[..., j, None] * g[:, None, k, :] + lam[..., k, None] * g[:, None, j, :])
                   for j, k in OTHER], -2)
    return val, np.concatenate([gv, ge], -2)


class PaperM1:
    def __init__(self, n, stab=0.1, stab2=1.0 * (DEG + 1)):
        self.mesh = mesh = fem.unit_square_mesh(n, n)
        self.fc = fc = mesh.facets()
        self.E, self.C, self.V = fc.verts.shape[0], mesh.num_cells, mesh.num_vertices
        E, C, V = self.E, self.C, self.V
        self.N = 3 * E + 5 * C + V
        self.stab, self.stab2 = stab, stab2
        self.gamma = stab / (DEG + 1) ** 3.5                                          # :144
        self.sym = _sym()
        self.D1, self.D2, self.gcoef, self.qmode, self.q_react, self.degree = D1, D2, 1.0, 0, 0.0, 6
        self.q_x2p, self.q_x2q, self.supg_stab = 1.0, 0.0, 0.0
        self._finish_setup()

    nat_markers = (31, 32)     # exterior facets carrying the natural p data (None: all of them)
    h_cell = False             # avg(hK) with hK = CellDiameter instead of FacetArea
    r2dir = 1.0                # r2vec = (r2dir, 0) in the q-equation (-1: ..._C0IP_otherBCs.py:58)

    def _finish_setup(self):
        mesh, fc = self.mesh, self.fc
        E, C, V = self.E, self.C, self.V
        X = mesh.coords
        left = lambda P: fem.near(P[..., 0], 0.0)
        right = lambda P: fem.near(P[..., 0], 1.0)
        rest = lambda P: fem.near(P[..., 1], 0.0) | fem.near(P[..., 1], 1.0)
        self.markers = fem.mark_facets(mesh, [(left, 30), (right, 31), (rest, 32)])
        cells = mesh.cells.astype(np.int64)
        fa = np.stack([cells[:, 1], cells[:, 0], cells[:, 0]], 1)
        fb = np.stack([cells[:, 2], cells[:, 2], cells[:, 1]], 1)
        key = np.minimum(fa, fb) * V + np.maximum(fa, fb)
        allkey = fc.verts[:, 0].astype(np.int64) * V + fc.verts[:, 1]
        self.cell_edges = ce = np.searchsorted(allkey, key)                           # [C,3]
        A, B = X[fc.verts[:, 0]], X[fc.verts[:, 1]]
        t = B - A
        self.elen = np.sqrt((t ** 2).sum(1))
        self.enorm = np.stack([t[:, 1], -t[:, 0]], 1) / self.elen[:, None]
        det, grads, Xc = fem.cell_geometry(mesh)
        self.area, self.Xc, self.grads = 0.5 * np.abs(det), Xc, grads
        mid = 0.5 * (A + B)[ce]
        self.sign = np.sign(((mid - Xc) * self.enorm[ce]).sum(-1))                    # [C,3]
        self.pts, self.wts = fem.triangle_rule(self.degree)
        self.lam = np.stack([1 - self.pts[:, 0] - self.pts[:, 1], self.pts[:, 0], self.pts[:, 1]], 1)   # [nq,3]
        self.P = np.einsum("qi,cid->cqd", self.lam, Xc)
        # global dof numbers of the 17 local dofs
        cid = np.arange(C)[:, None]
        sd = np.concatenate([np.stack([2 * ce[:, i], 2 * ce[:, i] + 1], 1) for i in range(3)]
                            + [2 * E + 2 * cid, 2 * E + 2 * cid + 1], 1)                 # [C,8]
        pd = 2 * E + 2 * C + 3 * cid + np.arange(3)[None, :]
        self.q0 = 2 * E + 5 * C
        qd = np.concatenate([self.q0 + cells, self.q0 + V + ce], 1)                    # [C,6]
        self.sd, self.pd, self.qd = sd, pd, qd
        self.cell_dofs = np.concatenate([sd, pd, qd], 1)                              # [C,17]
        ext = fc.exterior
        m = self.markers[ext]
        self.ext = ext
        self.k_adv = ((m == 30) | (m == 32)).astype(float)
        self.k_dat = (m == 31).astype(float)

    # -- bases at points given by barycentric coordinates lam [C, nq, 3] (or [nq, 3]) ------------------
    def rt_basis(self, lam, P, cells=None):
        """values R [C,nq,8,2], d/dx of the x component and d/dy of the y component [C,nq,8]."""
        idx = slice(None) if cells is None else cells
        Xc, g, area, sign = self.Xc[idx], self.grads[idx], self.area[idx], self.sign[idx]
        if lam.ndim == 2:
            lam = np.broadcast_to(lam[None], (Xc.shape[0],) + lam.shape)
        R, dxx, dyy = [], [], []
        funcs = [(i, a, sign[:, i]) for i in range(3) for a in OTHER[i]] + [(1, 1, None), (2, 2, None)]
        for i, a, sg in funcs:
            c = (1.0 if sg is None else sg) / (2 * area)                                # [C]
            d = P - Xc[:, None, i, :]                                                  # [C,nq,2]
            la = lam[..., a]
            R.append(la[..., None] * d * c[:, None, None])
            dxx.append(c[:, None] * (g[:, None, a, 0] * d[..., 0] + la))
            dyy.append(c[:, None] * (g[:, None, a, 1] * d[..., 1] + la))
        return np.stack(R, 2), np.stack(dxx, 2), np.stack(dyy, 2)

    def assemble(self, u, want_jac=True):
        mesh, C, D1, D2 = self.mesh, self.C, self.D1, self.D2
        lam, P = self.lam, self.P
        x, y = P[:, :, 0], P[:, :, 1]
        W = (4 * np.pi * x * y) ** 2
        a = self.wts[None, :] * (2 * self.area)[:, None] * W
        R, dxx, dyy = self.rt_basis(lam, P)
        Rx, Ry = R[..., 0], R[..., 1]
        dr = dxx + dyy + 2 * Rx / x[..., None] + 2 * Ry / y[..., None]                # div_rad(Psi)
        drD = D2 * (dxx + 2 * Rx / x[..., None]) + D1 * (dyy + 2 * Ry / y[..., None])  # div_rad(D Psi)
        chi = np.broadcast_to(lam[None], (C,) + lam.shape)                             # DG1
        lamc = chi
        phi, gphi = p2_basis(lamc, self.grads)
        se, pe, qe = u[self.sd], u[self.pd], u[self.qd]
        s = np.einsum("cqad,ca->cqd", R, se)
        p = np.einsum("cqb,cb->cq", chi, pe)
        q = np.einsum("cqk,ck->cq", phi, qe)
        x2 = x * x
        iq = 1.0 / (q + DOLFIN_EPS)
        gc = self.gcoef
        G, Gp, Gq = gc * x2 * p * p * iq, 2 * gc * x2 * p * iq, -gc * x2 * p * p * iq * iq
        f2, f3 = self.sym["f2"](x, y), self.sym["f3"](x, y)
        Fe = np.zeros((C, 17))
        Fe[:, :8] = np.einsum("cq,cqa->ca", a, np.einsum("cqd,cqad->cqa", s, R) + G[..., None] * Rx
                              - p[..., None] * dr)
        Fe[:, 8:11] = np.einsum("cq,cqb->cb", a * (-np.einsum("cqa,ca->cq", drD, se) + f2), chi)
        rd = self.r2dir                 # r2vec = (rd, 0)
        if self.qmode == 0:
            Fe[:, 11:] = (np.einsum("cq,cqk->ck", a * ((self.q_react - rd * 2 / x) * q + x2 * self.q_x2p * p - f3), phi)
                          - rd * np.einsum("cq,cqk->ck", a * q, gphi[..., 0]))
        else:   # (q_react q + dq/dx + x^2 (x2p p + x2q q) - f3)(w + tau dw/dx) W, tau = supg_stab hK/2
            qx = np.einsum("cqk,ck->cq", gphi[..., 0], qe)
            tw = phi + (0.5 * self.supg_stab * fem.cell_diameter(self.mesh))[:, None, None] * gphi[..., 0]
            res = self.q_react * q + qx + x2 * (self.q_x2p * p + self.q_x2q * q) - f3
            Fe[:, 11:] = np.einsum("cq,cqk->ck", a * res, tw)
        b = np.zeros(self.N)
        np.add.at(b, self.cell_dofs.ravel(), Fe.ravel())
        rows, cols, vals = [], [], []
        if want_jac:
            Je = np.zeros((C, 17, 17))
            Je[:, :8, :8] = np.einsum("cq,cqad,cqbd->cab", a, R, R)
            Je[:, :8, 8:11] = np.einsum("cq,cqa,cqb->cab", a, Gp[..., None] * Rx - dr, chi)
            Je[:, :8, 11:] = np.einsum("cq,cqa,cqk->cak", a * Gq, Rx, phi)
            Je[:, 8:11, :8] = -np.einsum("cq,cqb,cqa->cba", a, chi, drD)
            if self.qmode == 0:
                Je[:, 11:, 8:11] = np.einsum("cq,cqk,cqb->ckb", a * x2 * self.q_x2p, phi, chi)
                Je[:, 11:, 11:] = (np.einsum("cq,cqk,cql->ckl", a * (self.q_react - rd * 2 / x), phi, phi)
                                   - rd * np.einsum("cq,cqk,cql->ckl", a, gphi[..., 0], phi))
            else:
                Je[:, 11:, 8:11] = np.einsum("cq,cqk,cqb->ckb", a * x2 * self.q_x2p, tw, chi)
                Je[:, 11:, 11:] = np.einsum("cq,cqk,cql->ckl", a, tw,
                                            (self.q_react + x2 * self.q_x2q)[..., None] * phi + gphi[..., 0])
            rows.append(np.repeat(self.cell_dofs, 17, axis=1).ravel())
            cols.append(np.tile(self.cell_dofs, (1, 17)).ravel())
            vals.append(Je.ravel())
        self._exterior(u, b, rows, cols, vals, want_jac)
        self._interior(u, b, rows, cols, vals, want_jac)
        J = None
        if want_jac:
            J = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                              shape=(self.N, self.N)).tocsr()
        return J, b

    def _facet_lam(self, fidx, cell, s):
        """barycentric coordinates [n,nq,3] in `cell` of the points A + s (B - A) of facets fidx."""
        fc, cv = self.fc, self.mesh.cells[cell]
        va = fc.verts[fidx, 0][:, None] == cv
        vb = fc.verts[fidx, 1][:, None] == cv
        return va[:, None, :] * (1.0 - s)[None, :, None] + vb[:, None, :] * s[None, :, None]

    def _exterior(self, u, b, rows, cols, vals, want_jac):
        """q-rows: (r2vec.n) q w W ds(left,rest) + (r2vec.n) q_ex w W ds(right) + stab2 (hK neg(r2vec.n)^2 + 8/hK)
        (q - q_ex) w ds (:142-145,152-153); s-rows: + p_ex (tau.n) W ds(right,rest) (:148-149)."""
        mesh, fc, ext = self.mesh, self.fc, self.ext
        L, nrm, A, B = fem.facet_geometry(mesh, ext)
        s, w = fem.facet_rule(self.degree)
        cell = fc.cell0[ext]
        lam = self._facet_lam(ext, cell, s)
        Pq = A[:, None, :] + s[None, :, None] * (B - A)[:, None, :]
        x, y = Pq[..., 0], Pq[..., 1]
        W = (4 * np.pi * x * y) ** 2
        phi, _ = p2_basis(lam, self.grads[cell])
        qd = self.qd[cell]
        qh = np.einsum("nqk,nk->nq", phi, u[qd])
        bn = self.r2dir * nrm[:, 0]                                                  # r2vec = (+-1,0)
        neg = 0.5 * (np.abs(bn) - bn)
        pen = self.stab2 * (L * neg ** 2 + 8.0 / L)
        ds = w[None, :] * L[:, None]
        cq = (self.k_adv * bn)[:, None] * W + pen[:, None]
        cd = (self.k_dat * bn)[:, None] * W - pen[:, None]
        Fx = np.einsum("nq,nqk->nk", ds * (cq * qh + cd * q_ex(x, y)), phi)
        np.add.at(b, qd.ravel(), Fx.ravel())
        if want_jac:
            Jx = np.einsum("nq,nqk,nql->nkl", ds * cq, phi, phi)
            rows.append(np.repeat(qd, 6, axis=1).ravel())
            cols.append(np.tile(qd, (1, 6)).ravel())
            vals.append(Jx.ravel())
        # natural p data on the s-rows of the right / rest facets: only the two functions of the facet itself
        # have a normal trace there, lambda_end / |F| * sign_out
        sel = (np.ones(ext.size, dtype=bool) if self.nat_markers is None
               else np.isin(self.markers[ext], self.nat_markers))
        sgn = np.sign((nrm * self.enorm[ext]).sum(1))
        pw = ds * W * p_ex(x, y)                                                     # [n,nq]
        for slot, lamend in ((0, 1.0 - s), (1, s)):
            val = (pw * lamend[None, :]).sum(1) * sgn / L
            np.add.at(b, 2 * ext[sel] + slot, val[sel])

    def _interior(self, u, b, rows, cols, vals, want_jac):
        """gamma avg(hK)^2 (jump(grad q).n+)(jump(grad w).n+) W dS, hK = FacetArea (:144)."""
        mesh, fc = self.mesh, self.fc
        fi = fc.interior
        L, nrm, A, B = fem.facet_geometry(mesh, fi)
        s, w = fem.facet_rule(self.degree)
        Pq = A[:, None, :] + s[None, :, None] * (B - A)[:, None, :]
        W = (4 * np.pi * Pq[..., 0] * Pq[..., 1]) ** 2
        c0, c1 = fc.cell0[fi], fc.cell1[fi]
        d = []
        for cell, sg in ((c0, 1.0), (c1, -1.0)):
            lam = self._facet_lam(fi, cell, s)
            _, gphi = p2_basis(lam, self.grads[cell])
            d.append(sg * np.einsum("nqkd,nd->nqk", gphi, nrm))
        d = np.concatenate(d, 2)                                                     # [Fi,nq,12]
        md = np.concatenate([self.qd[c0], self.qd[c1]], 1)
        hbar = L
        if self.h_cell:
            hK = fem.cell_diameter(mesh)
            hbar = 0.5 * (hK[c0] + hK[c1])
        coef = self.gamma * (hbar ** 2)[:, None] * w[None, :] * L[:, None] * W
        jq = np.einsum("nqk,nk->nq", d, u[md])
        np.add.at(b, md.ravel(), np.einsum("nq,nqk->nk", coef * jq, d).ravel())
        if want_jac:
            Ji = np.einsum("nq,nqk,nql->nkl", coef, d, d)
            rows.append(np.repeat(md, 12, axis=1).ravel())
            cols.append(np.tile(md, (1, 12)).ravel())
            vals.append(Ji.ravel())

    def sigD(self):
        """project(sig_ex, RT2) (:132): unweighted L2 projection with the degree-6 rule."""
        R, _, _ = self.rt_basis(self.lam, self.P)
        a = self.wts[None, :] * (2 * self.area)[:, None]
        x, y = self.P[:, :, 0], self.P[:, :, 1]
        sex = np.stack([self.sym["sx"](x, y), self.sym["sy"](x, y)], -1)
        Me = np.einsum("cq,cqad,cqbd->cab", a, R, R)
        be = np.einsum("cq,cqad,cqd->ca", a, R, sex)
        ns = 2 * self.E + 2 * self.C
        M = sp.coo_matrix((Me.ravel(), (np.repeat(self.sd, 8, axis=1).ravel(), np.tile(self.sd, (1, 8)).ravel())),
                          shape=(ns, ns)).tocsc()
        rhs = np.zeros(ns)
        np.add.at(rhs, self.sd.ravel(), be.ravel())
        return spla.splu(M).solve(rhs)

    def solve(self, tol=1e-7, maxit=50):
        le = np.nonzero(self.markers == 30)[0]
        bc_dofs = np.concatenate([2 * le, 2 * le + 1])                                # both dofs of the left edges (:133)
        g = self.sigD()[bc_dofs]
        u = np.zeros(self.N)
        _, b = self.assemble(u, False)
        b[bc_dofs] = u[bc_dofs] - g
        r0 = r = np.linalg.norm(b)
        it = 0
        while not (r < tol or r / r0 < tol) and it < maxit:
            J, _ = self.assemble(u, True)
            J = J.tolil()
            J[bc_dofs, :] = 0.0
            J[bc_dofs, bc_dofs] = 1.0
            u -= spla.splu(J.tocsc()).solve(b)
            it += 1
            _, b = self.assemble(u, False)
            b[bc_dofs] = u[bc_dofs] - g
            r = np.linalg.norm(b)
        self.newton_its = it
        return u

    def errors(self, u):
        """(:175-177) e_div(s), e_0(p), e_1(q): weighted norms, degree-6 rule."""
        lam, P = self.lam, self.P
        x, y = P[:, :, 0], P[:, :, 1]
        W = (4 * np.pi * x * y) ** 2
        a = self.wts[None, :] * (2 * self.area)[:, None] * W
        R, dxx, dyy = self.rt_basis(lam, P)
        dr = dxx + dyy + 2 * R[..., 0] / x[..., None] + 2 * R[..., 1] / y[..., None]
        se = u[self.sd]
        s = np.einsum("cqad,ca->cqd", R, se)
        drs = np.einsum("cqa,ca->cq", dr, se)
        sx, sy = self.sym["sx"](x, y), self.sym["sy"](x, y)
        es = np.sqrt((a * ((sx - s[..., 0]) ** 2 + (sy - s[..., 1]) ** 2 + (self.sym["drs"](x, y) - drs) ** 2)).sum())
        ph = np.einsum("qb,cb->cq", lam, u[self.pd])
        ep = np.sqrt((a * (p_ex(x, y) - ph) ** 2).sum())
        phi, gphi = p2_basis(np.broadcast_to(lam[None], (self.C,) + lam.shape), self.grads)
        qh = np.einsum("cqk,ck->cq", phi, u[self.qd])
        qhx = np.einsum("cqk,ck->cq", gphi[..., 0], u[self.qd])
        eq = np.sqrt((a * ((q_ex(x, y) - qh) ** 2 + (self.sym["qx"](x, y) - qhx) ** 2)).sum())
        return float(es), float(ep), float(eq)


class MixedPrimalM1(PaperM1):
    """convergence/SphericalAdvectionDiffusionReaction_MixedPrimal_convergence.py (BASELINE config 2's own script):
    RT2 x DG1 x CG2 (deg = 1, :93-97), D1 = 1e-3, D2 = 1.5e-3 (:51-52), G = D2 4 pi r2^2 p^2/(q + eps) (:43-44),
    Galerkin q-equation (q + r2vec.grad q + r2^2 p) w W (:123), natural p data on the whole boundary (:126),
    q = q_ex on the facets marked 'left' = {x = 1} (:78,117), quadrature degree 4 (:25), Newton 1e-7 with 'umfpack'
    (:132-138); q error with the whole gradient (:148).  No output of this script is recorded: parity unpinned."""

    def __init__(self, n, D1=1e-3, D2=1.5e-3, supg=False):
        """supg=True: ...MixedPrimal_convergence_SUPG.py (:100-129): residual q + dq/dx + r2^2 q ("r2**2*q" as written
        there) tested with w + stab hK/2 dw/dx (stab = 1, hK = CellDiameter), q = q_ex on {x = 0}, Newton 1e-8."""
        self.supg = supg
        self.mesh = mesh = fem.unit_square_mesh(n, n)
        self.fc = fc = mesh.facets()
        self.E, self.C, self.V = fc.verts.shape[0], mesh.num_cells, mesh.num_vertices
        self.N = 3 * self.E + 5 * self.C + self.V
        self.stab = self.stab2 = self.gamma = 0.0
        self.D1, self.D2, self.gcoef, self.qmode, self.q_react, self.degree = D1, D2, D2 * 4 * np.pi, 1, 1.0, 4
        self.q_x2p, self.q_x2q, self.supg_stab = (0.0, 1.0, 1.0) if supg else (1.0, 0.0, 0.0)
        self.sym = self._data()
        self._finish_setup()
        xin = 0.0 if supg else 1.0            # the facets carrying the q data: {x = 0} (SUPG :77) or {x = 1} (:78)
        self.markers = fem.mark_facets(mesh, [(lambda P: fem.near(P[..., 0], xin), 31),
                                              (lambda P: fem.near(P[..., 0], 1.0 - xin) | fem.near(P[..., 1], 0.0)
                                               | fem.near(P[..., 1], 1.0), 32)])

    def _data(self):
        import sympy as sm
        x, y = sm.symbols("x y", positive=True)
        pe = sm.sin(sm.pi * x) * sm.sin(sm.pi * y)
        qe = sm.Rational(3, 2) + sm.cos(sm.pi * x * y)
        G = self.gcoef * x ** 2 * pe ** 2 / (qe + DOLFIN_EPS)
        sx, sy = -sm.diff(pe, x) - G, -sm.diff(pe, y)
        f2 = sm.diff(x ** 2 * self.D2 * sx, x) / x ** 2 + sm.diff(y ** 2 * self.D1 * sy, y) / y ** 2
        f3 = qe + sm.diff(qe, x) + x ** 2 * (qe if self.supg else pe)
        drs = sm.diff(x ** 2 * sx, x) / x ** 2 + sm.diff(y ** 2 * sy, y) / y ** 2
        L = lambda e: sm.lambdify((x, y), e, "numpy")
        return dict(sx=L(sx), sy=L(sy), f2=L(f2), f3=L(f3), drs=L(drs), qx=L(sm.diff(qe, x)), qy=L(sm.diff(qe, y)))

    def _exterior(self, u, b, rows, cols, vals, want_jac):
        """+ p_ex (tau.n) W ds on every exterior facet (:126)."""
        mesh, ext = self.mesh, self.ext
        L, nrm, A, B = fem.facet_geometry(mesh, ext)
        s, w = fem.facet_rule(self.degree)
        Pq = A[:, None, :] + s[None, :, None] * (B - A)[:, None, :]
        x, y = Pq[..., 0], Pq[..., 1]
        pw = w[None, :] * (4 * np.pi * x * y) ** 2 * p_ex(x, y)
        sgn = np.sign((nrm * self.enorm[ext]).sum(1))
        for slot, lamend in ((0, 1.0 - s), (1, s)):
            np.add.at(b, 2 * ext + slot, (pw * lamend[None, :]).sum(1) * sgn)

    def _interior(self, u, b, rows, cols, vals, want_jac):
        pass

    def bcs(self):
        fc, mesh = self.fc, self.mesh
        sel = np.nonzero((fc.cell1 < 0) & (self.markers == 31))[0]
        va, vb = fc.verts[sel, 0], fc.verts[sel, 1]
        XA, XB = mesh.coords[va], mesh.coords[vb]
        XM = 0.5 * (XA + XB)
        dofs = np.concatenate([self.q0 + va, self.q0 + vb, self.q0 + self.V + sel])
        vals = np.concatenate([q_ex(XA[:, 0], XA[:, 1]), q_ex(XB[:, 0], XB[:, 1]), q_ex(XM[:, 0], XM[:, 1])])
        d, i = np.unique(dofs, return_index=True)
        return d, vals[i]

    def solve(self, tol=None, maxit=50):
        tol = (1e-8 if self.supg else 1e-7) if tol is None else tol
        bc_dofs, g = self.bcs()
        u = np.zeros(self.N)
        _, b = self.assemble(u, False)
        b[bc_dofs] = u[bc_dofs] - g
        r0 = r = np.linalg.norm(b)
        it = 0
        while not (r < tol or r / r0 < tol) and it < maxit:
            J, _ = self.assemble(u, True)
            J = J.tolil()
            J[bc_dofs, :] = 0.0
            J[bc_dofs, bc_dofs] = 1.0
            u -= spla.splu(J.tocsc()).solve(b)
            it += 1
            _, b = self.assemble(u, False)
            b[bc_dofs] = u[bc_dofs] - g
            r = np.linalg.norm(b)
        self.newton_its = it
        return u

    def errors(self, u):
        es, ep, _ = PaperM1.errors(self, u)
        lam, P = self.lam, self.P
        x, y = P[:, :, 0], P[:, :, 1]
        a = self.wts[None, :] * (2 * self.area)[:, None] * (4 * np.pi * x * y) ** 2
        phi, gphi = p2_basis(np.broadcast_to(lam[None], (self.C,) + lam.shape), self.grads)
        qe = u[self.qd]
        qh = np.einsum("cqk,ck->cq", phi, qe)
        qhx = np.einsum("cqk,ck->cq", gphi[..., 0], qe)
        qhy = np.einsum("cqk,ck->cq", gphi[..., 1], qe)
        eq = np.sqrt((a * ((q_ex(x, y) - qh) ** 2 + (self.sym["qx"](x, y) - qhx) ** 2
                           + (self.sym["qy"](x, y) - qhy) ** 2)).sum())
        return es, ep, float(eq)


class MixedPrimalC0IPM1(MixedPrimalM1):
    """convergence/SphericalAdvectionDiffusionReaction_MixedPrimal_convergence_C0IP.py on RT2 x DG1 x CG2:
    D1 = .01, D2 = .015 (:63-64), G = r2^2 p^2/(q + eps) (:46-47), the q-equation integrated by parts with the
    inflow term on {x = 0} (:144-146) and the C0IP jump penalty with the MINUS sign, stab = 0.00075,
    hK = CellDiameter (:68,98,147); natural p data on the whole boundary (:149); essential: q = q_ex on {x = 1},
    s = s_ex = (0, pi sin(pi x)) on {x = 0}, whose normal trace there is zero (:132-135); Newton 1e-8 (:160-161)."""
    nat_markers = None
    h_cell = True

    def __init__(self, n, D1=0.01, D2=0.015, stab=0.00075, degree=4):
        self.supg = False
        self.mesh = mesh = fem.unit_square_mesh(n, n)
        self.fc = fc = mesh.facets()
        self.E, self.C, self.V = fc.verts.shape[0], mesh.num_cells, mesh.num_vertices
        self.N = 3 * self.E + 5 * self.C + self.V
        self.stab, self.stab2 = stab, 0.0
        self.gamma = -stab / (DEG + 1) ** 3.5
        self.D1, self.D2, self.gcoef, self.qmode, self.q_react, self.degree = D1, D2, 1.0, 0, 0.0, degree
        self.q_x2p, self.q_x2q, self.supg_stab = 1.0, 0.0, 0.0
        self.sym = self._data()
        self._finish_setup()
        self.markers = fem.mark_facets(mesh, [(lambda P: fem.near(P[..., 0], 0.0), 31),
                                              (lambda P: fem.near(P[..., 0], 1.0), 32),
                                              (lambda P: fem.near(P[..., 1], 0.0) | fem.near(P[..., 1], 1.0), 33)])
        m = self.markers[self.ext]
        self.k_adv = (m == 31).astype(float)
        self.k_dat = np.zeros(m.size)

    def _data(self):
        d = MixedPrimalM1._data(self)
        import sympy as sm
        x, y = sm.symbols("x y", positive=True)
        pe = sm.sin(sm.pi * x) * sm.sin(sm.pi * y)
        qe = sm.Rational(3, 2) + sm.cos(sm.pi * x * y)
        d["f3"] = sm.lambdify((x, y), sm.diff(qe, x) + x ** 2 * pe, "numpy")          # :130
        return d

    _exterior = PaperM1._exterior
    _interior = PaperM1._interior

    def bcs(self):
        fc, mesh = self.fc, self.mesh
        sel = np.nonzero((fc.cell1 < 0) & (self.markers == 32))[0]
        va, vb = fc.verts[sel, 0], fc.verts[sel, 1]
        XA, XB = mesh.coords[va], mesh.coords[vb]
        XM = 0.5 * (XA + XB)
        dofs = np.concatenate([self.q0 + va, self.q0 + vb, self.q0 + self.V + sel])
        vals = np.concatenate([q_ex(XA[:, 0], XA[:, 1]), q_ex(XB[:, 0], XB[:, 1]), q_ex(XM[:, 0], XM[:, 1])])
        d, i = np.unique(dofs, return_index=True)
        le = np.nonzero((fc.cell1 < 0) & (self.markers == 31))[0]
        sd = np.concatenate([2 * le, 2 * le + 1])
        return np.concatenate([sd, d]), np.concatenate([np.zeros(sd.size), vals[i]])

    def solve(self, tol=1e-8, maxit=50):
        return MixedPrimalM1.solve(self, tol, maxit)


class MixedPrimalC0IPOtherM1(MixedPrimalC0IPM1):
    """convergence/SphericalAdvectionDiffusionReaction_MixedPrimal_convergence_C0IP_otherBCs.py on RT2 x DG1 x CG2:
    r2vec = (-1,0) in the q-equation (:58; G keeps Constant((1,0)), :46-47), D1 = 2, D2 = 1.5 (:55-56), penalty with the
    PLUS sign, stab = 0.0075 (:60,134), hK = CellDiameter (:85); markers left 30 / right 31 / rest 32 (:78-82); inflow
    term on the left (:133), natural p data on right and rest (:136-137); essential: s = project(sig_ex) on the left,
    q = q_ex on the right (:121-124); f3 = grad(q_ex).r2vec + r2^2 p_ex (:117); Newton 1e-8 (:149-150)."""
    nat_markers = (31, 32)
    r2dir = -1.0

    def __init__(self, n, D1=2.0, D2=1.5, stab=0.0075):
        MixedPrimalC0IPM1.__init__(self, n, D1, D2, -stab)         # the parent stores gamma = -stab/(deg+1)^3.5
        self.markers = fem.mark_facets(self.mesh, [(lambda P: fem.near(P[..., 0], 0.0), 30),
                                                   (lambda P: fem.near(P[..., 0], 1.0), 31),
                                                   (lambda P: fem.near(P[..., 1], 0.0) | fem.near(P[..., 1], 1.0), 32)])
        self.k_adv = (self.markers[self.ext] == 30).astype(float)

    def _data(self):
        d = MixedPrimalM1._data(self)
        import sympy as sm
        x, y = sm.symbols("x y", positive=True)
        pe = sm.sin(sm.pi * x) * sm.sin(sm.pi * y)
        qe = sm.Rational(3, 2) + sm.cos(sm.pi * x * y)
        d["f3"] = sm.lambdify((x, y), -sm.diff(qe, x) + x ** 2 * pe, "numpy")
        return d

    def bcs(self):
        dq, gq = MixedPrimalM1.bcs(self)                           # q on the facets marked 31
        le = np.nonzero((self.fc.cell1 < 0) & (self.markers == 30))[0]
        sd = np.concatenate([2 * le, 2 * le + 1])
        return np.concatenate([sd, dq]), np.concatenate([self.sigD()[sd], gq])


class MixedPrimalC0IPModifiedM1(MixedPrimalC0IPM1):
    """convergence/SphericalAdvectionDiffusionReaction_MixedPrimal_convergence_C0IP_otherBCs_modified.py on
    RT2 x DG1 x CG2 (deg = 1 :49): the ancestor of test_for_paper_convergence.py -- r2vec = (1,0), D1 = 2, D2 = 1.5
    (:43-44), penalty +stab with stab = 1e-3 (:47) and hK = FacetArea (:111), inflow term on the left (:156), the
    outflow data of q imposed WEAKLY on the right (:164) and no boundary penalty, natural p data on right and rest
    (:160-161), s = project(sig_ex) on the left as the only essential condition (:146-147), quadrature degree 6 (:38),
    SNES with tolerances 1e-7 (:171-174; plain Newton steps here), q error with the full gradient (:188)."""
    nat_markers = (31, 32)
    h_cell = False

    def __init__(self, n, D1=2.0, D2=1.5, stab=1e-3):
        MixedPrimalC0IPM1.__init__(self, n, D1, D2, -stab, degree=6)
        self.markers = fem.mark_facets(self.mesh, [(lambda P: fem.near(P[..., 0], 0.0), 30),
                                                   (lambda P: fem.near(P[..., 0], 1.0), 31),
                                                   (lambda P: fem.near(P[..., 1], 0.0) | fem.near(P[..., 1], 1.0), 32)])
        m = self.markers[self.ext]
        self.k_adv, self.k_dat = (m == 30).astype(float), (m == 31).astype(float)

    def bcs(self):
        le = np.nonzero((self.fc.cell1 < 0) & (self.markers == 30))[0]
        sd = np.concatenate([2 * le, 2 * le + 1])
        return sd, self.sigD()[sd]

    def solve(self, tol=1e-7, maxit=50):
        return MixedPrimalM1.solve(self, tol, maxit)


def mixed_primal_c0ip_table(levels=5, other_bcs=False, modified=False):
    rows, prev = [], None
    for nk in range(levels):
        m = (MixedPrimalC0IPModifiedM1 if modified else MixedPrimalC0IPOtherM1 if other_bcs
             else MixedPrimalC0IPM1)(2 ** (nk + 1))
        u = m.solve()
        h = m.mesh.hmax()
        e = m.errors(u)
        r = [0.0, 0.0, 0.0] if prev is None else [np.log(a / b) / np.log(h / prev[0]) for a, b in zip(e, prev[1])]
        rows.append('{:6d} & {:.4f} & {:6.2e} & {:.3f} & {:6.2e} & {:.3f} & {:6.2e} & {:.3f} '.format(
            m.N, h, e[0], r[0], e[1], r[1], e[2], r[2]))
        prev = (h, e)
    return rows


def mixed_primal_table(levels=5, supg=False):
    rows, prev = [], None
    for nk in range(levels):
        m = MixedPrimalM1(2 ** (nk + 1), supg=supg)
        u = m.solve()
        h = m.mesh.hmax()
        e = m.errors(u)
        r = [0.0, 0.0, 0.0] if prev is None else [np.log(a / b) / np.log(h / prev[0]) for a, b in zip(e, prev[1])]
        rows.append('{:6d} & {:.4f} & {:6.2e} & {:.3f} & {:6.2e} & {:.3f} & {:6.2e} & {:.3f} '.format(
            m.N, h, e[0], r[0], e[1], r[1], e[2], r[2]))
        prev = (h, e)
    return rows


def table(levels=5, stab=0.1):
    rows, prev = [], None
    for nk in range(levels):
        n = 2 ** (nk + 1)
        m = PaperM1(n, stab=stab)
        u = m.solve()
        h = m.mesh.hmax()
        e = m.errors(u)
        r = [0.0, 0.0, 0.0] if prev is None else [np.log(a / b) / np.log(h / prev[0]) for a, b in zip(e, prev[1])]
        rows.append('{:6d} & {:.4f} & {:6.2e} & {:.3f} & {:6.2e} & {:.3f} & {:6.2e} & {:.3f} '.format(
            m.N, h, e[0], r[0], e[1], r[1], e[2], r[2]))
        prev = (h, e)
    return rows


if __name__ == "__main__":
    import sys
    for row in table(int(sys.argv[1]) if len(sys.argv) > 1 else 5, float(sys.argv[2]) if len(sys.argv) > 2 else 0.1):
        print(row)
