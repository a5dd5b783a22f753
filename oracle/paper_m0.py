"""Oracle for the lowest-order "paper" system RT1 x DG0 x CG1 of
convergence/test_for_paper_convergence.py (deg=0), whose k=0 error table is recorded at :200-207
(test infrastructure).

Why it is here although RT/DG are not on the GPU path yet: q is P1 in this system, and the recorded
table is the only reference output that exercises the P1 interior-facet (C0IP, :144), exterior-facet
(:142-143,152) and boundary-penalty (:145,153) q-terms.  The q-block facet tensors are computed by the
SAME functions the P-Q oracle uses (oracle.pq.element_exterior_facets / element_interior_facets), so
reproducing the table pins them.

Unknowns: s in RT1 (one dof per edge, basis sigma_i (x - x_i)/(2|K|) for the facet opposite local
vertex i), p in DG0 (one per cell), q in P1.  Global dofs [edges | cells | vertices] (DOLFIN's own
numbering of this mixed space is not reproducible without DOLFIN; results are
numbering-independent).  Quadrature degree 6 everywhere (:38), Newton from 0 (SNES "basic" = full
steps), tolerances 1e-7 (:161-164).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import fem, pq
from .fem import DOLFIN_EPS

D1, D2 = 1.0, 1.5                    # :63-64
STAB, STAB2 = 0.05, 1.0              # :70-71 (deg = 0)


def p_ex(x, y):
    return np.sin(np.pi * x) * np.sin(np.pi * y)


def q_ex(x, y):
    return 1.5 + np.cos(np.pi * x * y)


def _sym():
    import sympy as sm
    x, y = sm.symbols("x y", positive=True)
    pe = sm.sin(sm.pi * x) * sm.sin(sm.pi * y)
    qe = sm.Rational(3, 2) + sm.cos(sm.pi * x * y)
    G = x ** 2 * pe ** 2 / (qe + DOLFIN_EPS)
    sx = -sm.diff(pe, x) - G
    sy = -sm.diff(pe, y)
    f2 = sm.diff(x ** 2 * D2 * sx, x) / x ** 2 + sm.diff(y ** 2 * D1 * sy, y) / y ** 2   # div_rad(D*sig_ex)
    f3 = sm.diff(qe, x) + x ** 2 * pe                                                     # :128
    drs = sm.diff(x ** 2 * sx, x) / x ** 2 + sm.diff(y ** 2 * sy, y) / y ** 2             # div_rad(sig_ex)
    L = lambda e: sm.lambdify((x, y), e, "numpy")
    return dict(sx=L(sx), sy=L(sy), f2=L(f2), f3=L(f3), drs=L(drs), qx=L(sm.diff(qe, x)))


class PaperM0:
    def __init__(self, n):
        self.mesh = mesh = fem.unit_square_mesh(n, n)
        self.fc = fc = mesh.facets()
        self.E, self.C, self.V = fc.verts.shape[0], mesh.num_cells, mesh.num_vertices
        self.N = self.E + self.C + self.V
        self.sym = _sym()
        X = mesh.coords
        # markers :90-94
        left = lambda P: fem.near(P[..., 0], 0.0)
        right = lambda P: fem.near(P[..., 0], 1.0)
        rest = lambda P: fem.near(P[..., 1], 0.0) | fem.near(P[..., 1], 1.0)
        self.markers = fem.mark_facets(mesh, [(left, 30), (right, 31), (rest, 32)])
        # cell -> edges (local facet i opposite local vertex i) and orientation signs
        cells = mesh.cells.astype(np.int64)
        fa = np.stack([cells[:, 1], cells[:, 0], cells[:, 0]], 1)
        fb = np.stack([cells[:, 2], cells[:, 2], cells[:, 1]], 1)
        key = np.minimum(fa, fb) * self.V + np.maximum(fa, fb)
        allkey = fc.verts[:, 0].astype(np.int64) * self.V + fc.verts[:, 1]
        self.cell_edges = np.searchsorted(allkey, key)                                   # [C,3]
        # global normal of an edge: tangent (low -> high vertex) rotated clockwise
        A = X[fc.verts[:, 0]]
        B = X[fc.verts[:, 1]]
        t = B - A
        self.elen = np.sqrt((t ** 2).sum(1))
        self.enorm = np.stack([t[:, 1], -t[:, 0]], 1) / self.elen[:, None]
        det, grads, Xc = fem.cell_geometry(mesh)
        self.area = 0.5 * np.abs(det)
        self.Xc, self.grads = Xc, grads
        # sign: outward normal of the cell on local facet i vs the global normal
        mid = 0.5 * (A + B)[self.cell_edges]                                             # [C,3,2]
        out = mid - Xc                                                                   # from opposite vertex
        self.sign = np.sign((out * self.enorm[self.cell_edges]).sum(-1))                 # [C,3]
        self.pts, self.wts = fem.triangle_rule(6)
        self.phi = np.stack([1 - self.pts[:, 0] - self.pts[:, 1], self.pts[:, 0], self.pts[:, 1]], 1)
        self.P = np.einsum("qi,cid->cqd", self.phi, Xc)                                  # [C,nq,2]
        self.prm = pq.PQParams(D1=D1, D2=D2, gkind=pq.G_POWER, gcoef=1.0, q_adv=0.0, q_divr=1.0,
                               q_ibp=1.0, degree=6, c0ip_gamma=STAB / 1.0 ** 3.5, c0ip_h=pq.H_FACET_AREA,
                               pen_stab2=STAB2)
        # exterior-facet kinds of the q-equation (:142-145,152-153)
        ext = fc.exterior
        kinds = np.full(ext.size, pq.EF_PENALTY)
        m = self.markers[ext]
        kinds[(m == 30) | (m == 32)] |= pq.EF_ADV
        kinds[m == 31] |= pq.EF_DATA
        self.ext, self.ext_kinds = ext, kinds
        self._dofs()

    def _dofs(self):
        E, C = self.E, self.C
        ce = self.cell_edges
        cid = np.arange(C)[:, None]
        qd = E + C + self.mesh.cells.astype(np.int64)
        self.cell_dofs = np.concatenate([ce, E + cid, qd], 1)                            # [C,7]

    # RT basis values R[c,q,i,:] and derived quantities at the cell quadrature points
    def _rt(self):
        P, Xc = self.P, self.Xc
        R = (P[:, :, None, :] - Xc[:, None, :, :]) * (self.sign / (2 * self.area[:, None]))[:, None, :, None]
        div = self.sign / self.area[:, None]                                             # [C,3]
        return R, div

    def assemble(self, u, want_jac=True):
        mesh, prm = self.mesh, self.prm
        E, C, V = self.E, self.C, self.V
        P = self.P
        x, y = P[:, :, 0], P[:, :, 1]
        W = (4 * np.pi * x * y) ** 2
        a = self.wts[None, :] * (2 * self.area)[:, None] * W                             # [C,nq]
        R, div = self._rt()
        phi = self.phi
        Gx = self.grads[:, :, 0]
        se = u[self.cell_edges]                                                          # [C,3]
        pc = u[E:E + C]
        qe = u[E + C:][mesh.cells]
        s = np.einsum("cqid,ci->cqd", R, se)                                             # [C,nq,2]
        q = qe @ phi.T
        p = pc[:, None] * np.ones_like(q)
        G, Gp, Gq = pq.eval_G(prm, x, p, q)
        # dr(R_i) = div R_i + 2 R_x/x + 2 R_y/y ; drD(R_i) = D2 dRx/dx + D1 dRy/dy + 2 D2 Rx/x + 2 D1 Ry/y
        dRx = (self.sign / (2 * self.area[:, None]))                                     # d/dx of x-component
        dr = div[:, None, :] + 2 * R[..., 0] / x[:, :, None] + 2 * R[..., 1] / y[:, :, None]
        drD = (D2 + D1) * dRx[:, None, :] + 2 * D2 * R[..., 0] / x[:, :, None] + 2 * D1 * R[..., 1] / y[:, :, None]
        f2 = self.sym["f2"](x, y)
        f3 = self.sym["f3"](x, y)
        Fe = np.zeros((C, 7))
        Fe[:, :3] = np.einsum("cq,cqi->ci", a, np.einsum("cqd,cqid->cqi", s, R) + G[:, :, None] * R[..., 0]
                              - p[:, :, None] * dr)
        drDs = np.einsum("cqi,ci->cq", drD, se)
        Fe[:, 3] = (a * (-drDs + f2)).sum(1)
        Fe[:, 4:] = (np.einsum("cq,qi->ci", a * (-(2 / x) * q + x ** 2 * p - f3), phi)
                     - np.einsum("cq,ci->ci", a * q, Gx))
        b = np.zeros(self.N)
        fem.add_serial(b, self.cell_dofs, Fe)
        rows, cols, vals = [], [], []
        if want_jac:
            Je = np.zeros((C, 7, 7))
            Je[:, :3, :3] = np.einsum("cq,cqid,cqjd->cij", a, R, R)
            Je[:, :3, 3] = np.einsum("cq,cqi->ci", a, Gp[:, :, None] * R[..., 0] - dr)
            Je[:, :3, 4:] = np.einsum("cq,cqi,qj->cij", a * Gq, R[..., 0], phi)
            Je[:, 3, :3] = -np.einsum("cq,cqj->cj", a, drD)
            Je[:, 4:, 3] = np.einsum("cq,qi->ci", a * x ** 2, phi)
            Je[:, 4:, 4:] = (-np.einsum("cq,qi,qj->cij", a * (2 / x), phi, phi)
                             - np.einsum("cq,qj,ci->cij", a, phi, Gx))
            rows.append(np.repeat(self.cell_dofs, 7, axis=1).ravel())
            cols.append(np.tile(self.cell_dofs, (1, 7)).ravel())
            vals.append(Je.ravel())
        # ---- facet terms of the q-equation: the SAME functions as the P-Q oracle ------------------
        uq = np.zeros(2 * V)
        uq[1::2] = u[E + C:]
        cell, Fx, Jx = pq.element_exterior_facets(mesh, uq, prm, self.ext, self.ext_kinds, q_ex)
        qd = E + C + mesh.cells[cell].astype(np.int64)
        fem.add_serial(b, qd, Fx)
        mv, Fi, Ji = pq.element_interior_facets(mesh, uq, prm)
        qi = E + C + mv.astype(np.int64)
        fem.add_serial(b, qi, Fi)
        if want_jac:
            rows += [np.repeat(qd, 3, axis=1).ravel(), np.repeat(qi, 6, axis=1).ravel()]
            cols += [np.tile(qd, (1, 3)).ravel(), np.tile(qi, (1, 6)).ravel()]
            vals += [Jx.ravel(), Ji.ravel()]
        # ---- natural p data on the s-rows: + p_ex (tau.n) W ds(right), ds(rest)  (:148-149) -------
        sel = self.ext[(self.markers[self.ext] == 31) | (self.markers[self.ext] == 32)]
        L, nrm, A, B = fem.facet_geometry(mesh, sel)
        sq, wq = fem.facet_rule(6)
        Pq = A[:, None, :] + sq[None, :, None] * (B - A)[:, None, :]
        Wf = (4 * np.pi * Pq[:, :, 0] * Pq[:, :, 1]) ** 2
        # RT0 normal trace on its own facet: tau.n_global = 1/|F| -> tau.n_out = sign_out/|F|
        sgn = np.sign((nrm * self.enorm[sel]).sum(1))
        b[sel] += ((wq[None, :] * Wf * p_ex(Pq[:, :, 0], Pq[:, :, 1])).sum(1) * L) * sgn / L
        J = None
        if want_jac:
            J = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                              shape=(self.N, self.N)).tocsr()
        return J, b

    def sigD(self):
        """sigD = project(sig_ex, RT) (:132): unweighted L2 projection, degree-6 rule."""
        R, _ = self._rt()
        a = self.wts[None, :] * (2 * self.area)[:, None]
        x, y = self.P[:, :, 0], self.P[:, :, 1]
        sex = np.stack([self.sym["sx"](x, y), self.sym["sy"](x, y)], -1)
        Me = np.einsum("cq,cqid,cqjd->cij", a, R, R)
        be = np.einsum("cq,cqid,cqd->ci", a, R, sex)
        ce = self.cell_edges
        M = sp.coo_matrix((Me.ravel(), (np.repeat(ce, 3, axis=1).ravel(), np.tile(ce, (1, 3)).ravel())),
                          shape=(self.E, self.E)).tocsc()
        rhs = np.zeros(self.E)
        np.add.at(rhs, ce.ravel(), be.ravel())
        return spla.splu(M).solve(rhs)

    def solve(self, tol=1e-7, maxit=50):
        bc_dofs = np.nonzero(self.markers == 30)[0]                                      # left edges (:133)
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
        """(:175-177) e_div(s), e_0(p), e_1(q) with the weighted norms, degree-6 rule."""
        E, C = self.E, self.C
        x, y = self.P[:, :, 0], self.P[:, :, 1]
        W = (4 * np.pi * x * y) ** 2
        a = self.wts[None, :] * (2 * self.area)[:, None] * W
        R, div = self._rt()
        se = u[self.cell_edges]
        s = np.einsum("cqid,ci->cqd", R, se)
        dr = div[:, None, :] + 2 * R[..., 0] / x[:, :, None] + 2 * R[..., 1] / y[:, :, None]
        drs = np.einsum("cqi,ci->cq", dr, se)
        sx, sy = self.sym["sx"](x, y), self.sym["sy"](x, y)
        es = np.sqrt((a * ((sx - s[..., 0]) ** 2 + (sy - s[..., 1]) ** 2 + (self.sym["drs"](x, y) - drs) ** 2)).sum())
        ep = np.sqrt((a * (p_ex(x, y) - u[E:E + C][:, None]) ** 2).sum())
        qe = u[E + C:][self.mesh.cells]
        qh = qe @ self.phi.T
        qhx = np.einsum("ci,ci->c", qe, self.grads[:, :, 0])[:, None]
        eq = np.sqrt((a * ((q_ex(x, y) - qh) ** 2 + (self.sym["qx"](x, y) - qhx) ** 2)).sum())
        return float(es), float(ep), float(eq)


def table(levels=7):
    rows, prev = [], None
    for nk in range(levels):
        n = 2 ** (nk + 1)
        m = PaperM0(n)
        u = m.solve()
        h = m.mesh.hmax()
        e = m.errors(u)
        r = [0.0, 0.0, 0.0] if prev is None else [np.log(a / b) / np.log(h / prev[0]) for a, b in zip(e, prev[1])]
        rows.append('{:6d} & {:.4f} & {:6.2e} & {:.3f} & {:6.2e} & {:.3f} & {:6.2e} & {:.3f} '.format(
            m.N, h, e[0], r[0], e[1], r[1], e[2], r[2]))
        prev = (h, e)
    return rows
