"""TEST INFRASTRUCTURE (oracle) -- never imported by the product path.

CPU restatement (numpy / scipy) of the two scalar CG2 pure-advection scripts of the reference, written directly as
scalar problems (independent of oracle/paper_m1.py's 17 x 17 element tensors, whose q-q block the GPU path reuses):

* convergence/SUPG_pure_advection_convergence.py:108-112 (kind 'supg'):
      resid (w + stab hK/(2 |r2vec|) r2vec.grad w) W dx,  resid = q + r2vec.grad q + r2^2 q,
      f3 = q_ex + dq_ex/dx + r2^2 q_ex (:97), stab = 0.1 (:55), hK = CellDiameter (:82), q = q_ex on the inlet x = 0
      (:101);
* convergence/C0IP_pure_advection_spherical_convergence.py:107-117 (kind 'c0ip'):
      (q - div_rad(r2vec) q) w W dx - (r2vec.grad w) q W dx + (r2vec.n) q w W ds(inlet, char)
      + stab/(deg+1)^3.5 avg(he)^2 (jump(grad q).n+)(jump(grad w).n+) W dS,
      f3 = q_ex + dq_ex/dx (:98), stab = 0.005 (:53), he = CellDiameter (:80), q = q_ex on the outlet x = 1 (:102).

Both: RectangleMesh n x n on the unit square, n = 2^(nk+1) (:69-70), CG2 (:90), W = (4 pi r1 r2)^2, r2vec = (1,0),
quadrature degree 4 everywhere (:25), error e_1(q) = weighted full H1 norm (:120-122), table rows
'{:6d} & {:.4f} & {:6.2e} & {:.3f} ' (:137).  q_ex is evaluated analytically instead of through its degree-5
interpolant.  parity unpinned: the reference records no output of either script; pinned only through the shared
building blocks (CG2 basis, facet geometry, rules) that reproduce the recorded k = 1 table in oracle/paper_m1.py.
Dof numbering: vertices, then edges (DoF count V + E as in the scripts' first column)."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from . import fem
from .paper_m1 import p2_basis

STAB = {"supg": 0.1, "c0ip": 0.005}
INLET, OUTLET, CHAR = 31, 32, 33


def q_ex(x, y):
    return 1.5 + np.cos(np.pi * x * y)


def q_ex_x(x, y):
    return -np.pi * y * np.sin(np.pi * x * y)


def q_ex_y(x, y):
    return -np.pi * x * np.sin(np.pi * x * y)


class PureAdvection:
    def __init__(self, n, kind, stab=None):
        assert kind in ("supg", "c0ip")
        self.kind, self.stab = kind, STAB[kind] if stab is None else stab
        self.mesh = mesh = fem.unit_square_mesh(n, n)
        self.fc = fc = mesh.facets()
        self.V, self.E, self.C = mesh.num_vertices, fc.verts.shape[0], mesh.num_cells
        self.N = self.V + self.E
        cells = mesh.cells.astype(np.int64)
        fa = np.stack([cells[:, 1], cells[:, 0], cells[:, 0]], 1)
        fb = np.stack([cells[:, 2], cells[:, 2], cells[:, 1]], 1)
        key = np.minimum(fa, fb) * self.V + np.maximum(fa, fb)
        ce = np.searchsorted(fc.verts[:, 0].astype(np.int64) * self.V + fc.verts[:, 1], key)
        self.qd = np.concatenate([cells, self.V + ce], 1)                              # [C,6]
        det, self.grads, self.Xc = fem.cell_geometry(mesh)
        self.area = 0.5 * np.abs(det)
        pts, self.wts = fem.triangle_rule(4)
        self.lam = np.stack([1 - pts[:, 0] - pts[:, 1], pts[:, 0], pts[:, 1]], 1)
        self.P = np.einsum("qi,cid->cqd", self.lam, self.Xc)
        self.markers = fem.mark_facets(mesh, [(lambda P: fem.near(P[..., 0], 0.0), INLET),
                                              (lambda P: fem.near(P[..., 0], 1.0), OUTLET),
                                              (lambda P: fem.near(P[..., 1], 0.0) | fem.near(P[..., 1], 1.0), CHAR)])

    def _facet_lam(self, fidx, cell, s):
        cv = self.mesh.cells[cell]
        va = self.fc.verts[fidx, 0][:, None] == cv
        vb = self.fc.verts[fidx, 1][:, None] == cv
        return va[:, None, :] * (1.0 - s)[None, :, None] + vb[:, None, :] * s[None, :, None]

    def assemble(self):
        """(A, b) of lhs == rhs before the boundary condition."""
        x, y = self.P[..., 0], self.P[..., 1]
        W = (4 * np.pi * x * y) ** 2
        a = self.wts[None, :] * (2 * self.area)[:, None] * W
        phi, gphi = p2_basis(np.broadcast_to(self.lam[None], (self.C,) + self.lam.shape), self.grads)
        phx = gphi[..., 0]
        if self.kind == "supg":
            tw = phi + (0.5 * self.stab * fem.cell_diameter(self.mesh))[:, None, None] * phx   # |r2vec| = 1
            f3 = q_ex(x, y) + q_ex_x(x, y) + x * x * q_ex(x, y)
            Ae = np.einsum("cq,cqk,cql->ckl", a, tw, (1.0 + x * x)[..., None] * phi + phx)
            be = np.einsum("cq,cqk->ck", a * f3, tw)
        else:
            f3 = q_ex(x, y) + q_ex_x(x, y)
            Ae = (np.einsum("cq,cqk,cql->ckl", a * (1.0 - 2.0 / x), phi, phi)          # div_rad((1,0)) = 2/r2
                  - np.einsum("cq,cqk,cql->ckl", a, phx, phi))
            be = np.einsum("cq,cqk->ck", a * f3, phi)
        rows = [np.repeat(self.qd, 6, axis=1).ravel()]
        cols = [np.tile(self.qd, (1, 6)).ravel()]
        vals = [Ae.ravel()]
        b = np.zeros(self.N)
        np.add.at(b, self.qd.ravel(), be.ravel())
        if self.kind == "c0ip":
            s, w = fem.facet_rule(4)
            fc, mesh = self.fc, self.mesh
            # (r2vec.n) q w W ds on the inlet and the characteristic sides
            ext = fc.exterior
            ext = ext[np.isin(self.markers[ext], (INLET, CHAR))]
            L, nrm, A0, B0 = fem.facet_geometry(mesh, ext)
            Pq = A0[:, None, :] + s[None, :, None] * (B0 - A0)[:, None, :]
            Wf = (4 * np.pi * Pq[..., 0] * Pq[..., 1]) ** 2
            cell = fc.cell0[ext]
            ph, _ = p2_basis(self._facet_lam(ext, cell, s), self.grads[cell])
            Jx = np.einsum("nq,nqk,nql->nkl", w[None, :] * (L * nrm[:, 0])[:, None] * Wf, ph, ph)
            rows.append(np.repeat(self.qd[cell], 6, axis=1).ravel())
            cols.append(np.tile(self.qd[cell], (1, 6)).ravel())
            vals.append(Jx.ravel())
            # jump penalty on the interior facets
            fi = fc.interior
            L, nrm, A0, B0 = fem.facet_geometry(mesh, fi)
            Pq = A0[:, None, :] + s[None, :, None] * (B0 - A0)[:, None, :]
            Wf = (4 * np.pi * Pq[..., 0] * Pq[..., 1]) ** 2
            c0, c1 = fc.cell0[fi], fc.cell1[fi]
            d = []
            for cell, sg in ((c0, 1.0), (c1, -1.0)):
                _, g = p2_basis(self._facet_lam(fi, cell, s), self.grads[cell])
                d.append(sg * np.einsum("nqkd,nd->nqk", g, nrm))
            d = np.concatenate(d, 2)
            md = np.concatenate([self.qd[c0], self.qd[c1]], 1)
            hK = fem.cell_diameter(mesh)
            coef = (self.stab / 2.0 ** 3.5 * (0.5 * (hK[c0] + hK[c1])) ** 2 * L)[:, None] * w[None, :] * Wf
            rows.append(np.repeat(md, 12, axis=1).ravel())
            cols.append(np.tile(md, (1, 12)).ravel())
            vals.append(np.einsum("nq,nqk,nql->nkl", coef, d, d).ravel())
        A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))),
                          shape=(self.N, self.N)).tocsr()
        return A, b

    def bcs(self):
        fc, mesh = self.fc, self.mesh
        sel = np.nonzero((fc.cell1 < 0) & (self.markers == (INLET if self.kind == "supg" else OUTLET)))[0]
        va, vb = fc.verts[sel, 0], fc.verts[sel, 1]
        XA, XB = mesh.coords[va], mesh.coords[vb]
        XM = 0.5 * (XA + XB)
        dofs = np.concatenate([va, vb, self.V + sel])
        vals = np.concatenate([q_ex(XA[:, 0], XA[:, 1]), q_ex(XB[:, 0], XB[:, 1]), q_ex(XM[:, 0], XM[:, 1])])
        d, i = np.unique(dofs, return_index=True)
        return d, vals[i]

    def solve(self):
        A, b = self.assemble()
        d, g = self.bcs()
        A = A.tolil()
        for r, v in zip(d, g):            # DirichletBC.apply: identity row, data on the right-hand side
            A.rows[r], A.data[r] = [int(r)], [1.0]
            b[r] = v
        return spla.splu(A.tocsc()).solve(b)

    def error(self, q):
        x, y = self.P[..., 0], self.P[..., 1]
        a = self.wts[None, :] * (2 * self.area)[:, None] * (4 * np.pi * x * y) ** 2
        phi, gphi = p2_basis(np.broadcast_to(self.lam[None], (self.C,) + self.lam.shape), self.grads)
        qh = np.einsum("cqk,ck->cq", phi, q[self.qd])
        gx = np.einsum("cqk,ck->cq", gphi[..., 0], q[self.qd])
        gy = np.einsum("cqk,ck->cq", gphi[..., 1], q[self.qd])
        return float(np.sqrt((a * ((q_ex(x, y) - qh) ** 2 + (q_ex_x(x, y) - gx) ** 2
                                   + (q_ex_y(x, y) - gy) ** 2)).sum()))


def table(kind, levels=6):
    rows, prev = [], None
    for nk in range(levels):
        m = PureAdvection(2 ** (nk + 1), kind)
        e, h = m.error(m.solve()), m.mesh.hmax()
        r = 0.0 if prev is None else np.log(e / prev[1]) / np.log(h / prev[0])
        rows.append('{:6d} & {:.4f} & {:6.2e} & {:.3f} '.format(m.N, h, e, r))
        prev = (h, e)
    return rows
