"""Lowest-order mixed-primal system RT1 x DG0 x CG1 (deg = 0) of
convergence/test_for_paper_convergence.py:68,108-167 on the GPU.

Host side = topology tables (edge numbering, orientation signs, the scalar CSR pattern and the per-slot
gather lists that make the scatter-add deterministic) built once per mesh with torch; per Newton step
the work is in the CUDA library: cm_mixed_m0_cells (element tensors), cm_csr_gather (assembly),
cm_assemble_c0ip_facets / cm_assemble_exterior_facets (q-q facet operator, shared with the P-Q path),
cm_apply_dirichlet, and the pivoting banded LU cm_blu_* ('mumps' at :162).  Global dofs
[edges | cells | vertices] (DOLFIN's own numbering of this mixed space is not reproducible without
DOLFIN; the results are numbering independent).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _capi
from .assembler import PQAssembler
from .forms import EF_ADV, EF_DATA, EF_PENALTY, PQPrimalForm
from .manufactured import _torchify, p_exact, q_exact
from .mesh import DOLFIN_EPS
from .quadrature import collapsed_triangle_rule, facet_rule, triangle_rule
from .solver import BandedLULinearSolver
from .space import FiniteElement, FunctionSpace, MixedElement


def paper_data(D1, D2, gcoef=1.0, q_react=0.0, x2_on_q=False, x2_coef=1.0, r2dir=1.0):
    """sig_ex = -grad(p_ex) - G(p_ex,q_ex), f2_ex = div_rad(D*sig_ex), f3_ex = q_react q_ex + dq_ex/dx + r2^2 p_ex
    (test_for_paper_convergence.py:124-128; ...MixedPrimal_convergence.py:109-113 with gcoef = D2 4 pi, q_react = 1),
    div_rad(sig_ex) and grad q_ex for the error norms, as torch callables."""
    import sympy as sm
    x, y = sm.symbols("x y", positive=True)
    pe = sm.sin(sm.pi * x) * sm.sin(sm.pi * y)
    qe = sm.Rational(3, 2) + sm.cos(sm.pi * x * y)
    G = gcoef * x ** 2 * pe ** 2 / (qe + DOLFIN_EPS)
    sx = -sm.diff(pe, x) - G
    sy = -sm.diff(pe, y)
    f2 = sm.diff(x ** 2 * D2 * sx, x) / x ** 2 + sm.diff(y ** 2 * D1 * sy, y) / y ** 2
    # "r2**2*q_ex" in ..._SUPG.py:112; grad(q_ex).r2vec with r2vec = (-1,0) in ..._C0IP_otherBCs.py:117
    f3 = q_react * qe + r2dir * sm.diff(qe, x) + x2_coef * x ** 2 * (qe if x2_on_q else pe)
    drs = sm.diff(x ** 2 * sx, x) / x ** 2 + sm.diff(y ** 2 * sy, y) / y ** 2
    T = lambda e: _torchify(e, (x, y))
    return dict(sx=T(sx), sy=T(sy), f2=T(f2), f3=T(f3), drs=T(drs), qx=T(sm.diff(qe, x)), qy=T(sm.diff(qe, y)))


class _Holder:
    """What BandedLULinearSolver needs from an assembler: mesh.device / .structured, topo, vals, b, bs."""

    class _M:
        structured = False

    class _T:
        pass

    def __init__(self, device, rowptr, col, n, vals, b):
        self.mesh = self._M()
        self.mesh.device = device
        self.topo = self._T()
        self.topo.nrowptr, self.topo.ncol, self.topo.nv = rowptr, col, n
        self.vals, self.b, self.bs = vals, b, 1


def _csr_from_pairs(rows, cols, n):
    key = torch.unique(rows * n + cols)
    r = key // n
    rp = torch.zeros(n + 1, dtype=torch.int64, device=rows.device)
    rp[1:] = torch.cumsum(torch.bincount(r, minlength=n), 0)
    return key, rp.to(torch.int32).contiguous(), (key % n).to(torch.int32).contiguous()


def _gather_lists(slot, nslots):
    """Per-slot lists of the contributing entries, ascending entry index (= ascending cell)."""
    _, order = torch.sort(slot, stable=True)
    ptr = torch.zeros(nslots + 1, dtype=torch.int64, device=slot.device)
    ptr[1:] = torch.cumsum(torch.bincount(slot, minlength=nslots), 0)
    return ptr.to(torch.int32).contiguous(), order


class PaperMixedSystem:
    """Assembly + Newton solve of the k=0 'paper' system on UnitSquareMesh-like meshes with the facet
    markers left=30, right=31, rest=32 (:90-94)."""

    def __init__(self, mesh, bdry, D1=1.0, D2=1.5, stab=0.05, stab2=1.0, degree=6, tables_only=False):
        if not tables_only:
            _capi.require_cuda(mesh.coords)
            self.lib = _capi.lib()
        self.mesh, self.bdry = mesh, bdry
        self.D1, self.D2 = float(D1), float(D2)
        dev = mesh.device
        fc = mesh.facets()
        V, C, E = mesh.num_vertices(), mesh.num_cells(), fc.num
        self.V, self.C, self.E, self.N = V, C, E, E + C + V
        cells = mesh.cells.long()
        X = mesh.coords
        # cell -> edges (local facet i opposite local vertex i) and orientation signs
        fa = torch.stack([cells[:, 1], cells[:, 0], cells[:, 0]], 1)
        fb = torch.stack([cells[:, 2], cells[:, 2], cells[:, 1]], 1)
        key = torch.minimum(fa, fb) * V + torch.maximum(fa, fb)
        allkey = fc.verts[:, 0].long() * V + fc.verts[:, 1].long()
        ce = torch.searchsorted(allkey, key)
        A, B = X[fc.verts[:, 0].long()], X[fc.verts[:, 1].long()]
        t = B - A
        self.elen = t.pow(2).sum(1).sqrt()
        self.enorm = torch.stack([t[:, 1], -t[:, 0]], 1) / self.elen[:, None]
        mid = 0.5 * (A + B)[ce]
        sign = torch.sign(((mid - X[cells]) * self.enorm[ce]).sum(-1))
        self.cell_edges = ce.to(torch.int32).contiguous()
        self.sign = sign.to(torch.float64).contiguous()
        # quadrature rule and forcing at the cell points
        pts, wts = triangle_rule(degree)
        self.rule = np.ascontiguousarray(np.concatenate([pts, wts[:, None]], 1), dtype=np.float64)
        self.nq = self.rule.shape[0]
        phi = torch.tensor(np.stack([1.0 - pts[:, 0] - pts[:, 1], pts[:, 0], pts[:, 1]], 1), device=dev)
        P = torch.einsum("qi,cid->cqd", phi, X[cells])
        self.xq, self.yq = P[:, :, 0].contiguous(), P[:, :, 1].contiguous()
        self.data = paper_data(D1, D2)
        self.f2q = self.data["f2"](self.xq, self.yq).to(torch.float64).contiguous()
        self.f3q = self.data["f3"](self.xq, self.yq).to(torch.float64).contiguous()
        # q-q facet operator: the C0IP 'paper' q-equation on the CG1 x CG1 machinery (same kernels)
        self.qtopo = mesh.topology(True)
        if not tables_only:
            P1 = FiniteElement("CG", "triangle", 1)
            W = FunctionSpace(mesh, MixedElement([P1, P1]))
            form = PQPrimalForm(W, D1, D2, G=("power", 1.0), q_adv=0.0, q_divr=1.0, q_ibp=1.0,
                                c0ip_gamma=stab / 1.0 ** 3.5, c0ip_h="facet", pen_stab2=stab2, degree=degree,
                                ext_terms=[(bdry, 30, EF_ADV), (bdry, 32, EF_ADV), (bdry, 31, EF_DATA),
                                           (bdry, "on_boundary", EF_PENALTY)], q_data=q_exact)
            qasm = PQAssembler(form)
            assert qasm.topo is self.qtopo
            self.Kq, self.cq = qasm.Kq, qasm.cq
        # scalar CSR pattern: cell blocks + the (vertex x vertex) facet pattern
        cid = torch.arange(C, device=dev)[:, None]
        cd = torch.cat([ce, E + cid, E + C + cells], 1)                               # [C,7]
        self.cell_dofs = cd
        rows_c = cd.repeat_interleave(7, dim=1).flatten()
        cols_c = cd.repeat(1, 7).flatten()
        qdeg = (self.qtopo.nrowptr[1:] - self.qtopo.nrowptr[:-1]).long()
        qrows = E + C + torch.repeat_interleave(torch.arange(V, device=dev), qdeg)
        qcols = E + C + self.qtopo.ncol.long()
        N = self.N
        self._key, self.rowptr, self.col = _csr_from_pairs(torch.cat([rows_c, qrows]), torch.cat([cols_c, qcols]), N)
        self.nnz = int(self.col.numel())
        slot_c = torch.searchsorted(self._key, rows_c * N + cols_c)                     # cell-major, (i,j) inside
        # element storage is entry-major ([49][C]): entry index of (c, e) is e*C + c
        ent = (torch.arange(49, device=dev)[None, :] * C + torch.arange(C, device=dev)[:, None]).flatten()
        self.mat_ptr, order = _gather_lists(slot_c, self.nnz)
        self.mat_idx = ent[order].to(torch.int32).contiguous()
        entv = (torch.arange(7, device=dev)[None, :] * C + torch.arange(C, device=dev)[:, None]).flatten()
        self.vec_ptr, order = _gather_lists(cd.flatten(), N)
        self.vec_idx = entv[order].to(torch.int32).contiguous()
        self.kq_map = torch.searchsorted(self._key, qrows * N + qcols).to(torch.int32).contiguous()
        self.diag_slot_all = (torch.searchsorted(self._key, torch.arange(N, device=dev) * (N + 1))
                              - self.rowptr[:-1].long()).to(torch.int32)
        if tables_only:
            return
        # buffers
        f64 = dict(dtype=torch.float64, device=dev)
        self.Je = torch.empty(49 * C, **f64)
        self.Fe = torch.empty(7 * C, **f64)
        self.vals = torch.zeros(self.nnz, **f64)
        self.b = torch.zeros(N, **f64)
        self._kqq = torch.empty(V, **f64)
        self.bnat = self._natural_data(degree)
        self._lin = None
        self.newton_its = 0

    # -- state-independent pieces ----------------------------------------------------------------
    def _natural_data(self, degree):
        """+ p_ex (tau.n) W ds(right), ds(rest) on the s-rows (:148-149): RT0 normal trace = 1/|F| on its
        own facet, so only the facet's own dof is touched."""
        mesh, fc = self.mesh, self.mesh.facets()
        dev = mesh.device
        sel = torch.nonzero((fc.cell1 < 0) & ((self.bdry.values == 31) | (self.bdry.values == 32))).flatten()
        A = mesh.coords[fc.verts[sel, 0].long()]
        B = mesh.coords[fc.verts[sel, 1].long()]
        t = B - A
        L = t.pow(2).sum(1).sqrt()
        nrm = torch.stack([t[:, 1], -t[:, 0]], 1) / L[:, None]
        opp = mesh.coords[mesh.cells.long()[fc.cell0[sel].long(), fc.loc0[sel].long()]]
        flip = ((opp - A) * nrm).sum(1) > 0
        nrm = torch.where(flip[:, None], -nrm, nrm)
        m = (degree + 2) // 2
        p, w = np.polynomial.legendre.leggauss(m)
        s = torch.tensor(0.5 * (p + 1.0), device=dev)
        wq = torch.tensor(0.5 * w, device=dev)
        Pq = A[:, None, :] + s[None, :, None] * t[:, None, :]
        Wf = (4.0 * math.pi * Pq[:, :, 0] * Pq[:, :, 1]) ** 2
        sgn = torch.sign((nrm * self.enorm[sel]).sum(1))
        out = torch.zeros(self.N, dtype=torch.float64, device=dev)
        out[sel] = (wq[None, :] * Wf * p_exact(Pq[:, :, 0], Pq[:, :, 1])).sum(1) * sgn
        return out

    def sigD(self):
        """project(sig_ex, RT) (:132): unweighted L2 projection; RT mass matrix assembled with the same
        gather machinery and solved with the banded LU."""
        dev = self.mesh.device
        C, E = self.C, self.E
        cells = self.mesh.cells.long()
        Xc = self.mesh.coords[cells]
        det = ((Xc[:, 1, 0] - Xc[:, 0, 0]) * (Xc[:, 2, 1] - Xc[:, 0, 1])
               - (Xc[:, 2, 0] - Xc[:, 0, 0]) * (Xc[:, 1, 1] - Xc[:, 0, 1]))
        area = 0.5 * det.abs()
        wts = torch.tensor(self.rule[:, 2], device=dev)
        a = wts[None, :] * (2 * area)[:, None]
        P = torch.stack([self.xq, self.yq], -1)
        R = (P[:, :, None, :] - Xc[:, None, :, :]) * (self.sign / (2 * area[:, None]))[:, None, :, None]
        sex = torch.stack([self.data["sx"](self.xq, self.yq), self.data["sy"](self.xq, self.yq)], -1)
        Me = torch.einsum("cq,cqid,cqjd->cij", a, R, R)                                # [C,3,3]
        be = torch.einsum("cq,cqid,cqd->ci", a, R, sex)
        ce = self.cell_edges.long()
        rows = ce.repeat_interleave(3, dim=1).flatten()
        cols = ce.repeat(1, 3).flatten()
        key, rp, col = _csr_from_pairs(rows, cols, E)
        slot = torch.searchsorted(key, rows * E + cols)
        ptr, order = _gather_lists(slot, key.numel())
        vals = torch.empty(key.numel(), dtype=torch.float64, device=dev)
        st = _capi.stream_ptr()
        src = Me.flatten().contiguous()
        _capi.check(self.lib.cm_csr_gather(key.numel(), _capi.ptr(ptr), _capi.ptr(order.to(torch.int32).contiguous()),
                                           _capi.ptr(src), _capi.ptr(vals), 0, st))
        ptr2, order2 = _gather_lists(ce.flatten(), E)
        rhs = torch.empty(E, dtype=torch.float64, device=dev)
        src2 = be.flatten().contiguous()
        _capi.check(self.lib.cm_csr_gather(E, _capi.ptr(ptr2), _capi.ptr(order2.to(torch.int32).contiguous()),
                                           _capi.ptr(src2), _capi.ptr(rhs), 0, st))
        lin = BandedLULinearSolver(_Holder(dev, rp, col, E, vals, rhs))
        lin.setup()
        x = torch.empty_like(rhs)
        lin.solve(rhs, x)
        return x

    # -- per Newton step ----------------------------------------------------------------------------
    def assemble(self, u, want_jac=True):
        """F(u) -> self.b and (optionally) J(u) -> self.vals on the scalar CSR (rowptr, col)."""
        lib, st, mesh = self.lib, _capi.stream_ptr(), self.mesh
        _capi.check(lib.cm_mixed_m0_cells(
            self.nq, self.rule.ctypes.data, self.D1, self.D2, DOLFIN_EPS, self.C, _capi.ptr(mesh.cells),
            _capi.ptr(mesh.coords), _capi.ptr(self.cell_edges), _capi.ptr(self.sign), self.E, _capi.ptr(u),
            _capi.ptr(self.f2q), _capi.ptr(self.f3q), _capi.ptr(self.Je) if want_jac else None,
            _capi.ptr(self.Fe), st))
        _capi.check(lib.cm_csr_gather(self.N, _capi.ptr(self.vec_ptr), _capi.ptr(self.vec_idx),
                                      _capi.ptr(self.Fe), _capi.ptr(self.b), 0, st))
        # q-rows: + Kq q + cq (facet terms), s-rows: + natural p data
        uq = u[self.E + self.C:]
        t = self.qtopo
        _capi.check(lib.cm_spmv(1, self.V, _capi.ptr(t.nrowptr), _capi.ptr(t.ncol), _capi.ptr(self.Kq),
                                uq.data_ptr(), _capi.ptr(self._kqq), st))
        self.b[self.E + self.C:] += self._kqq + self.cq
        self.b += self.bnat
        if want_jac:
            _capi.check(lib.cm_csr_gather(self.nnz, _capi.ptr(self.mat_ptr), _capi.ptr(self.mat_idx),
                                          _capi.ptr(self.Je), _capi.ptr(self.vals), 0, st))
            _capi.check(lib.cm_csr_add_mapped(self.kq_map.numel(), _capi.ptr(self.kq_map), _capi.ptr(self.Kq),
                                              _capi.ptr(self.vals), st))
        return (self.vals if want_jac else None), self.b

    def _apply_bc(self, u, with_matrix):
        _capi.check(self.lib.cm_apply_dirichlet(
            1, self.bc_dofs.numel(), _capi.ptr(self.bc_dofs), _capi.ptr(self.bc_diag), _capi.ptr(self.bc_vals),
            _capi.ptr(u), _capi.ptr(self.rowptr), _capi.ptr(self.vals) if with_matrix else None,
            _capi.ptr(self.b), _capi.stream_ptr()))

    def solve(self, tol=1e-7, maxit=50):
        """SNES 'basic' Newton from u = 0 with 'mumps' (:159-165): abs / rel tolerance 1e-7."""
        dev = self.mesh.device
        fc = self.mesh.facets()
        bc = torch.nonzero((fc.cell1 < 0) & (self.bdry.values == 30)).flatten()          # left edges (:133)
        self.bc_dofs = bc.to(torch.int32).contiguous()
        self.bc_vals = self.sigD()[bc].contiguous()
        self.bc_diag = self.diag_slot_all[bc].contiguous()
        u = torch.zeros(self.N, dtype=torch.float64, device=dev)
        self.assemble(u, False)
        self._apply_bc(u, False)
        r0 = r = float(torch.linalg.vector_norm(self.b))
        if self._lin is None:
            self._lin = BandedLULinearSolver(_Holder(dev, self.rowptr, self.col, self.N, self.vals, self.b))
        dx = torch.empty_like(u)
        it = 0
        while not (r < tol or r / r0 < tol) and it < maxit:
            self.assemble(u, True)
            self._apply_bc(u, True)
            self._lin.setup()
            self._lin.solve(self.b, dx)
            u -= dx
            it += 1
            self.assemble(u, False)
            self._apply_bc(u, False)
            r = float(torch.linalg.vector_norm(self.b))
            if r != r:
                raise RuntimeError("Newton solver diverged (NaN residual)")
        if not (r < tol or r / r0 < tol):
            raise RuntimeError("Newton solver did not converge because maximum number of iterations reached")
        self.newton_its = it
        return u

    def errors(self, u):
        """(e_div(s), e_0(p), e_1(q)) of :175-177 (weighted norms, degree-6 rule)."""
        d, x, y = self.data, self.xq, self.yq
        exq = torch.stack([d["sx"](x, y), d["sy"](x, y), d["drs"](x, y), p_exact(x, y), q_exact(x, y),
                           d["qx"](x, y)], -1).to(torch.float64).contiguous()
        work = torch.empty(int(self.lib.cm_mixed_workspace_doubles()), dtype=torch.float64,
                           device=self.mesh.device)
        _capi.check(self.lib.cm_mixed_m0_errors(
            self.nq, self.rule.ctypes.data, self.C, _capi.ptr(self.mesh.cells), _capi.ptr(self.mesh.coords),
            _capi.ptr(self.cell_edges), _capi.ptr(self.sign), self.E, _capi.ptr(u), _capi.ptr(exq),
            _capi.ptr(work), _capi.stream_ptr()))
        s = work[:3].cpu()
        return tuple(math.sqrt(float(v)) for v in s)


# ---------------------------------------------------------------------------------------------------------
# k = 1: RT2 x DG1 x CG2 (test_for_paper_convergence.py with deg = 1, recorded table :211-217)
# ---------------------------------------------------------------------------------------------------------
_OTHER = ((1, 2), (0, 2), (0, 1))      # end points (local vertices) of the facet opposite local vertex i


def _p2_basis(lam, g):
    """lam [n, nq, 3], g [n, 3, 2] -> CG2 values [n, nq, 6] and gradients [n, nq, 6, 2] (vertices, then facets)."""
    val = torch.cat([lam * (2 * lam - 1),
                     torch.stack([4 * lam[..., j] * lam[..., k] for j, k in _OTHER], -1)], -1)
    gv = (4 * lam - 1)[..., None] * g[:, None, :, :]
    ge = torch.stack([4 * (lam[..., j, None] * g[:, None, k, :] + lam[..., k, None] * g[:, None, j, :])
                      for j, k in _OTHER], -2)
    return val, torch.cat([gv, ge], -2)


class PaperMixedSystemK1:
    """RT2 x DG1 x CG2 'paper' system on the GPU: per Newton step cm_mixed_m1_cells (17 x 17 element tensors),
    cm_csr_gather, cm_spmv, cm_apply_dirichlet and the pivoting LU (cm_blu_*).  The q-q facet operator (C0IP
    jump penalty on CG2, weak outflow data, boundary penalty; :142-145,152-153) and the natural p data on the
    flux rows (:148-149) are state independent: cm_cg2_exterior_facets / cm_cg2_interior_facets tabulate them
    once per mesh and they enter every step as a constant matrix / vector (_facet_tables is the torch
    restatement of the same operator for the CPU table test)."""

    def __init__(self, mesh, bdry, D1=1.0, D2=1.5, stab=0.1, stab2=2.0, degree=6, tables_only=False,
                 variant="paper", supg_stab=1.0):
        """variant 'paper': test_for_paper_convergence.py (deg = 1; markers left 30 / right 31 / rest 32);
        variant 'galerkin': ...MixedPrimal_convergence.py (BASELINE config 2's own script: G = D2 4 pi r2^2 p^2/q,
        Galerkin q-equation with reaction, natural p data on the whole boundary, q = q_ex on the facets marked 31,
        no facet stabilisation; pass D1 = 1e-3, D2 = 1.5e-3, degree = 4); variant 'supg': ..._SUPG.py, the same with
        the SUPG test function and the script's residual q + dq/dx + r2^2 q."""
        if not tables_only:
            _capi.require_cuda(mesh.coords)
            self.lib = _capi.lib()
        self.mesh, self.bdry = mesh, bdry
        self.D1, self.D2 = float(D1), float(D2)
        if variant not in ("paper", "galerkin", "supg", "c0ip", "c0ip_other", "pure_supg", "pure_c0ip"):
            raise ValueError(f"unknown variant '{variant}'")
        # facet-term layout of the integrated-by-parts q-equation:
        #   'paper' test_for_paper_convergence.py:142-153 (left 30, right 31, rest 32);
        #   'c0ip'  ...MixedPrimal_convergence_C0IP.py:144-149 (left 31, right 32, rest 33): inflow term on the left only,
        #           no weak data / penalty, natural p data on the whole boundary, avg(hK) with hK = CellDiameter, the
        #           caller passes the signed stab (the script uses -0.00075) and stab2 = 0
        #   'pure_c0ip' C0IP_pure_advection_spherical_convergence.py:109-113 (inlet 31, outlet 32, char 33): the q-q block
        #           only (PureAdvectionCG2): inflow term on inlet and char, reaction q, no p coupling, penalty +stab
        #   'c0ip_other' ..._C0IP_otherBCs.py:131-137 (left 30, right 31, rest 32): r2vec = (-1,0) in the q-equation,
        #           inflow term on the left, natural p data on right and rest, s = project(sig_ex) on the left, q = q_ex
        #           on the right, penalty +stab, hK = CellDiameter
        self.adv_markers, self.dat_markers, self.nat_markers, self.h_cell = (
            ((31,), (), None, True) if variant == "c0ip" else ((31, 33), (), None, True) if variant == "pure_c0ip"
            else ((30,), (), (31, 32), True) if variant == "c0ip_other" else ((30, 32), (31,), (31, 32), False))
        self.q_bc_marker, self.s_bc_marker = ((32, 31) if variant in ("c0ip", "pure_c0ip")
                                              else (31, 30) if variant == "c0ip_other" else (None, 30))
        self.r2dir = -1.0 if variant == "c0ip_other" else 1.0
        self.variant = variant
        # cell-only q-equation, natural p data on all of the boundary
        self.galerkin = variant in ("galerkin", "supg", "pure_supg")
        self.gcoef = self.D2 * 4.0 * math.pi if self.galerkin else 1.0
        self.qmode, self.q_react = ((1, 1.0) if self.galerkin else (2, 0.0) if variant == "c0ip_other"
                                    else (0, 1.0 if variant == "pure_c0ip" else 0.0))
        # 'supg': ...MixedPrimal_convergence_SUPG.py:118-129 (residual q + dq/dx + r2^2 q, test w + stab hK/2 dw/dx);
        # 'pure_supg': SUPG_pure_advection_convergence.py:106-109, the same q-q block with stab = 0.1
        self.q_x2p, self.q_x2q, self.supg_stab = ((0.0, 1.0, float(supg_stab)) if variant in ("supg", "pure_supg")
                                                  else (0.0, 0.0, 0.0) if variant == "pure_c0ip" else (1.0, 0.0, 0.0))
        self.gamma = stab / 2.0 ** 3.5                       # stab/(deg+1)^3.5 (:144)
        self.stab2 = float(stab2)
        dev = mesh.device
        fc = mesh.facets()
        V, C, E = mesh.num_vertices(), mesh.num_cells(), fc.num
        self.V, self.C, self.E = V, C, E
        self.N = N = 3 * E + 5 * C + V
        cells = mesh.cells.long()
        X = mesh.coords
        fa = torch.stack([cells[:, 1], cells[:, 0], cells[:, 0]], 1)
        fb = torch.stack([cells[:, 2], cells[:, 2], cells[:, 1]], 1)
        key = torch.minimum(fa, fb) * V + torch.maximum(fa, fb)
        allkey = fc.verts[:, 0].long() * V + fc.verts[:, 1].long()
        ce = torch.searchsorted(allkey, key)
        A, B = X[fc.verts[:, 0].long()], X[fc.verts[:, 1].long()]
        t = B - A
        self.elen = t.pow(2).sum(1).sqrt()
        self.enorm = torch.stack([t[:, 1], -t[:, 0]], 1) / self.elen[:, None]
        Xc = X[cells]
        self.sign = torch.sign(((0.5 * (A + B)[ce] - Xc) * self.enorm[ce]).sum(-1)).to(torch.float64).contiguous()
        self.cell_edges = ce.to(torch.int32).contiguous()
        det = ((Xc[:, 1, 0] - Xc[:, 0, 0]) * (Xc[:, 2, 1] - Xc[:, 0, 1])
               - (Xc[:, 2, 0] - Xc[:, 0, 0]) * (Xc[:, 1, 1] - Xc[:, 0, 1]))
        g1 = torch.stack([Xc[:, 2, 1] - Xc[:, 0, 1], -(Xc[:, 2, 0] - Xc[:, 0, 0])], 1) / det[:, None]
        g2 = torch.stack([-(Xc[:, 1, 1] - Xc[:, 0, 1]), Xc[:, 1, 0] - Xc[:, 0, 0]], 1) / det[:, None]
        self.grads = torch.stack([-(g1 + g2), g1, g2], 1)                              # [C,3,2]
        self.area = 0.5 * det.abs()
        pts, wts = triangle_rule(degree)
        self.rule = np.ascontiguousarray(np.concatenate([pts, wts[:, None]], 1), dtype=np.float64)
        self.nq = self.rule.shape[0]
        self.lamq = torch.tensor(np.stack([1.0 - pts[:, 0] - pts[:, 1], pts[:, 0], pts[:, 1]], 1), device=dev)
        P = torch.einsum("qi,cid->cqd", self.lamq, Xc)
        self.xq, self.yq = P[:, :, 0].contiguous(), P[:, :, 1].contiguous()
        self.data = paper_data(D1, D2, self.gcoef, self.q_react, variant in ("supg", "pure_supg"),
                               0.0 if variant == "pure_c0ip" else 1.0, self.r2dir)
        self.f2q = self.data["f2"](self.xq, self.yq).to(torch.float64).contiguous()
        self.f3q = self.data["f3"](self.xq, self.yq).to(torch.float64).contiguous()
        # global dofs of the 17 local ones: [2 per edge | 2 per cell | 3 per cell | vertices | edges]
        cid = torch.arange(C, device=dev)[:, None]
        sd = torch.cat([torch.stack([2 * ce[:, i], 2 * ce[:, i] + 1], 1) for i in range(3)]
                       + [2 * E + 2 * cid, 2 * E + 2 * cid + 1], 1)
        pd = 2 * E + 2 * C + 3 * cid + torch.arange(3, device=dev)[None, :]
        self.q0 = q0 = 2 * E + 5 * C
        qd = torch.cat([q0 + cells, q0 + V + ce], 1)
        self.sd, self.pd, self.qd = sd, pd, qd
        self.cell_dofs = cd = torch.cat([sd, pd, qd], 1)                               # [C,17]
        # constant facet operator (COO) and vectors: CUDA kernels (cm_cg2_*_facets); the torch tabulation
        # _facet_tables is the host restatement used by the CPU table test and the GPU cross-check
        if self.galerkin:      # no facet operator: only the natural p data on every exterior facet (:126)
            fr = fcn = torch.zeros(0, dtype=torch.int64, device=dev)
            fv = torch.zeros(0, dtype=torch.float64, device=dev)
            self.cvec = self._natural_all(degree)
            cv_rows = cv_vals = None
        elif tables_only:
            fr, fcn, fv, self.cvec = self._facet_tables(degree)
            cv_rows = cv_vals = None
        else:
            fr, fcn, fv, cv_rows, cv_vals = self._facet_tables_cuda(degree)
        rows_c = cd.repeat_interleave(17, dim=1).flatten()
        cols_c = cd.repeat(1, 17).flatten()
        self._key, self.rowptr, self.col = _csr_from_pairs(torch.cat([rows_c, fr]), torch.cat([cols_c, fcn]), N)
        self.nnz = int(self.col.numel())
        slot_c = torch.searchsorted(self._key, rows_c * N + cols_c)
        ent = (torch.arange(289, device=dev)[None, :] * C + torch.arange(C, device=dev)[:, None]).flatten()
        self.mat_ptr, order = _gather_lists(slot_c, self.nnz)
        self.mat_idx = ent[order].to(torch.int32).contiguous()
        entv = (torch.arange(17, device=dev)[None, :] * C + torch.arange(C, device=dev)[:, None]).flatten()
        self.vec_ptr, order = _gather_lists(cd.flatten(), N)
        self.vec_idx = entv[order].to(torch.int32).contiguous()
        fslot = torch.searchsorted(self._key, fr * N + fcn)
        if tables_only or self.galerkin:
            self.Kf = torch.zeros(self.nnz, dtype=torch.float64, device=dev)
            self.Kf.index_add_(0, fslot, fv)
        else:   # deterministic: per-slot gather lists in facet order (exterior, then interior)
            st = _capi.stream_ptr()
            ptr, order = _gather_lists(fslot, self.nnz)
            self.Kf = torch.empty(self.nnz, dtype=torch.float64, device=dev)
            fv = fv.contiguous()
            _capi.check(self.lib.cm_csr_gather(self.nnz, _capi.ptr(ptr), _capi.ptr(order.to(torch.int32).contiguous()),
                                               _capi.ptr(fv), _capi.ptr(self.Kf), 0, st))
            ptr, order = _gather_lists(cv_rows, N)
            self.cvec = torch.empty(N, dtype=torch.float64, device=dev)
            cv_vals = cv_vals.contiguous()
            _capi.check(self.lib.cm_csr_gather(N, _capi.ptr(ptr), _capi.ptr(order.to(torch.int32).contiguous()),
                                               _capi.ptr(cv_vals), _capi.ptr(self.cvec), 0, st))
            torch.cuda.current_stream().synchronize()
        self.diag_slot_all = (torch.searchsorted(self._key, torch.arange(N, device=dev) * (N + 1))
                              - self.rowptr[:-1].long()).to(torch.int32)
        if tables_only:
            return
        f64 = dict(dtype=torch.float64, device=dev)
        self.Je = torch.empty(289 * C, **f64)
        self.Fe = torch.empty(17 * C, **f64)
        self.vals = torch.zeros(self.nnz, **f64)
        self.b = torch.zeros(N, **f64)
        self._ku = torch.empty(N, **f64)
        self._lin = None
        self.newton_its = 0

    # -- state-independent facet terms -----------------------------------------------------------------------
    def _facet_lam(self, fidx, cell, s):
        fc, cv = self.mesh.facets(), self.mesh.cells.long()[cell]
        va = (fc.verts[fidx, 0].long()[:, None] == cv).to(torch.float64)
        vb = (fc.verts[fidx, 1].long()[:, None] == cv).to(torch.float64)
        return va[:, None, :] * (1.0 - s)[None, :, None] + vb[:, None, :] * s[None, :, None]

    def _facet_geometry(self, fidx):
        mesh, fc = self.mesh, self.mesh.facets()
        A = mesh.coords[fc.verts[fidx, 0].long()]
        B = mesh.coords[fc.verts[fidx, 1].long()]
        t = B - A
        L = t.pow(2).sum(1).sqrt()
        nrm = torch.stack([t[:, 1], -t[:, 0]], 1) / L[:, None]
        opp = mesh.coords[mesh.cells.long()[fc.cell0[fidx].long(), fc.loc0[fidx].long()]]
        flip = ((opp - A) * nrm).sum(1) > 0
        return L, torch.where(flip[:, None], -nrm, nrm), A, B

    def _facet_tables(self, degree):
        """COO (rows, cols, vals) of the constant q-q facet matrix and the constant vector (q_ex data on the
        q-rows, natural p data on the flux rows)."""
        mesh, fc, dev = self.mesh, self.mesh.facets(), self.mesh.device
        m = (degree + 2) // 2
        p, w = np.polynomial.legendre.leggauss(m)
        s = torch.tensor(0.5 * (p + 1.0), device=dev)
        w = torch.tensor(0.5 * w, device=dev)
        cvec = torch.zeros(self.N, dtype=torch.float64, device=dev)
        rows, cols, vals = [], [], []
        # exterior facets
        ext = torch.nonzero(fc.cell1 < 0).flatten()
        L, nrm, A, B = self._facet_geometry(ext)
        cell = fc.cell0[ext].long()
        lam = self._facet_lam(ext, cell, s)
        Pq = A[:, None, :] + s[None, :, None] * (B - A)[:, None, :]
        x, y = Pq[..., 0], Pq[..., 1]
        W = (4.0 * math.pi * x * y) ** 2
        phi, _ = _p2_basis(lam, self.grads[cell])
        mk = self.bdry.values[ext]
        isin = lambda ms: (torch.zeros_like(mk, dtype=torch.bool) if not ms
                           else torch.stack([mk == m_ for m_ in ms]).any(0))
        k_adv = isin(self.adv_markers).to(torch.float64)
        k_dat = isin(self.dat_markers).to(torch.float64)
        bn = self.r2dir * nrm[:, 0]
        neg = 0.5 * (bn.abs() - bn)
        pen = self.stab2 * (L * neg ** 2 + 8.0 / L)
        ds = w[None, :] * L[:, None]
        cq = (k_adv * bn)[:, None] * W + pen[:, None]
        cdat = (k_dat * bn)[:, None] * W - pen[:, None]
        qd = self.qd[cell]
        cvec.index_add_(0, qd.flatten(), torch.einsum("nq,nqk->nk", ds * cdat * q_exact(x, y), phi).flatten())
        Jx = torch.einsum("nq,nqk,nql->nkl", ds * cq, phi, phi)
        rows.append(qd.repeat_interleave(6, dim=1).flatten())
        cols.append(qd.repeat(1, 6).flatten())
        vals.append(Jx.flatten())
        sel = torch.ones_like(mk, dtype=torch.bool) if self.nat_markers is None else isin(self.nat_markers)
        sgn = torch.sign((nrm * self.enorm[ext]).sum(1))
        pw = ds * W * p_exact(x, y)
        for slot, lamend in ((0, 1.0 - s), (1, s)):
            val = (pw * lamend[None, :]).sum(1) * sgn / L
            cvec.index_add_(0, (2 * ext + slot)[sel], val[sel])
        # interior facets: gamma avg(hK)^2 (jump(grad q).n+)(jump(grad w).n+) W dS, hK = FacetArea
        fi = torch.nonzero(fc.cell1 >= 0).flatten()
        L, nrm, A, B = self._facet_geometry(fi)
        Pq = A[:, None, :] + s[None, :, None] * (B - A)[:, None, :]
        W = (4.0 * math.pi * Pq[..., 0] * Pq[..., 1]) ** 2
        c0, c1 = fc.cell0[fi].long(), fc.cell1[fi].long()
        d = []
        for cell, sg in ((c0, 1.0), (c1, -1.0)):
            lam = self._facet_lam(fi, cell, s)
            _, gphi = _p2_basis(lam, self.grads[cell])
            d.append(sg * torch.einsum("nqkd,nd->nqk", gphi, nrm))
        d = torch.cat(d, 2)
        md = torch.cat([self.qd[c0], self.qd[c1]], 1)
        hbar = self._hbar(c0, c1) if self.h_cell else L
        coef = self.gamma * (hbar ** 2)[:, None] * w[None, :] * L[:, None] * W
        Ji = torch.einsum("nq,nqk,nql->nkl", coef, d, d)
        rows.append(md.repeat_interleave(12, dim=1).flatten())
        cols.append(md.repeat(1, 12).flatten())
        vals.append(Ji.flatten())
        return torch.cat(rows), torch.cat(cols), torch.cat(vals), cvec

    def _hbar(self, c0, c1):
        """avg(CellDiameter) of the two cells of every interior facet."""
        Xc = self.mesh.coords[self.mesh.cells.long()]
        e = torch.stack([Xc[:, 1] - Xc[:, 0], Xc[:, 2] - Xc[:, 0], Xc[:, 2] - Xc[:, 1]], 1)
        hK = e.pow(2).sum(-1).sqrt().max(1).values
        return 0.5 * (hK[c0] + hK[c1])

    def _natural_all(self, degree):
        """+ p_ex (tau.n) W ds on every exterior facet (...MixedPrimal_convergence.py:126): boundary data vector."""
        fc, dev = self.mesh.facets(), self.mesh.device
        m = (degree + 2) // 2
        p, w = np.polynomial.legendre.leggauss(m)
        s = torch.tensor(0.5 * (p + 1.0), device=dev)
        w = torch.tensor(0.5 * w, device=dev)
        ext = torch.nonzero(fc.cell1 < 0).flatten()
        L, nrm, A, B = self._facet_geometry(ext)
        Pq = A[:, None, :] + s[None, :, None] * (B - A)[:, None, :]
        x, y = Pq[..., 0], Pq[..., 1]
        pw = w[None, :] * (4.0 * math.pi * x * y) ** 2 * p_exact(x, y)
        sgn = torch.sign((nrm * self.enorm[ext]).sum(1))
        out = torch.zeros(self.N, dtype=torch.float64, device=dev)
        for slot, lamend in ((0, 1.0 - s), (1, s)):
            out[2 * ext + slot] = (pw * lamend[None, :]).sum(1) * sgn
        return out

    def _facet_tables_cuda(self, degree):
        """The same operator from the CUDA kernels: returns COO (rows, cols, vals) of the q-q facet matrix and
        (rows, vals) of the constant vector, both in kernel storage order (entry-major over facets)."""
        mesh, fc, dev, lib = self.mesh, self.mesh.facets(), self.mesh.device, self.lib
        m = (degree + 2) // 2
        p, w = np.polynomial.legendre.leggauss(m)
        rule = np.ascontiguousarray(np.stack([0.5 * (p + 1.0), 0.5 * w], 1), dtype=np.float64)
        s = torch.tensor(rule[:, 0], device=dev)
        st = _capi.stream_ptr()
        f64 = dict(dtype=torch.float64, device=dev)
        # exterior
        ext = torch.nonzero(fc.cell1 < 0).flatten()
        nfe = ext.numel()
        fv_e = fc.verts[ext].to(torch.int32).contiguous()
        fc_e = fc.cell0[ext].to(torch.int32).contiguous()
        mk = self.bdry.values[ext]
        isin = lambda ms: (torch.zeros_like(mk, dtype=torch.bool) if not ms
                           else torch.stack([mk == m_ for m_ in ms]).any(0))
        nat = torch.ones_like(mk, dtype=torch.bool) if self.nat_markers is None else isin(self.nat_markers)
        kind = (isin(self.adv_markers).to(torch.int32) + 2 * isin(self.dat_markers).to(torch.int32)
                + 4 * nat.to(torch.int32) + (8 if self.r2dir < 0 else 0)).contiguous()
        A = mesh.coords[fv_e[:, 0].long()]
        B = mesh.coords[fv_e[:, 1].long()]
        Pq = A[:, None, :] + s[None, :, None] * (B - A)[:, None, :]
        qex = q_exact(Pq[..., 0], Pq[..., 1]).to(torch.float64).contiguous()
        pex = p_exact(Pq[..., 0], Pq[..., 1]).to(torch.float64).contiguous()
        Jx, cx, nat = torch.empty(36 * nfe, **f64), torch.empty(6 * nfe, **f64), torch.empty(2 * nfe, **f64)
        _capi.check(lib.cm_cg2_exterior_facets(m, rule.ctypes.data, self.stab2, nfe, _capi.ptr(fv_e), _capi.ptr(fc_e),
                                               _capi.ptr(kind), _capi.ptr(mesh.cells), _capi.ptr(mesh.coords),
                                               _capi.ptr(qex), _capi.ptr(pex), _capi.ptr(Jx), _capi.ptr(cx),
                                               _capi.ptr(nat), st))
        qd_e = self.qd[fc_e.long()].T.contiguous()                                    # [6,nfe]
        ia = torch.arange(6, device=dev).repeat_interleave(6)
        ib = torch.arange(6, device=dev).repeat(6)
        rows = [qd_e[ia].flatten()]
        cols = [qd_e[ib].flatten()]
        vals = [Jx]
        cv_rows = [qd_e.flatten(), 2 * ext, 2 * ext + 1]
        cv_vals = [cx, nat]
        # interior
        fi = torch.nonzero(fc.cell1 >= 0).flatten()
        nfi = fi.numel()
        fv_i = fc.verts[fi].to(torch.int32).contiguous()
        fc_i = torch.stack([fc.cell0[fi], fc.cell1[fi]], 1).to(torch.int32).contiguous()
        Ji = torch.empty(144 * nfi, **f64)
        hbar = self._hbar(fc_i[:, 0].long(), fc_i[:, 1].long()).contiguous() if self.h_cell else None
        _capi.check(lib.cm_cg2_interior_facets(m, rule.ctypes.data, self.gamma, nfi, _capi.ptr(fv_i), _capi.ptr(fc_i),
                                               _capi.ptr(mesh.cells), _capi.ptr(mesh.coords), _capi.ptr(hbar),
                                               _capi.ptr(Ji), st))
        md = torch.cat([self.qd[fc_i[:, 0].long()], self.qd[fc_i[:, 1].long()]], 1).T.contiguous()   # [12,nfi]
        ja = torch.arange(12, device=dev).repeat_interleave(12)
        jb = torch.arange(12, device=dev).repeat(12)
        rows.append(md[ja].flatten())
        cols.append(md[jb].flatten())
        vals.append(Ji)
        return torch.cat(rows), torch.cat(cols), torch.cat(vals), torch.cat(cv_rows), torch.cat(cv_vals)

    def _rt_values(self):
        """RT2 basis values at the cell quadrature points [C,nq,8,2] (setup: the projection of sig_ex)."""
        Xc = self.mesh.coords[self.mesh.cells.long()]
        P = torch.stack([self.xq, self.yq], -1)
        lam = self.lamq[None].expand(self.C, -1, -1)
        R = []
        funcs = [(i, a, self.sign[:, i]) for i in range(3) for a in _OTHER[i]] + [(1, 1, None), (2, 2, None)]
        for i, a, sg in funcs:
            c = (1.0 if sg is None else sg) / (2 * self.area)
            R.append(lam[..., a, None] * (P - Xc[:, None, i, :]) * c[:, None, None])
        return torch.stack(R, 2)

    def sigD(self):
        """project(sig_ex, RT2) (:132), unweighted; mass matrix through cm_csr_gather, solve through cm_blu_*."""
        dev, E, C = self.mesh.device, self.E, self.C
        R = self._rt_values()
        a = torch.tensor(self.rule[:, 2], device=dev)[None, :] * (2 * self.area)[:, None]
        sex = torch.stack([self.data["sx"](self.xq, self.yq), self.data["sy"](self.xq, self.yq)], -1)
        Me = torch.einsum("cq,cqad,cqbd->cab", a, R, R)
        be = torch.einsum("cq,cqad,cqd->ca", a, R, sex)
        ns = 2 * E + 2 * C
        rows = self.sd.repeat_interleave(8, dim=1).flatten()
        cols = self.sd.repeat(1, 8).flatten()
        key, rp, col = _csr_from_pairs(rows, cols, ns)
        ptr, order = _gather_lists(torch.searchsorted(key, rows * ns + cols), key.numel())
        vals = torch.empty(key.numel(), dtype=torch.float64, device=dev)
        st = _capi.stream_ptr()
        src = Me.flatten().contiguous()
        _capi.check(self.lib.cm_csr_gather(key.numel(), _capi.ptr(ptr), _capi.ptr(order.to(torch.int32).contiguous()),
                                           _capi.ptr(src), _capi.ptr(vals), 0, st))
        ptr2, order2 = _gather_lists(self.sd.flatten(), ns)
        rhs = torch.empty(ns, dtype=torch.float64, device=dev)
        src2 = be.flatten().contiguous()
        _capi.check(self.lib.cm_csr_gather(ns, _capi.ptr(ptr2), _capi.ptr(order2.to(torch.int32).contiguous()),
                                           _capi.ptr(src2), _capi.ptr(rhs), 0, st))
        lin = BandedLULinearSolver(_Holder(dev, rp, col, ns, vals, rhs))
        lin.setup()
        x = torch.empty_like(rhs)
        lin.solve(rhs, x)
        return x

    # -- per Newton step ---------------------------------------------------------------------------------------
    def assemble(self, u, want_jac=True):
        lib, st, mesh = self.lib, _capi.stream_ptr(), self.mesh
        _capi.check(lib.cm_mixed_m1_cells(
            self.nq, self.rule.ctypes.data, self.D1, self.D2, DOLFIN_EPS, self.gcoef, self.qmode, self.q_react,
            self.q_x2p, self.q_x2q, self.supg_stab, self.C, self.E, self.V,
            _capi.ptr(mesh.cells), _capi.ptr(mesh.coords), _capi.ptr(self.cell_edges), _capi.ptr(self.sign),
            _capi.ptr(u), _capi.ptr(self.f2q), _capi.ptr(self.f3q), _capi.ptr(self.Je) if want_jac else None,
            _capi.ptr(self.Fe), st))
        _capi.check(lib.cm_csr_gather(self.N, _capi.ptr(self.vec_ptr), _capi.ptr(self.vec_idx),
                                      _capi.ptr(self.Fe), _capi.ptr(self.b), 0, st))
        _capi.check(lib.cm_spmv(1, self.N, _capi.ptr(self.rowptr), _capi.ptr(self.col), _capi.ptr(self.Kf),
                                _capi.ptr(u), _capi.ptr(self._ku), st))
        self.b += self._ku
        self.b += self.cvec
        if want_jac:
            self.vals.copy_(self.Kf)
            _capi.check(lib.cm_csr_gather(self.nnz, _capi.ptr(self.mat_ptr), _capi.ptr(self.mat_idx),
                                          _capi.ptr(self.Je), _capi.ptr(self.vals), 1, st))
        return (self.vals if want_jac else None), self.b

    def _apply_bc(self, u, with_matrix):
        _capi.check(self.lib.cm_apply_dirichlet(
            1, self.bc_dofs.numel(), _capi.ptr(self.bc_dofs), _capi.ptr(self.bc_diag), _capi.ptr(self.bc_vals),
            _capi.ptr(u), _capi.ptr(self.rowptr), _capi.ptr(self.vals) if with_matrix else None,
            _capi.ptr(self.b), _capi.stream_ptr()))

    def solve(self, tol=1e-7, maxit=50):
        dev = self.mesh.device
        fc = self.mesh.facets()
        val = torch.zeros(self.N, dtype=torch.float64, device=dev)
        has = torch.zeros(self.N, dtype=torch.bool, device=dev)
        q_marker = 31 if self.galerkin else self.q_bc_marker
        if q_marker is not None:   # q = q_ex at the CG2 nodes of the marked facets (...MixedPrimal_convergence.py:117)
            sel = torch.nonzero((fc.cell1 < 0) & (self.bdry.values == q_marker)).flatten()
            va, vb = fc.verts[sel, 0].long(), fc.verts[sel, 1].long()
            XA, XB = self.mesh.coords[va], self.mesh.coords[vb]
            XM = 0.5 * (XA + XB)
            for d, Xn in ((self.q0 + va, XA), (self.q0 + vb, XB), (self.q0 + self.V + sel, XM)):
                val[d] = q_exact(Xn[:, 0], Xn[:, 1])
                has[d] = True
        if not self.galerkin:      # flux data on the marked facets: project(sig_ex) for the paper system (:132-133);
            le = torch.nonzero((fc.cell1 < 0) & (self.bdry.values == self.s_bc_marker)).flatten()
            sdofs = torch.cat([2 * le, 2 * le + 1])
            # ..._C0IP.py:120,134: s_ex = (0, pi sin(pi x)) has a vanishing normal trace on {x = 0}: zero dofs
            val[sdofs] = self.sigD()[sdofs] if self.variant in ("paper", "c0ip_other") else 0.0
            has[sdofs] = True
        bc = torch.nonzero(has).flatten()
        self.bc_vals = val[bc].contiguous()
        self.bc_dofs = bc.to(torch.int32).contiguous()
        self.bc_diag = self.diag_slot_all[bc].contiguous()
        u = torch.zeros(self.N, dtype=torch.float64, device=dev)
        self.assemble(u, False)
        self._apply_bc(u, False)
        r0 = r = float(torch.linalg.vector_norm(self.b))
        if self._lin is None:
            self._lin = BandedLULinearSolver(_Holder(dev, self.rowptr, self.col, self.N, self.vals, self.b))
        dx = torch.empty_like(u)
        it = 0
        while not (r < tol or r / r0 < tol) and it < maxit:
            self.assemble(u, True)
            self._apply_bc(u, True)
            self._lin.setup()
            self._lin.solve(self.b, dx)
            u -= dx
            it += 1
            self.assemble(u, False)
            self._apply_bc(u, False)
            r = float(torch.linalg.vector_norm(self.b))
            if r != r:
                raise RuntimeError("Newton solver diverged (NaN residual)")
        if not (r < tol or r / r0 < tol):
            raise RuntimeError("Newton solver did not converge because maximum number of iterations reached")
        self.newton_its = it
        return u

    def errors(self, u, full_grad=None):
        """e_div(s), e_0(p), e_1(q); e_1(q) with d/dx only (test_for_paper_convergence.py:175, the 'paper' default) or
        with the whole gradient (the MixedPrimal scripts, e.g. ..._C0IP_otherBCs_modified.py:188)."""
        if full_grad is None:
            full_grad = self.variant != "paper"
        d, x, y = self.data, self.xq, self.yq
        exq = torch.stack([d["sx"](x, y), d["sy"](x, y), d["drs"](x, y), p_exact(x, y), q_exact(x, y),
                           d["qx"](x, y), d["qy"](x, y)], -1).to(torch.float64).contiguous()
        work = torch.empty(int(self.lib.cm_mixed_workspace_doubles()), dtype=torch.float64,
                           device=self.mesh.device)
        _capi.check(self.lib.cm_mixed_m1_errors(
            self.nq, self.rule.ctypes.data, 1 if full_grad else 0, self.C, self.E, self.V,
            _capi.ptr(self.mesh.cells),
            _capi.ptr(self.mesh.coords), _capi.ptr(self.cell_edges), _capi.ptr(self.sign), _capi.ptr(u),
            _capi.ptr(exq), _capi.ptr(work), _capi.stream_ptr()))
        return tuple(math.sqrt(float(v)) for v in work[:3].cpu())


# ---------------------------------------------------------------------------------------------------------
# scalar CG2 pure-advection problems (SUPG_pure_advection_convergence.py, C0IP_pure_advection_spherical_convergence.py)
# ---------------------------------------------------------------------------------------------------------
class PureAdvectionCG2:
    """solve(lhs == rhs, q_h, bcs=bcQ) of the two scalar CG2 scripts on the GPU.  Their forms are the q-equation of the
    RT2 x DG1 x CG2 systems above without the p coupling, so the element / facet tensors are the q-q block of
    cm_mixed_m1_cells (qmode 1 with stab = 0.1: SUPG_pure_advection_convergence.py:106-109; qmode 0 with reaction:
    C0IP_pure_advection_spherical_convergence.py:109-110) and cm_cg2_exterior_facets / cm_cg2_interior_facets
    (:111-113, penalty +stab/(deg+1)^3.5 avg(he)^2, he = CellDiameter); that block is cut out of the mixed CSR
    pattern once per mesh, cm_apply_dirichlet imposes q = q_ex on the inlet (31, SUPG :101) or the outlet (32, C0IP
    :102) and cm_blu_* solves.  Markers inlet 31 / outlet 32 / char 33 (:72-78).  Dofs: vertices, then edges.
    NB: with the C0IP script's positive stab = 0.005 the penalty enters against the sign of the (negative) form
    and the discrete problem is nearly singular (the later ..._MixedPrimal_convergence_C0IP.py:147 flips the sign);
    `stab` is an argument so that both can be run."""

    STAB = {"supg": 0.1, "c0ip": 0.005}

    def __init__(self, mesh, bdry, kind, stab=None, tables_only=False):
        """tables_only: host tables (pattern of the q-q block, boundary dofs) without touching the GPU -- CPU tests."""
        if kind not in self.STAB:
            raise ValueError(f"unknown kind '{kind}'")
        self.kind = kind
        self.stab = stab = self.STAB[kind] if stab is None else float(stab)
        self.sys = S = PaperMixedSystemK1(mesh, bdry, D1=1e-3, D2=1.5e-3, stab=stab if kind == "c0ip" else 0.0,
                                          stab2=0.0, degree=4, variant="pure_" + kind,
                                          supg_stab=stab if kind == "supg" else 0.0, tables_only=tables_only)
        dev = mesh.device
        self.N = N = S.V + S.E
        rows = torch.repeat_interleave(torch.arange(S.N, device=dev), torch.diff(S.rowptr.long()))
        keep = (rows >= S.q0) & (S.col >= S.q0)
        self.slots = torch.nonzero(keep).flatten()
        r = rows[keep] - S.q0
        self.col = (S.col[keep] - S.q0).to(torch.int32).contiguous()
        rp = torch.zeros(N + 1, dtype=torch.int64, device=dev)
        rp[1:] = torch.cumsum(torch.bincount(r, minlength=N), 0)
        self.rowptr = rp.to(torch.int32).contiguous()
        self.diag_slot = (torch.searchsorted(r * N + self.col.long(), torch.arange(N, device=dev) * (N + 1))
                          - rp[:-1]).to(torch.int32).contiguous()          # position inside the row
        fc = mesh.facets()
        sel = torch.nonzero((fc.cell1 < 0) & (bdry.values == (31 if kind == "supg" else 32))).flatten()
        va, vb = fc.verts[sel, 0].long(), fc.verts[sel, 1].long()
        X = mesh.coords
        val = torch.zeros(N, dtype=torch.float64, device=dev)
        has = torch.zeros(N, dtype=torch.bool, device=dev)
        for d, Xn in ((va, X[va]), (vb, X[vb]), (S.V + sel, 0.5 * (X[va] + X[vb]))):
            val[d] = q_exact(Xn[:, 0], Xn[:, 1])
            has[d] = True
        bc = torch.nonzero(has).flatten()
        self.bc_dofs, self.bc_vals = bc.to(torch.int32).contiguous(), val[bc].contiguous()
        self.bc_diag = self.diag_slot[bc].contiguous()
        self._lin = None

    def assemble(self):
        """(values of the CG2 matrix on (rowptr, col), right-hand side) before the boundary condition."""
        S = self.sys
        vals, F0 = S.assemble(torch.zeros(S.N, dtype=torch.float64, device=S.mesh.device), True)
        if getattr(self, "vals", None) is None:   # persistent buffers: the LU holder keeps pointing at them
            self.vals = vals[self.slots].contiguous()
            self.b = (-F0[S.q0:]).contiguous()    # residual at q = 0 is -rhs
        else:
            self.vals.copy_(vals[self.slots])
            self.b.copy_(-F0[S.q0:])
        return self.vals, self.b

    def solve(self):
        S, dev = self.sys, self.sys.mesh.device
        self.assemble()
        q = torch.zeros(self.N, dtype=torch.float64, device=dev)
        # cm_apply_dirichlet writes identity rows and b[row] = x[row] - g (the Newton form of DirichletBC.apply); with
        # x = 0 and the data negated the rows of the linear problem read q = g
        _capi.check(S.lib.cm_apply_dirichlet(
            1, self.bc_dofs.numel(), _capi.ptr(self.bc_dofs), _capi.ptr(self.bc_diag),
            _capi.ptr((-self.bc_vals).contiguous()), _capi.ptr(q), _capi.ptr(self.rowptr), _capi.ptr(self.vals),
            _capi.ptr(self.b), _capi.stream_ptr()))
        if self._lin is None:
            self._lin = BandedLULinearSolver(_Holder(dev, self.rowptr, self.col, self.N, self.vals, self.b))
        self._lin.setup()
        self._lin.solve(self.b, q)
        return q

    def error(self, q):
        """e_1(q) = (int ((q_ex - q_h)^2 + |grad(q_ex - q_h)|^2) W dx)^(1/2) (:120-122), degree-4 rule."""
        S = self.sys
        u = torch.zeros(S.N, dtype=torch.float64, device=S.mesh.device)
        u[S.q0:] = q
        return S.errors(u)[2]


# ---------------------------------------------------------------------------------------------------------
# production mixed formulation DG2 x RT3 x CG3 (advection_diffusion_mixed.py:94-269)
# ---------------------------------------------------------------------------------------------------------
class ProductionMixedSystem:
    """solve_problem_instance of advection_diffusion_mixed.py (deg = 2: DG2 x RT3 x CG3, :147-150) on the GPU:
    cm_mixed_p2_cells (31 x 31 element tensors, production Gstar), cm_csr_gather, cm_apply_dirichlet, cm_blu_*
    ('mumps', :223).  Markers right = 21, top = 22, left = 23, bottom = 24.  The natural p data on the right / top
    facets (:203-204) and the Dirichlet values are boundary data tabulated once with torch; the RT3 projection of
    the flux (:243) assembles its mass matrix with the same gather machinery and solves it with the LU."""

    RIGHT, TOP, LEFT, BOTTOM = 21, 22, 23, 24
    CELL_DEGREE, FACET_POINTS = 10, 6

    def __init__(self, mesh, bdry, concentration, sigma, gamma, V1, D1, D2, dt=None):
        _capi.require_cuda(mesh.coords)
        self.lib = _capi.lib()
        self.mesh, self.bdry = mesh, bdry
        self.c, self.sigma, self.gamma, self.V1 = float(concentration), float(sigma), float(gamma), float(V1)
        self.D1, self.D2 = float(D1), float(D2)
        self.inv_dt = 0.0 if dt is None else 1.0 / float(dt)
        dev = mesh.device
        fc = mesh.facets()
        V, C, E = mesh.num_vertices(), mesh.num_cells(), fc.num
        self.V, self.C, self.E = V, C, E
        self.np_, self.ns = 6 * C, 3 * E + 6 * C
        self.q0 = self.np_ + self.ns
        self.N = N = self.q0 + V + 2 * E + C
        cells = mesh.cells.long()
        X = mesh.coords
        fa = torch.stack([cells[:, 1], cells[:, 0], cells[:, 0]], 1)
        fb = torch.stack([cells[:, 2], cells[:, 2], cells[:, 1]], 1)
        key = torch.minimum(fa, fb) * V + torch.maximum(fa, fb)
        allkey = fc.verts[:, 0].long() * V + fc.verts[:, 1].long()
        ce = torch.searchsorted(allkey, key)
        A, B = X[fc.verts[:, 0].long()], X[fc.verts[:, 1].long()]
        t = B - A
        self.elen = t.pow(2).sum(1).sqrt()
        self.enorm = torch.stack([t[:, 1], -t[:, 0]], 1) / self.elen[:, None]
        Xc = X[cells]
        self.Xc = Xc
        self.sign = torch.sign(((0.5 * (A + B)[ce] - Xc) * self.enorm[ce]).sum(-1)).to(torch.float64).contiguous()
        self.cell_edges = ce.to(torch.int32).contiguous()
        det = ((Xc[:, 1, 0] - Xc[:, 0, 0]) * (Xc[:, 2, 1] - Xc[:, 0, 1])
               - (Xc[:, 2, 0] - Xc[:, 0, 0]) * (Xc[:, 1, 1] - Xc[:, 0, 1]))
        g1 = torch.stack([Xc[:, 2, 1] - Xc[:, 0, 1], -(Xc[:, 2, 0] - Xc[:, 0, 0])], 1) / det[:, None]
        g2 = torch.stack([-(Xc[:, 1, 1] - Xc[:, 0, 1]), Xc[:, 1, 0] - Xc[:, 0, 0]], 1) / det[:, None]
        self.grads = torch.stack([-(g1 + g2), g1, g2], 1)
        self.area = 0.5 * det.abs()
        pts, wts = collapsed_triangle_rule(self.CELL_DEGREE)
        self.rule = np.ascontiguousarray(np.concatenate([pts, wts[:, None]], 1), dtype=np.float64)
        self.nq = self.rule.shape[0]
        self.lamq = torch.tensor(np.stack([1.0 - pts[:, 0] - pts[:, 1], pts[:, 0], pts[:, 1]], 1), device=dev)
        self.Pq = torch.einsum("qi,cid->cqd", self.lamq, Xc)
        cid = torch.arange(C, device=dev)[:, None]
        ar = lambda k: torch.arange(k, device=dev)[None, :]
        self.pd = 6 * cid + ar(6)
        s0 = self.np_
        self.sd = torch.cat([s0 + 3 * ce[:, i, None] + ar(3) for i in range(3)] + [s0 + 3 * E + 6 * cid + ar(6)], 1)
        q0 = self.q0
        self.qd = torch.cat([q0 + cells] + [q0 + V + 2 * ce[:, i, None] + ar(2) for i in range(3)]
                            + [q0 + V + 2 * E + cid], 1)
        self.cell_dofs = cd = torch.cat([self.pd, self.sd, self.qd], 1)                  # [C,31]
        rows_c = cd.repeat_interleave(31, dim=1).flatten()
        cols_c = cd.repeat(1, 31).flatten()
        self._key, self.rowptr, self.col = _csr_from_pairs(rows_c, cols_c, N)
        self.nnz = int(self.col.numel())
        slot_c = torch.searchsorted(self._key, rows_c * N + cols_c)
        ent = (torch.arange(961, device=dev)[None, :] * C + torch.arange(C, device=dev)[:, None]).flatten()
        self.mat_ptr, order = _gather_lists(slot_c, self.nnz)
        self.mat_idx = ent[order].to(torch.int32).contiguous()
        entv = (torch.arange(31, device=dev)[None, :] * C + torch.arange(C, device=dev)[:, None]).flatten()
        self.vec_ptr, order = _gather_lists(cd.flatten(), N)
        self.vec_idx = entv[order].to(torch.int32).contiguous()
        self.diag_slot_all = (torch.searchsorted(self._key, torch.arange(N, device=dev) * (N + 1))
                              - self.rowptr[:-1].long()).to(torch.int32)
        self.nat = self._natural()
        self._bc_tables()
        f64 = dict(dtype=torch.float64, device=dev)
        self.Je = torch.empty(961 * C, **f64)
        self.Fe = torch.empty(31 * C, **f64)
        self.vals = torch.zeros(self.nnz, **f64)
        self.b = torch.zeros(N, **f64)
        self.u = torch.zeros(N, **f64)
        self.p_old = torch.zeros(6 * C, **f64)
        self._lin = None
        self.newton_its, self.converged = 0, False

    # -- boundary data ---------------------------------------------------------------------------------------
    def p_initial(self, x, y):          # PInitial.eval (:40-44)
        sb = self.sigma * torch.exp(-4 / 3 * math.pi * self.gamma * x ** 3)
        base = self.c / self.V1 * torch.exp(-4 / 3 * math.pi * self.c * x ** 3)
        return torch.where(y > 0, base * (1 - sb / torch.where(y > 0, y, torch.ones_like(y))), base)

    def q_initial(self, x):             # QInitial.eval (:87-88)
        return 1 / (4 * math.pi * self.V1) * torch.exp(-4 / 3 * math.pi * self.c * x ** 3)

    def _facet_geometry(self, fidx):
        mesh, fc = self.mesh, self.mesh.facets()
        A = mesh.coords[fc.verts[fidx, 0].long()]
        B = mesh.coords[fc.verts[fidx, 1].long()]
        t = B - A
        L = t.pow(2).sum(1).sqrt()
        nrm = torch.stack([t[:, 1], -t[:, 0]], 1) / L[:, None]
        opp = mesh.coords[mesh.cells.long()[fc.cell0[fidx].long(), fc.loc0[fidx].long()]]
        flip = ((opp - A) * nrm).sum(1) > 0
        return L, torch.where(flip[:, None], -nrm, nrm), A, B

    def _gauss(self):
        p, w = np.polynomial.legendre.leggauss(self.FACET_POINTS)
        dev = self.mesh.device
        return torch.tensor(0.5 * (p + 1.0), device=dev), torch.tensor(0.5 * w, device=dev)

    def _natural(self):
        """+ p_right (tau.n) W ds(right) + p_top (tau.n) W ds(top) (:203-204): the facet's own three functions."""
        fc, dev = self.mesh.facets(), self.mesh.device
        sel = torch.nonzero((fc.cell1 < 0) & ((self.bdry.values == self.RIGHT) | (self.bdry.values == self.TOP))).flatten()
        L, nrm, A, B = self._facet_geometry(sel)
        s, w = self._gauss()
        P = A[:, None, :] + s[None, :, None] * (B - A)[:, None, :]
        x, y = P[..., 0], P[..., 1]
        W = (4.0 * math.pi * x * y) ** 2
        sgn = torch.sign((nrm * self.enorm[sel]).sum(1))
        pw = w[None, :] * W * self.p_initial(x, y)
        out = torch.zeros(self.N, dtype=torch.float64, device=dev)
        for m, tr in enumerate(((1 - s) ** 2, (1 - s) * s, s ** 2)):
            out[self.np_ + 3 * sel + m] = (pw * tr[None, :]).sum(1) * sgn
        return out

    def _bc_tables(self):
        """s = 0 on the left facets (:181-182), q = q_initial at the CG3 nodes of the right facets (:187)."""
        fc, mesh, dev = self.mesh.facets(), self.mesh, self.mesh.device
        left = torch.nonzero((fc.cell1 < 0) & (self.bdry.values == self.LEFT)).flatten()
        sdofs = (self.np_ + 3 * left[:, None] + torch.arange(3, device=dev)[None, :]).flatten()
        right = torch.nonzero((fc.cell1 < 0) & (self.bdry.values == self.RIGHT)).flatten()
        va, vb = fc.verts[right, 0].long(), fc.verts[right, 1].long()
        xa, xb = mesh.coords[va, 0], mesh.coords[vb, 0]
        val = torch.zeros(self.N, dtype=torch.float64, device=dev)
        has = torch.zeros(self.N, dtype=torch.bool, device=dev)
        for d, xv in ((self.q0 + va, xa), (self.q0 + vb, xb), (self.q0 + self.V + 2 * right, (2 * xa + xb) / 3),
                      (self.q0 + self.V + 2 * right + 1, (xa + 2 * xb) / 3)):
            val[d] = self.q_initial(xv)
            has[d] = True
        has[sdofs] = True
        dofs = torch.nonzero(has).flatten()
        self.bc_dofs = dofs.to(torch.int32).contiguous()
        self.bc_vals = val[dofs].contiguous()
        self.bc_diag = self.diag_slot_all[dofs].contiguous()

    # -- per Newton step ----------------------------------------------------------------------------------------
    def assemble(self, u, want_jac=True):
        lib, st, mesh = self.lib, _capi.stream_ptr(), self.mesh
        _capi.check(lib.cm_mixed_p2_cells(
            self.nq, self.rule.ctypes.data, self.D1, self.D2, self.c, self.inv_dt, self.C, self.E, self.V,
            _capi.ptr(mesh.cells), _capi.ptr(mesh.coords), _capi.ptr(self.cell_edges), _capi.ptr(self.sign),
            _capi.ptr(u), _capi.ptr(self.p_old) if self.inv_dt else None,
            _capi.ptr(self.Je) if want_jac else None, _capi.ptr(self.Fe), st))
        _capi.check(lib.cm_csr_gather(self.N, _capi.ptr(self.vec_ptr), _capi.ptr(self.vec_idx),
                                      _capi.ptr(self.Fe), _capi.ptr(self.b), 0, st))
        self.b += self.nat
        if want_jac:
            _capi.check(lib.cm_csr_gather(self.nnz, _capi.ptr(self.mat_ptr), _capi.ptr(self.mat_idx),
                                          _capi.ptr(self.Je), _capi.ptr(self.vals), 0, st))
        return (self.vals if want_jac else None), self.b

    def _apply_bc(self, u, with_matrix):
        _capi.check(self.lib.cm_apply_dirichlet(
            1, self.bc_dofs.numel(), _capi.ptr(self.bc_dofs), _capi.ptr(self.bc_diag), _capi.ptr(self.bc_vals),
            _capi.ptr(u), _capi.ptr(self.rowptr), _capi.ptr(self.vals) if with_matrix else None,
            _capi.ptr(self.b), _capi.stream_ptr()))

    def solve(self, tol=1e-8, maxit=10):
        """Newton from the current self.u with 'mumps' (:219-226): abs / rel tolerance 1e-8, at most 10 iterations."""
        u, dev = self.u, self.mesh.device
        self.assemble(u, False)
        self._apply_bc(u, False)
        r0 = r = float(torch.linalg.vector_norm(self.b))
        if self._lin is None:
            self._lin = BandedLULinearSolver(_Holder(dev, self.rowptr, self.col, self.N, self.vals, self.b))
        dx = torch.empty_like(u)
        it = 0
        while not (r < tol or (r0 > 0 and r / r0 < tol)) and it < maxit:
            self.assemble(u, True)
            self._apply_bc(u, True)
            self._lin.setup()
            self._lin.solve(self.b, dx)
            u -= dx
            it += 1
            self.assemble(u, False)
            self._apply_bc(u, False)
            r = float(torch.linalg.vector_norm(self.b))
            if r != r:
                raise RuntimeError("Newton solver diverged (NaN residual)")
        self.newton_its, self.converged = it, bool(r < tol or (r0 > 0 and r / r0 < tol))
        if not self.converged:
            raise RuntimeError("Newton solver did not converge because maximum number of iterations reached")
        return u

    # -- post-processing (:243, 261-263) -------------------------------------------------------------------------
    def _rt_values(self, lam, P, cells=None):
        idx = slice(None) if cells is None else cells
        Xc, area, sign = self.Xc[idx], self.area[idx], self.sign[idx]
        funcs = []
        for i in range(3):
            j, k = _OTHER[i]
            funcs += [(i, j, j, sign[:, i]), (i, j, k, sign[:, i]), (i, k, k, sign[:, i])]
        for i in (1, 2):
            funcs += [(i, i, m, None) for m in range(3)]
        R = []
        for i, a, b, sg in funcs:
            c = (1.0 if sg is None else sg) / (2 * area)
            R.append((lam[..., a] * lam[..., b])[..., None] * (P - Xc[:, None, i, :]) * c[:, None, None])
        return torch.stack(R, 2)

    # -- L2 projections (project(): advection_diffusion_mixed.py:238-254) -----------------------------------------
    def _p3_values(self):
        """CG3 basis at the cell quadrature points [nq, 10] (vertices, two nodes per facet, bubble; oracle p3_basis)."""
        lam = self.lamq
        val = [0.5 * lam[:, i] * (3 * lam[:, i] - 1) * (3 * lam[:, i] - 2) for i in range(3)]
        for j, k in _OTHER:
            lj, lk = lam[:, j], lam[:, k]
            val += [4.5 * lj * lk * (3 * lj - 1), 4.5 * lk * lj * (3 * lk - 1)]
        val.append(27 * lam[:, 0] * lam[:, 1] * lam[:, 2])
        return torch.stack(val, -1)

    def _p2_values(self):
        lam = self.lamq
        return torch.cat([lam * (2 * lam - 1), torch.stack([4 * lam[:, j] * lam[:, k] for j, k in _OTHER], -1)], -1)

    def _qweights(self):
        return torch.tensor(self.rule[:, 2], device=self.mesh.device)[None, :] * (2 * self.area)[:, None]   # [C,nq]

    def _global_projector(self, kind):
        """Mass matrix of the space (constant per mesh): assembled with cm_csr_gather and factorised with cm_blu_*
        ONCE; returns (basis values at the quadrature points, local dof table, gather lists, solver)."""
        cache = self.__dict__.setdefault("_proj_cache", {})
        if kind in cache:
            return cache[kind]
        dev = self.mesh.device
        a = self._qweights()
        if kind == "rt3":
            lam = self.lamq[None].expand(self.C, -1, -1)
            R = self._rt_values(lam, self.Pq)                               # [C,nq,15,2]
            Me = torch.einsum("cq,cqad,cqbd->cab", a, R, R)
            sl, n = self.sd - self.np_, self.ns
        else:
            R = self._p3_values()                                           # [nq,10]
            Me = torch.einsum("cq,qa,qb->cab", a, R, R)
            sl, n = self.qd - self.q0, self.V + 2 * self.E + self.C
        k = sl.shape[1]
        rows = sl.repeat_interleave(k, dim=1).flatten()
        cols = sl.repeat(1, k).flatten()
        key, rp, col = _csr_from_pairs(rows, cols, n)
        ptr, order = _gather_lists(torch.searchsorted(key, rows * n + cols), key.numel())
        vals = torch.empty(key.numel(), dtype=torch.float64, device=dev)
        st = _capi.stream_ptr()
        src = Me.flatten().contiguous()
        _capi.check(self.lib.cm_csr_gather(key.numel(), _capi.ptr(ptr), _capi.ptr(order.to(torch.int32).contiguous()),
                                           _capi.ptr(src), _capi.ptr(vals), 0, st))
        ptr2, order2 = _gather_lists(sl.flatten(), n)
        rhs = torch.empty(n, dtype=torch.float64, device=dev)
        lin = BandedLULinearSolver(_Holder(dev, rp, col, n, vals, rhs))
        lin.setup()
        cache[kind] = (R, sl, (ptr2, order2.to(torch.int32).contiguous()), rhs, lin, n)
        return cache[kind]

    def project(self, kind, fq):
        """Unweighted L2 projection of a field given at the cell quadrature points (fq [C,nq], or [C,nq,2] for
        'rt3') onto 'dg2' (cell-local: the reference mass matrix is inverted once on the host), 'cg3' or 'rt3'."""
        a = self._qweights()
        if kind == "dg2":
            Phi = self._p2_values()                                          # [nq,6]
            w = torch.tensor(self.rule[:, 2], device=self.mesh.device)
            Minv = torch.linalg.inv(torch.einsum("q,qa,qb->ab", w, Phi, Phi).cpu()).to(Phi.device)   # 6 x 6, host
            be = torch.einsum("cq,qa,cq->ca", a, Phi, fq)
            return (be / (2 * self.area)[:, None]) @ Minv.T                  # [C,6] -> DG2 dofs, cell-major
        R, sl, (ptr2, order2), rhs, lin, n = self._global_projector(kind)
        be = torch.einsum("cq,cqad,cqd->ca", a, R, fq) if kind == "rt3" else torch.einsum("cq,qa,cq->ca", a, R, fq)
        src2 = be.flatten().contiguous()
        _capi.check(self.lib.cm_csr_gather(n, _capi.ptr(ptr2), _capi.ptr(order2), _capi.ptr(src2), _capi.ptr(rhs), 0,
                                           _capi.stream_ptr()))
        out = torch.empty_like(rhs)
        lin.solve(rhs, out)
        return out

    def evaluate(self, kind, coef):
        """Field values at the cell quadrature points from a coefficient vector of 'dg2' / 'cg3' / 'rt3'."""
        if kind == "dg2":
            return coef.reshape(self.C, 6) @ self._p2_values().T
        if kind == "cg3":
            return coef[self.qd - self.q0] @ self._p3_values().T
        R = self._global_projector("rt3")[0]
        return torch.einsum("cqad,ca->cqd", R, coef[self.sd - self.np_])

    def flux(self, u=None):
        """project(D 4 pi r2^2 s_h, RT3) (:243): RT3 coefficient vector [3E + 6C]."""
        u = self.u if u is None else u
        x = self.Pq[..., 0]
        D = torch.tensor([self.D2, self.D1], device=self.mesh.device)[None, None, :]
        s = self.evaluate("rt3", u[self.np_:self.q0])
        return self.project("rt3", 4 * math.pi * (x * x)[..., None] * s * D)

    def s_initial(self, x, y):          # SInitial.eval (:59-68): (0, -c/V1 exp(-4/3 pi c x^3) sigma_b(x) / y^2)
        sb = self.sigma * torch.exp(-4 / 3 * math.pi * self.gamma * x ** 3)
        ok = y * y > 0
        v = -self.c / self.V1 * torch.exp(-4 / 3 * math.pi * self.c * x ** 3) * sb / torch.where(ok, y * y, torch.ones_like(y))
        return torch.stack([torch.zeros_like(x), torch.where(ok, v, torch.zeros_like(v))], -1)

    def postprocess(self, u=None):
        """The fields advection_diffusion_mixed.py:238-254 writes per time step, as coefficient vectors:
        p = project(4 pi r2^2 max(p_h, 0), DG2), q = project(max(q_h, 0), CG3), flux = project(D 4 pi r2^2 s_h, RT3),
        dif = project(p - p_approx, DG2), dif_flux = project(flux - flux_approx, RT3), with p_approx = 4 pi r2^2
        interpolate(PInitial, DG2) and flux_approx = D 4 pi r2^2 interpolate(SInitial, RT3) (:232-233).  Deviation:
        the RT3 interpolant of SInitial is replaced by its L2 projection (FIAT's moment-based RT3 interpolation is
        not restated; both are third-order approximations of the same field)."""
        u = self.u if u is None else u
        x, y = self.Pq[..., 0], self.Pq[..., 1]
        fourpix2 = 4 * math.pi * x * x
        p_h = self.evaluate("dg2", u[:self.np_])
        q_h = self.evaluate("cg3", u[self.q0:])
        p = self.project("dg2", fourpix2 * torch.clamp(p_h, min=0.0))
        q = self.project("cg3", torch.clamp(q_h, min=0.0))
        fl = self.flux(u)
        Xc = self.Xc
        nodes = torch.cat([Xc, 0.5 * (Xc[:, [1, 0, 0]] + Xc[:, [2, 2, 1]])], 1)             # DG2 nodes [C,6,2]
        p_steady = self.p_initial(nodes[..., 0], nodes[..., 1]).flatten()
        p_approx_q = fourpix2 * self.evaluate("dg2", p_steady)
        dif = self.project("dg2", self.evaluate("dg2", p.flatten()) - p_approx_q)
        D = torch.tensor([self.D2, self.D1], device=self.mesh.device)[None, None, :]
        s_int = self.project("rt3", self.s_initial(x, y))
        flux_approx_q = fourpix2[..., None] * self.evaluate("rt3", s_int) * D
        dif_flux = self.project("rt3", self.evaluate("rt3", fl) - flux_approx_q)
        return {"p": p.flatten(), "q": q, "flux": fl, "dif": dif.flatten(), "dif_flux": dif_flux}

    def vertex_values(self, kind, coef):
        """What XDMFFile.write stores: the field at the mesh vertices ([V], or [V,2] for 'rt3'); discontinuous
        fields are averaged over the cells around a vertex."""
        cells = self.mesh.cells.long()
        V = self.V
        cnt = torch.zeros(V, dtype=torch.float64, device=coef.device).index_add_(
            0, cells.flatten(), torch.ones(cells.numel(), dtype=torch.float64, device=coef.device))
        if kind == "cg3":
            return coef[:V].clone()
        if kind == "dg2":
            vals = coef.reshape(self.C, 6)[:, :3]
            return torch.zeros(V, dtype=torch.float64, device=coef.device).index_add_(0, cells.flatten(), vals.flatten()) / cnt
        lam = torch.eye(3, dtype=torch.float64, device=coef.device)[None].expand(self.C, -1, -1)   # the three vertices
        R = self._rt_values(lam, self.Xc)
        f = torch.einsum("cqad,ca->cqd", R, coef[self.sd - self.np_])                              # [C,3,2]
        out = torch.zeros(V, 2, dtype=torch.float64, device=coef.device).index_add_(0, cells.flatten(), f.reshape(-1, 2))
        return out / cnt[:, None]

    def bottom_flux(self, fl):
        """(r2 part, r1 part, total) of V1 * assemble(dot(flux, n) 4 pi r1^2 ds(bottom)) (:261-263)."""
        fc, mesh = self.mesh.facets(), self.mesh
        sel = torch.nonzero((fc.cell1 < 0) & (self.bdry.values == self.BOTTOM)).flatten()
        L, nrm, A, B = self._facet_geometry(sel)
        s, w = self._gauss()
        cell = fc.cell0[sel].long()
        cv = mesh.cells.long()[cell]
        va = (fc.verts[sel, 0].long()[:, None] == cv).to(torch.float64)
        vb = (fc.verts[sel, 1].long()[:, None] == cv).to(torch.float64)
        lam = va[:, None, :] * (1.0 - s)[None, :, None] + vb[:, None, :] * s[None, :, None]
        P = A[:, None, :] + s[None, :, None] * (B - A)[:, None, :]
        R = self._rt_values(lam, P, cell)
        f = torch.einsum("nqad,na->nqd", R, fl[(self.sd - self.np_)[cell]])
        wgt = w[None, :] * L[:, None] * 4 * math.pi * P[..., 1] ** 2
        r2 = self.V1 * float((wgt * f[..., 0] * nrm[:, None, 0]).sum())
        r1 = self.V1 * float((wgt * f[..., 1] * nrm[:, None, 1]).sum())
        return r2, r1, r2 + r1


class XDMFFile:
    """Minimal XDMFFile(filename).write(name, vertex_values, t) (advection_diffusion_mixed.py:153-156, 246-256 with
    functions_share_mesh = True): the mesh once, one temporal grid per time with one Attribute per written
    function (vertex values; vectors get a zero third component, like DOLFIN).  Heavy data go to a raw binary
    side file <name>.bin (XDMF Format="Binary", little endian, 8-byte reals / 4-byte ints): no HDF5 here."""

    def __init__(self, filename, mesh):
        import os
        self.filename = filename
        self.binname = os.path.splitext(filename)[0] + ".bin"
        self.mesh = mesh
        self._bin = open(self.binname, "wb")
        self._off = 0
        self._steps = []            # [(t, [(name, kind, offset, n)])]
        self.parameters = {"rewrite_function_mesh": False, "flush_output": True, "functions_share_mesh": True}
        self._cells = self._put(mesh.cells.to(torch.int32).cpu().numpy())
        X = mesh.coords.cpu().numpy()
        self._coords = self._put(np.ascontiguousarray(X, dtype="<f8"))
        self._flush()

    def _put(self, arr):
        arr = np.ascontiguousarray(arr)
        off = self._off
        self._bin.write(arr.tobytes())
        self._off += arr.nbytes
        return off

    def write(self, name, values, t):
        v = values.detach().cpu().numpy()
        if v.ndim == 2:
            v = np.concatenate([v, np.zeros((v.shape[0], 1))], 1)
        off = self._put(np.ascontiguousarray(v, dtype="<f8"))
        if not self._steps or self._steps[-1][0] != t:
            self._steps.append((t, []))
        self._steps[-1][1].append((name, "Vector" if v.ndim == 2 else "Scalar", off, v.shape))
        if self.parameters["flush_output"]:
            self._flush()

    def _flush(self):
        import os
        self._bin.flush()
        V, C = self.mesh.num_vertices(), self.mesh.num_cells()
        b = os.path.basename(self.binname)
        topo = (f'<Topology TopologyType="Triangle" NumberOfElements="{C}"><DataItem Format="Binary" DataType="Int" '
                f'Precision="4" Endian="Little" Seek="{self._cells}" Dimensions="{C} 3">{b}</DataItem></Topology>')
        geom = (f'<Geometry GeometryType="XY"><DataItem Format="Binary" DataType="Float" Precision="8" Endian="Little" '
                f'Seek="{self._coords}" Dimensions="{V} 2">{b}</DataItem></Geometry>')
        out = ['<?xml version="1.0"?>', '<Xdmf Version="3.0"><Domain>',
               '<Grid Name="TimeSeries" GridType="Collection" CollectionType="Temporal">']
        for t, fields in self._steps:
            out.append(f'<Grid Name="mesh" GridType="Uniform">{topo}{geom}<Time Value="{t}"/>')
            for name, kind, off, shape in fields:
                dims = " ".join(str(d) for d in shape)
                out.append(f'<Attribute Name="{name}" AttributeType="{kind}" Center="Node"><DataItem Format="Binary" '
                           f'DataType="Float" Precision="8" Endian="Little" Seek="{off}" Dimensions="{dims}">{b}'
                           '</DataItem></Attribute>')
            out.append('</Grid>')
        out += ['</Grid>', '</Domain></Xdmf>']
        with open(self.filename, "w") as f:
            f.write("\n".join(out))

    def close(self):
        self._flush()
        self._bin.close()


def solve_problem_instance_mixed(concentration, t_final, dt, mesh, bdry, sigma, gamma, results, V1, D1, D2,
                                 output_filename=None, steady_state=False):
    """advection_diffusion_mixed.py:94-269 with its own discretisation (DG2 x RT3 x CG3) on the GPU: same
    signature, `results` row (:268) and return value (flux, total_flux).  With an output_filename the fields of
    :238-254 (clipped p, clipped q, flux, and their differences to the steady-state approximation) are written per
    time step as vertex values into an XDMF time series (XDMFFile above)."""
    system = ProductionMixedSystem(mesh, bdry, concentration, sigma, gamma, V1, D1, D2, None if steady_state else dt)
    if steady_state:
        t_final = 0
    t = 0
    if not steady_state:   # p_old = interpolate(p_initial, DG2): nodal values at the vertices and edge mid points
        Xc = system.Xc
        nodes = torch.cat([Xc, 0.5 * (Xc[:, [1, 0, 0]] + Xc[:, [2, 2, 1]])], 1)            # [C,6,2]
        system.p_old.copy_(system.p_initial(nodes[..., 0], nodes[..., 1]).flatten())
    out = XDMFFile(output_filename, mesh) if output_filename else None
    while t <= t_final:
        print("t=%.3f" % t)
        system.solve()
        if out is not None:                            # :238-254
            f = system.postprocess()
            for name, kind in (("p", "dg2"), ("q", "cg3"), ("flux", "rt3"), ("dif", "dg2"), ("dif_flux", "rt3")):
                out.write(name, system.vertex_values(kind, f[name]), t)
        system.p_old.copy_(system.u[:system.np_])      # assign(p_old, p_h)
        t += dt
    if out is not None:
        out.close()
    flux = system.flux()
    r2f, r1f, total = system.bottom_flux(flux)
    approx = 4 * math.pi * D1 * sigma * (concentration / (gamma + concentration))           # :264
    results.append([sigma, gamma, concentration, r1f, r2f, total, approx])
    solve_problem_instance_mixed.last_system = system
    return flux, total
