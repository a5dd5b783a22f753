"""Assembly front end: owns the device buffers of one form on one mesh and calls the C ABI.

Replaces the `assemble(b, FF)` / `assemble(A, Tang)` / `bc.apply` calls DOLFIN's NewtonSolver makes
[ext] when the scripts call solver.solve() (advection_diffusion_convergence.py:219).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _capi
from .quadrature import facet_rule, nq_of_degree, triangle_rule
from .space import DirichletBC


def _quad_points(mesh, degree, c0, c1):
    """x, y [n, nq] of the kernel's quadrature points for cells c0:c1."""
    pts, _ = triangle_rule(degree)
    dev = mesh.device
    phi = torch.tensor(np.stack([1.0 - pts[:, 0] - pts[:, 1], pts[:, 0], pts[:, 1]], 1), device=dev)
    X = mesh.coords[mesh.cells[c0:c1].long()]          # [n,3,2]
    P = torch.einsum("qi,cid->cqd", phi, X)
    return P[:, :, 0], P[:, :, 1]


class _Base:
    def __init__(self, form):
        self.form = form
        self.mesh = form.space.mesh
        _capi.require_cuda(self.mesh.coords)
        self.lib = _capi.lib()

    def _bc_tables(self, bcs):
        dofs, vals = DirichletBC.merge(bcs, self.form.space.dim())
        dev = self.mesh.device
        dofs, vals = dofs.to(dev), vals.to(dev)
        node = dofs.long() // self.bs
        diag = self.topo.diag_slot[node].contiguous()
        return dofs, vals, diag

    def apply_bcs(self, bc_tables, x, with_matrix=True, blocks=None):
        """blocks: the stencil form of the Jacobian the assembly wrote next to the CSR values (PQAssembler.assemble
        with blocks=...): its Dirichlet rows get the same treatment."""
        dofs, g, diag = bc_tables
        _capi.check(self.lib.cm_apply_dirichlet(
            self.bs, dofs.numel(), _capi.ptr(dofs), _capi.ptr(diag), _capi.ptr(g), _capi.ptr(x),
            _capi.ptr(self.topo.nrowptr), _capi.ptr(self.vals) if with_matrix else None,
            _capi.ptr(self.b), _capi.stream_ptr()))
        if blocks is not None and with_matrix:
            _capi.check(self.lib.cm_apply_dirichlet_blocks(dofs.numel(), _capi.ptr(dofs), self.topo.nv,
                                                           C.byref(blocks), _capi.stream_ptr()))

    def _load_vector(self, ncomp, weighted, supg_stab, degree):
        mesh, t = self.mesh, self.topo
        nq = nq_of_degree(degree)
        Cn = mesh.num_cells()
        fq = torch.empty((Cn, nq, ncomp), dtype=torch.float64, device=mesh.device)
        chunk = 1 << 21
        for c0 in range(0, Cn, chunk):
            c1 = min(Cn, c0 + chunk)
            x, y = _quad_points(mesh, degree, c0, c1)
            f = self.form.forcing(x, y)
            if ncomp == 1:
                fq[c0:c1, :, 0] = f
            else:
                fq[c0:c1, :, 0] = f[0]
                fq[c0:c1, :, 1] = f[1]
        load = torch.empty(ncomp * mesh.num_vertices(), dtype=torch.float64, device=mesh.device)
        _capi.check(self.lib.cm_assemble_load(
            ncomp, weighted, float(supg_stab), degree, mesh.num_vertices(), Cn, _capi.ptr(mesh.cells),
            _capi.ptr(mesh.coords), _capi.ptr(t.v2c_ptr), _capi.ptr(t.v2c_ent), _capi.ptr(fq),
            _capi.ptr(load), _capi.stream_ptr()))
        torch.cuda.current_stream().synchronize()
        return load

    # scalar CSR view (DOLFIN / PETSc AIJ order) for parity checks
    def csr(self):
        rp, col = self.topo.scalar_csr(self.bs)
        return rp, col, self.vals


class PQAssembler(_Base):
    bs = 2

    def __init__(self, form, use_sweep=True):
        super().__init__(form)
        self.use_sweep = use_sweep
        mesh = self.mesh
        self.topo = t = mesh.topology(form.with_interior_facets)
        dev = mesh.device
        self.vals = torch.zeros(4 * t.nnz, dtype=torch.float64, device=dev)
        self.b = torch.zeros(2 * t.nv, dtype=torch.float64, device=dev)
        self.load = None
        if form.forcing is not None:
            self.load = self._load_vector(2, 1, form.params.supg_stab, form.params.degree)
        self.Kq = self.cq = None
        if form.with_interior_facets or form.ext_terms:
            self._facet_operator()

    # -- constant facet operator of the C0IP q-equation -------------------------------------------
    def _facet_operator(self):
        mesh, t, form = self.mesh, self.topo, self.form
        dev = mesh.device
        V = t.nv
        fc = mesh.facets()
        self.Kq = torch.zeros(t.nnz, dtype=torch.float64, device=dev)
        self.cq = torch.zeros(V, dtype=torch.float64, device=dev)
        prm = C.byref(form.params)
        st = _capi.stream_ptr()
        cells = mesh.cells.long()
        if form.with_interior_facets:
            fi = fc.interior()
            fverts = fc.verts[fi].contiguous()
            fcell = torch.stack([fc.cell0[fi], fc.cell1[fi]], 1).contiguous()
            o0 = cells[fc.cell0[fi].long(), fc.loc0[fi].long()]
            o1 = cells[fc.cell1[fi].long(), fc.loc1[fi].long()]
            macro = torch.stack([fverts[:, 0].long(), fverts[:, 1].long(), o0, o1], 1)  # [Fi,4]
            ptr_, ent = _vertex_adjacency(macro, V)
            _capi.check(self.lib.cm_assemble_c0ip_facets(
                prm, V, fi.numel(), _capi.ptr(mesh.cells), _capi.ptr(mesh.coords), _capi.ptr(fverts),
                _capi.ptr(fcell), _capi.ptr(ptr_), _capi.ptr(ent), _capi.ptr(t.nrowptr),
                _capi.ptr(t.ncol), _capi.ptr(self.Kq), st))
        if form.ext_terms:
            kind = torch.zeros(fc.num, dtype=torch.int32, device=dev)
            ext = fc.cell1 < 0
            for mf, marker, bits in form.ext_terms:
                sel = ext if marker == "on_boundary" else (ext & (mf.values == int(marker)))
                kind[sel] |= int(bits)
            fe = torch.nonzero(kind != 0).flatten()
            efverts = fc.verts[fe].contiguous()
            efopp = cells[fc.cell0[fe].long(), fc.loc0[fe].long()].to(torch.int32).contiguous()
            efkind = kind[fe].contiguous()
            s, _ = facet_rule(form.params.degree)
            s = torch.tensor(s, device=dev)
            A = mesh.coords[efverts[:, 0].long()]
            B = mesh.coords[efverts[:, 1].long()]
            P = A[:, None, :] + s[None, :, None] * (B - A)[:, None, :]
            qdata = None
            if form.q_data is not None:
                qdata = form.q_data(P[:, :, 0], P[:, :, 1]).to(torch.float64).contiguous()
            ptr_, ent = _vertex_adjacency(efverts.long(), V)
            _capi.check(self.lib.cm_assemble_exterior_facets(
                prm, V, fe.numel(), _capi.ptr(mesh.coords), _capi.ptr(efverts), _capi.ptr(efopp),
                _capi.ptr(efkind), _capi.ptr(qdata), _capi.ptr(ptr_), _capi.ptr(ent),
                _capi.ptr(t.nrowptr), _capi.ptr(t.ncol), _capi.ptr(self.Kq), _capi.ptr(self.cq), st))
        torch.cuda.current_stream().synchronize()

    def writes_blocks(self):
        """True when assemble(..., blocks=...) can hand the linear solver its stencil blocks directly: structured
        sweep, Jacobian = cell integrals only (7-point pattern)."""
        mesh, t = self.mesh, self.topo
        return bool(self.use_sweep and mesh.structured and self.Kq is None and not t.with_interior_facets)

    def assemble(self, u, want_jac=True, blocks=None):
        """Fused J (self.vals, DOLFIN CSR order) + F (self.b) at state u [2V].  blocks (a _capi.PQBlocksC of a
        structured solver context): also write the stencil form of J and the row scalings there."""
        mesh, t, form = self.mesh, self.topo, self.form
        p_old = form.p_old.vector() if form.p_old is not None else None
        if self.writes_blocks() and blocks is not None and want_jac:
            _capi.check(self.lib.cm_assemble_pq_cells_structured_blocks(
                C.byref(form.params), mesh.nx, mesh.ny, _capi.ptr(mesh.coords), _capi.ptr(t.nrowptr),
                _capi.ptr(u), _capi.ptr(p_old), _capi.ptr(self.load), _capi.ptr(self.vals), _capi.ptr(self.b),
                C.byref(blocks), 0, -1, _capi.stream_ptr()))
            return self.vals, self.b
        if self.use_sweep and mesh.structured and self.Kq is None and not t.with_interior_facets:
            # structured fast path: no connectivity / gather tables are read
            _capi.check(self.lib.cm_assemble_pq_cells_structured(
                C.byref(form.params), mesh.nx, mesh.ny, _capi.ptr(mesh.coords), _capi.ptr(t.nrowptr),
                _capi.ptr(u), _capi.ptr(p_old), _capi.ptr(self.load),
                _capi.ptr(self.vals) if want_jac else None, _capi.ptr(self.b), _capi.stream_ptr()))
            return (self.vals if want_jac else None), self.b
        _capi.check(self.lib.cm_assemble_pq_cells(
            C.byref(form.params), t.nv, mesh.num_cells(), _capi.ptr(mesh.cells), _capi.ptr(mesh.coords),
            _capi.ptr(t.v2c_ptr), _capi.ptr(t.v2c_ent), _capi.ptr(t.v2c_slot), _capi.ptr(t.nrowptr),
            _capi.ptr(t.ncol), t.maxdeg, _capi.ptr(u), _capi.ptr(p_old), _capi.ptr(self.load),
            _capi.ptr(self.Kq), _capi.ptr(self.cq), _capi.ptr(self.vals) if want_jac else None,
            _capi.ptr(self.b), 0, _capi.stream_ptr()))
        return (self.vals if want_jac else None), self.b


class ScalarAssembler(_Base):
    bs = 1

    def __init__(self, form):
        super().__init__(form)
        mesh = self.mesh
        self.topo = t = mesh.topology(False)
        dev = mesh.device
        self.vals = torch.zeros(t.nnz, dtype=torch.float64, device=dev)
        self.b = torch.zeros(t.nv, dtype=torch.float64, device=dev)
        self.load = None
        if form.forcing is not None:
            self.load = self._load_vector(1, 0, 0.0, form.load_degree)

    def assemble(self, u, want_jac=True):
        mesh, t, form = self.mesh, self.topo, self.form
        _capi.check(self.lib.cm_assemble_scalar_cells(
            C.byref(form.params), t.nv, mesh.num_cells(), _capi.ptr(mesh.cells), _capi.ptr(mesh.coords),
            _capi.ptr(t.v2c_ptr), _capi.ptr(t.v2c_ent), _capi.ptr(t.v2c_slot), _capi.ptr(t.nrowptr),
            t.maxdeg, _capi.ptr(u), _capi.ptr(self.load), _capi.ptr(self.vals) if want_jac else None,
            _capi.ptr(self.b), _capi.stream_ptr()))
        return (self.vals if want_jac else None), self.b


def _vertex_adjacency(ent_verts, V):
    """CSR vertex -> entity index for entity vertex lists [n,k] (ascending entity index per vertex)."""
    n, k = ent_verts.shape
    flat = ent_verts.flatten()
    _, order = torch.sort(flat, stable=True)
    ent = (order // k).to(torch.int32).contiguous()
    cnt = torch.bincount(flat, minlength=V)
    ptr_ = torch.zeros(V + 1, dtype=torch.int64, device=flat.device)
    ptr_[1:] = torch.cumsum(cnt, 0)
    return ptr_.to(torch.int32).contiguous(), ent


def make_assembler(form):
    from .forms import PQPrimalForm
    return PQAssembler(form) if isinstance(form, PQPrimalForm) else ScalarAssembler(form)
