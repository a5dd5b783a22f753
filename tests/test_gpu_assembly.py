"""CUDA assembly (through the C ABI) against the oracle: bit-exact sparsity, values <= 1e-12 relative
to the row inf-norm, two runs bit-identical."""
import math

import numpy as np
import pytest
import torch

import closestmolecule_b200 as cm
from closestmolecule_b200.assembler import PQAssembler, ScalarAssembler
from oracle import fem, pq, scalar
from oracle.newton import apply_bc_matrix_vec, apply_bc_residual

from util import (np_forcing, np_qdata, oracle_ext, oracle_mesh, oracle_params, random_state,
                  rel_row_err)

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _mesh(kind, nx, ny):
    if kind == "unit":
        return cm.RectangleMesh(cm.Point(0, 0), cm.Point(1, 1), nx, ny)
    if kind == "flat":
        return cm.RectangleMesh(cm.Point(0, 0.1), cm.Point(5, 5), nx, ny)
    return cm.MappedRectangleMesh(nx, ny, 5.0, 5.0, 0.1, 1.0)


def _forms(V, mesh):
    mf = cm.MeshFunction("size_t", mesh, 1, 0)
    cm.CompiledSubDomain("near(x[0],0) && on_boundary").mark(mf, 30)
    cm.CompiledSubDomain("near(x[0],1) && on_boundary").mark(mf, 31)
    cm.CompiledSubDomain("(near(x[1],0) || near(x[1],1)) && on_boundary").mark(mf, 32)
    p_old = cm.Function(V.sub(0).collapse()).interpolate(lambda x, y: 0.2 + 0.1 * x * y)
    return {
        "production_steady": cm.PQPrimalForm.flat_boundary(V, 1.0, 2.0, 1.5),
        "production_transient": cm.PQPrimalForm.flat_boundary(V, 1.0, 2.0, 5.5, dt=0.1, p_old=p_old,
                                                              production_gstar=True),
        "galerkin": cm.PQPrimalForm.manufactured(V, "galerkin"),
        "supg": cm.PQPrimalForm.manufactured(V, "supg"),
        "c0ip_minus": cm.PQPrimalForm.manufactured(
            V, "c0ip", D1=0.01, D2=0.015, ext_terms=[(mf, 30, cm.EF_ADV)]),
        "c0ip_paper": cm.PQPrimalForm.manufactured(
            V, "c0ip", D1=1.0, D2=1.5, stab=0.05, sign=+1.0, c0ip_h="facet", degree=6, pen_stab2=1.0,
            ext_terms=[(mf, 30, cm.EF_ADV), (mf, 32, cm.EF_ADV), (mf, 31, cm.EF_DATA),
                       (mf, "on_boundary", cm.EF_PENALTY)]),
    }


@pytest.mark.parametrize("name", ["production_steady", "production_transient", "galerkin", "supg",
                                  "c0ip_minus", "c0ip_paper"])
@pytest.mark.parametrize("meshkind,nx,ny", [("unit", 9, 7), ("unit", 40, 24)])
def test_pq_jacobian_residual_parity(name, meshkind, nx, ny):
    mesh = _mesh(meshkind, nx, ny)
    V = cm.FunctionSpace(mesh, cm.MixedElement([cm.FiniteElement("CG", "triangle", 1)] * 2))
    form = _forms(V, mesh)[name]
    asm = PQAssembler(form)
    un = random_state(mesh.num_vertices())
    u = torch.tensor(un, device=mesh.device)
    vals, b = asm.assemble(u, True)
    vals1, b1 = vals.clone(), b.clone()
    # oracle
    om = oracle_mesh(mesh)
    prm = oracle_params(form)
    ef, ek = oracle_ext(form, om)
    S = pq.PQSystem(om, prm, ef, ek)
    p_old = form.p_old.vector().cpu().numpy() if form.p_old is not None else None
    ovals, ob = S.assemble(un, True, p_old, np_forcing(form), np_qdata(form))
    rp, col, _ = asm.csr()
    assert np.array_equal(rp.cpu().numpy(), S.rowptr.astype(np.int64))       # bit-exact sparsity
    assert np.array_equal(col.cpu().numpy(), S.col.astype(np.int64))
    assert rel_row_err(vals.cpu().numpy(), ovals, S.rowptr) < TOL
    assert np.max(np.abs(b.cpu().numpy() - ob)) / np.max(np.abs(ob)) < TOL
    # residual-only path and determinism
    _, b2 = asm.assemble(u, False)
    assert np.max(np.abs(b2.cpu().numpy() - ob)) / np.max(np.abs(ob)) < TOL
    vals3, b3 = asm.assemble(u, True)
    assert torch.equal(vals3, vals1) and torch.equal(b3, b1)


@pytest.mark.parametrize("meshkind", ["flat", "mapped"])
def test_pq_production_geometry_with_bcs(meshkind):
    """F-C on the production-like domain with the BC list of advection_diffusion_convergence.py:174-183,
    state u = 0 (the Newton start): exercises the eager conditional at p = q = 0."""
    nx, ny = 25, 20
    mesh = _mesh(meshkind, nx, ny)
    V = cm.FunctionSpace(mesh, cm.MixedElement([cm.FiniteElement("CG", "triangle", 1)] * 2))
    form = cm.PQPrimalForm.flat_boundary(V, 1.0, 2.0, 1.5)
    asm = PQAssembler(form)
    bd = cm.structured_boundary_markers(mesh)
    sigma, c, V1 = 0.1, 1.0, 4 * math.pi * (5 - 0.1) ** 3 / 3
    pin = lambda x, y: c / V1 * torch.exp(-4 / 3 * math.pi * c * x ** 3) * (1 - sigma / y)
    qin = lambda x, y: 1 / (4 * math.pi * V1) * torch.exp(-4 / 3 * math.pi * c * x ** 3)
    bcs = [cm.DirichletBC(V.sub(0), pin, bd, 21), cm.DirichletBC(V.sub(0), pin, bd, 22),
           cm.DirichletBC(V.sub(0), cm.Constant(0.0), bd, 24), cm.DirichletBC(V.sub(1), qin, bd, 21)]
    tables = asm._bc_tables(bcs)
    for state in ("zero", "random"):
        un = np.zeros(2 * mesh.num_vertices()) if state == "zero" else 1e-4 * random_state(mesh.num_vertices())
        u = torch.tensor(un, device=mesh.device)
        vals, b = asm.assemble(u, True)
        asm.apply_bcs(tables, u, True)
        om = oracle_mesh(mesh)
        S = pq.PQSystem(om, oracle_params(form))
        ovals, ob = S.assemble(un, True)
        dofs, g, _ = tables
        dn, gn = dofs.cpu().numpy().astype(np.int64), g.cpu().numpy()
        apply_bc_matrix_vec(S.rowptr, S.col, ovals, dn)
        apply_bc_residual(ob, un, dn, gn)
        assert np.all(np.isfinite(vals.cpu().numpy()))
        assert rel_row_err(vals.cpu().numpy(), ovals, S.rowptr) < TOL
        assert np.max(np.abs(b.cpu().numpy() - ob)) / np.max(np.abs(ob)) < TOL
        # BC rows exactly identity / x-g
        v = vals.cpu().numpy()
        for d in dn[:: max(1, dn.size // 20)]:
            row = v[S.rowptr[d]:S.rowptr[d + 1]]
            assert row.sum() == 1.0 and np.count_nonzero(row) == 1


@pytest.mark.parametrize("which", ["fb", "fa"])
def test_scalar_forms_parity(which):
    n = 24
    mesh = cm.RectangleMesh(cm.Point(0, 0), cm.Point(1, 1), n, n)
    V = cm.FunctionSpace(mesh, "CG", 1)
    form = cm.ScalarADRForm.nonlinear_adr(V) if which == "fb" else cm.ScalarADRForm.laplace(V)
    asm = ScalarAssembler(form)
    rng = np.random.default_rng(1)
    un = rng.random(mesh.num_vertices())
    u = torch.tensor(un, device=mesh.device)
    vals, b = asm.assemble(u, True)
    om = oracle_mesh(mesh)
    S = scalar.ScalarP1System(om)
    if which == "fb":
        Fe, Je = scalar.element_fb(om, un, 1e-3, 1e-4, 4, True)
    else:
        Je, be = scalar.element_fa(om, 1e-4, 2, 6)
        Fe = np.einsum("cij,cj->ci", Je, un[om.cells]) - be
    ovals, ob = S.assemble(Fe, Je)
    assert np.array_equal(asm.topo.nrowptr.cpu().numpy(), S.rowptr)
    assert np.array_equal(asm.topo.ncol.cpu().numpy(), S.col)
    assert rel_row_err(vals.cpu().numpy(), ovals, S.rowptr) < TOL
    assert np.max(np.abs(b.cpu().numpy() - ob)) / np.max(np.abs(ob)) < 1e-11


def test_spmv_matches_scipy():
    import scipy.sparse as sp
    from closestmolecule_b200 import _capi
    mesh = _mesh("unit", 33, 17)
    V = cm.FunctionSpace(mesh, cm.MixedElement([cm.FiniteElement("CG", "triangle", 1)] * 2))
    asm = PQAssembler(cm.PQPrimalForm.manufactured(V, "supg"))
    u = torch.tensor(random_state(mesh.num_vertices()), device=mesh.device)
    vals, _ = asm.assemble(u, True)
    rp, col, _ = asm.csr()
    n = 2 * mesh.num_vertices()
    A = sp.csr_matrix((vals.cpu().numpy(), col.cpu().numpy(), rp.cpu().numpy()), shape=(n, n))
    x = torch.tensor(np.random.default_rng(3).standard_normal(n), device=mesh.device)
    y = torch.empty_like(x)
    t = asm.topo
    _capi.check(_capi.lib().cm_spmv(2, t.nv, _capi.ptr(t.nrowptr), _capi.ptr(t.ncol), _capi.ptr(vals),
                                    _capi.ptr(x), _capi.ptr(y), _capi.stream_ptr()))
    ref = A @ x.cpu().numpy()
    assert np.max(np.abs(y.cpu().numpy() - ref)) / np.max(np.abs(ref)) < 1e-13


@pytest.mark.parametrize("name", ["production_steady", "production_transient", "galerkin", "supg"])
@pytest.mark.parametrize("meshkind,nx,ny", [("unit", 9, 7), ("unit", 300, 70), ("mapped", 130, 33)])
def test_structured_sweep_matches_gather_and_oracle(name, meshkind, nx, ny):
    """The structured sweep kernel (cells computed once, shared-memory row accumulators) against the
    row-owner gather kernel (same summation order -> expected bit-identical or 1 ulp) and the oracle."""
    mesh = _mesh(meshkind, nx, ny)
    V = cm.FunctionSpace(mesh, cm.MixedElement([cm.FiniteElement("CG", "triangle", 1)] * 2))
    form = _forms(V, mesh)[name]
    a_sweep = PQAssembler(form, use_sweep=True)
    a_gather = PQAssembler(form, use_sweep=False)
    un = random_state(mesh.num_vertices(), seed=4)
    u = torch.tensor(un, device=mesh.device)
    for want_jac in (True, False):
        v1, b1 = a_sweep.assemble(u, want_jac)
        v2, b2 = a_gather.assemble(u, want_jac)
        scale = b2.abs().max()
        assert (b1 - b2).abs().max() / scale < 1e-13
        if want_jac:
            rp = a_sweep.topo.scalar_csr(2)[0].cpu().numpy()
            assert rel_row_err(v1.cpu().numpy(), v2.cpu().numpy(), rp) < 1e-13
    if nx * ny < 2000:
        om = oracle_mesh(mesh)
        S = pq.PQSystem(om, oracle_params(form))
        p_old = form.p_old.vector().cpu().numpy() if form.p_old is not None else None
        ovals, ob = S.assemble(un, True, p_old, np_forcing(form))
        v1, b1 = a_sweep.assemble(u, True)
        assert rel_row_err(v1.cpu().numpy(), ovals, S.rowptr) < TOL
        assert np.max(np.abs(b1.cpu().numpy() - ob)) / np.max(np.abs(ob)) < TOL
    v3, b3 = a_sweep.assemble(u, True)
    v4, b4 = a_sweep.assemble(u, True)
    assert torch.equal(v3, v4) and torch.equal(b3, b4)


@pytest.mark.parametrize("with_facets", [False, True])
def test_library_sparsity_matches_host_pattern(with_facets, monkeypatch):
    """a9: cm_build_p1_sparsity (the path CUDA meshes take) against the torch pattern of the same mesh on the CPU and
    against the oracle's DOLFIN-pattern restatement, on a structured, a mapped and a renumbered mesh."""
    from oracle import fem
    import closestmolecule_b200.mesh as cmmesh
    monkeypatch.setattr(cmmesh, "LIBRARY_SPARSITY", True)
    g = torch.Generator().manual_seed(5)
    for mesh in (_mesh("unit", 37, 11), _mesh("mapped", 20, 33)):
        for shuffle in (False, True):
            m = mesh
            if shuffle:
                V = mesh.num_vertices()
                new = torch.randperm(V, generator=g).to(mesh.device)
                coords = torch.empty_like(mesh.coords)
                coords[new] = mesh.coords
                cells, _ = torch.sort(new[mesh.cells.long()], dim=1)
                m = cm.Mesh(coords, cells.to(torch.int32), device=mesh.device)
            t = m.topology(with_facets)
            host = cm.Mesh(m.coords.cpu(), m.cells.cpu(), device="cpu").topology(with_facets)
            assert torch.equal(t.nrowptr.cpu(), host.nrowptr) and torch.equal(t.ncol.cpu(), host.ncol)
            om = fem.Mesh(m.coords.cpu().numpy(), m.cells.cpu().numpy())
            rp, col = fem.p1_node_pattern(om, with_facets)
            assert np.array_equal(t.nrowptr.cpu().numpy(), rp) and np.array_equal(t.ncol.cpu().numpy(), col)
