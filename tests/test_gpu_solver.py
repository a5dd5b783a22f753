"""GPU linear solver and Newton loop against the oracle's sparse-LU Newton (SuperLU standing in for
MUMPS/UMFPACK): Newton solutions within 1e-8 relative L2 (north-star tolerance)."""
import math

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import torch

import closestmolecule_b200 as cm
from closestmolecule_b200.solver import JacobiKrylovLinearSolver, StructuredPQLinearSolver
from oracle import fem, pq, scalar
from oracle.newton import merge_bcs, newton_solve

from util import np_forcing, oracle_mesh, oracle_params, random_state

pytestmark = pytest.mark.gpu
P1 = cm.FiniteElement("CG", "triangle", 1)


@pytest.fixture(autouse=True)
def _krylov_by_default(monkeypatch):
    """This file tests the multigrid / Jacobi preconditioned Krylov paths: 'default' (which now selects
    the direct LU whenever its factors fit, tests/test_gpu_direct.py) is mapped to 'gmres' here."""
    orig = cm.NonlinearVariationalSolver._defaults

    def defaults():
        d = orig()
        d["linear_solver"] = "gmres"
        return d
    monkeypatch.setattr(cm.NonlinearVariationalSolver, "_defaults", staticmethod(defaults))


def _supg_problem(nx, ny):
    mesh = cm.RectangleMesh(cm.Point(0, 0), cm.Point(1, 1), nx, ny)
    V = cm.FunctionSpace(mesh, cm.MixedElement([P1, P1]))
    bd = cm.MeshFunction("size_t", mesh, 1, 0)
    cm.CompiledSubDomain("near(x[0],0) && on_boundary").mark(bd, 31)   # …_SUPG.py:77
    form = cm.PQPrimalForm.manufactured(V, "supg")
    from closestmolecule_b200.manufactured import p_exact, q_exact
    bcs = [cm.DirichletBC(V.sub(0), p_exact, "on_boundary"), cm.DirichletBC(V.sub(1), q_exact, bd, 31)]
    return mesh, V, form, bcs


def _oracle_newton(mesh, form, bcs, atol, rtol, maxit=50):
    om = oracle_mesh(mesh)
    S = pq.PQSystem(om, oracle_params(form))
    dofs, vals = cm.DirichletBC.merge(bcs, 2 * mesh.num_vertices())
    dofs, vals = dofs.cpu().numpy().astype(np.int64), vals.cpu().numpy()
    f = np_forcing(form)
    return S, dofs, vals, newton_solve(lambda x, wj: S.assemble(x, wj, None, f), np.zeros(S.N), S.rowptr,
                                       S.col, dofs, vals, atol=atol, rtol=rtol, maxit=maxit,
                                       return_history=True)


@pytest.mark.parametrize("nx,ny", [(32, 32), (64, 32), (40, 24)])
def test_structured_linear_solve_matches_lu(nx, ny):
    mesh, V, form, bcs = _supg_problem(nx, ny)
    s = cm.NonlinearVariationalSolver(cm.NonlinearVariationalProblem(form, cm.Function(V), bcs))
    asm = s.asm
    un = random_state(mesh.num_vertices())
    u = torch.tensor(un, device=mesh.device)
    bc = asm._bc_tables(bcs)
    asm.assemble(u, True)
    asm.apply_bcs(bc, u, True)
    lin = StructuredPQLinearSolver(asm, restart=40)
    lin.setup()
    dx = torch.empty_like(u)
    assert lin.solve(asm.b, dx, 1e-12, 0.0, 200)
    rp, col, vals = asm.csr()
    n = u.numel()
    A = sp.csr_matrix((vals.cpu().numpy(), col.cpu().numpy(), rp.cpu().numpy()), shape=(n, n))
    ref = spla.splu(A.tocsc()).solve(asm.b.cpu().numpy())
    err = np.linalg.norm(dx.cpu().numpy() - ref) / np.linalg.norm(ref)
    assert err < 1e-9, (err, lin.last_iterations)
    assert lin.last_iterations < 60


@pytest.mark.parametrize("nx,ny", [(32, 32), (128, 64)])
def test_newton_supg_matches_oracle(nx, ny):
    mesh, V, form, bcs = _supg_problem(nx, ny)
    u = cm.Function(V)
    s = cm.NonlinearVariationalSolver(cm.NonlinearVariationalProblem(form, u, bcs))
    s.parameters["newton_solver"]["linear_solver"] = "gmres"
    s.parameters["newton_solver"]["absolute_tolerance"] = 1e-8
    s.parameters["newton_solver"]["relative_tolerance"] = 1e-8
    it, conv = s.solve()
    S, dofs, vals, (x, oit, oconv, hist) = _oracle_newton(mesh, form, bcs, 1e-8, 1e-8)
    assert conv and oconv and it == oit
    err = np.linalg.norm(u.vector().cpu().numpy() - x) / np.linalg.norm(x)
    assert err < 1e-8, err
    # same residual history as DOLFIN-style Newton with LU
    assert np.allclose(s.history, hist, rtol=1e-5, atol=1e-12)


def test_newton_nonconvergence_raises():
    mesh, V, form, bcs = _supg_problem(16, 16)
    u = cm.Function(V)
    s = cm.NonlinearVariationalSolver(cm.NonlinearVariationalProblem(form, u, bcs))
    s.parameters["newton_solver"]["relative_tolerance"] = 1e-30
    s.parameters["newton_solver"]["absolute_tolerance"] = 1e-30
    s.parameters["newton_solver"]["maximum_iterations"] = 2
    with pytest.raises(RuntimeError, match="maximum number of iterations"):
        s.solve()


def test_galerkin_q_block_reports_breakdown_or_solves():
    """Unstabilised Galerkin q-block: the reference needs pivoting LU; the structured solver
    must either converge or fail loudly -- never return garbage silently."""
    mesh = cm.RectangleMesh(cm.Point(0, 0), cm.Point(1, 1), 16, 16)
    V = cm.FunctionSpace(mesh, cm.MixedElement([P1, P1]))
    bd = cm.MeshFunction("size_t", mesh, 1, 0)
    cm.CompiledSubDomain("near(x[0],1) && on_boundary").mark(bd, 31)
    form = cm.PQPrimalForm.manufactured(V, "galerkin")
    from closestmolecule_b200.manufactured import p_exact, q_exact
    bcs = [cm.DirichletBC(V.sub(0), p_exact, "on_boundary"), cm.DirichletBC(V.sub(1), q_exact, bd, 31)]
    u = cm.Function(V)
    s = cm.NonlinearVariationalSolver(cm.NonlinearVariationalProblem(form, u, bcs))
    s.parameters["newton_solver"]["absolute_tolerance"] = 1e-7
    s.parameters["newton_solver"]["relative_tolerance"] = 1e-7
    try:
        it, conv = s.solve()
    except RuntimeError:
        return
    S, dofs, vals, (x, oit, oconv, hist) = _oracle_newton(mesh, form, bcs, 1e-7, 1e-7)
    assert np.linalg.norm(u.vector().cpu().numpy() - x) / np.linalg.norm(x) < 1e-6


def test_scalar_newton_reproduces_readme_rows():
    """GPU path for SphericalAdvectionDiffusionReaction_convergence.py: rows of README.md:109-115."""
    from closestmolecule_b200.manufactured import u_exact
    rows = {8: ("4.45e-01", "7.51e-03"), 32: ("1.09e-01", "4.29e-04")}
    for n, (e1s, e0s) in rows.items():
        mesh = cm.RectangleMesh(cm.Point(0, 0), cm.Point(1, 1), n, n)
        V = cm.FunctionSpace(mesh, "CG", 1)
        u = cm.Function(V)
        s = cm.NonlinearVariationalSolver(cm.NonlinearVariationalProblem(
            cm.ScalarADRForm.nonlinear_adr(V), u, [cm.DirichletBC(V, u_exact, "on_boundary")]))
        s.parameters["newton_solver"]["linear_solver"] = "mumps"
        s.parameters["newton_solver"]["absolute_tolerance"] = 1e-7
        s.parameters["newton_solver"]["relative_tolerance"] = 1e-7
        it, conv = s.solve()
        assert conv
        e1, e0 = scalar.errors(oracle_mesh(mesh), u.vector().cpu().numpy(), 4)
        assert f"{e1:6.2e}" == e1s and f"{e0:6.2e}" == e0s
        ref = scalar.solve_fb(n)
        assert np.linalg.norm(u.vector().cpu().numpy() - ref["u"]) / np.linalg.norm(ref["u"]) < 1e-8


def test_ported_convergence_script_prints_readme_table():
    """scripts/run_scalar_adr_convergence.py (the reference driver on the GPU
    path) reproduces README.md:109-115 digit for digit."""
    import importlib.util
    import os
    from test_oracle_golden import README_TABLE
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts",
                        "run_scalar_adr_convergence.py")
    spec = importlib.util.spec_from_file_location("sadr_conv", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rows = mod.main(7)
    assert [r.split() for r in rows] == [r.split() for r in README_TABLE]


def test_zebra_line_relaxation_building_blocks():
    """cm_stencil_line_factor / cm_stencil_zebra / cm_stencil_residual on a random diagonally dominant
    7-point stencil: after a sweep the residual vanishes on the lines solved last, and one colour solve
    equals scipy's tridiagonal solve of the same lines."""
    from closestmolecule_b200 import _capi
    nx, ny = 200, 37
    n1, n = nx + 1, (nx + 1) * (ny + 1)
    dev = torch.device("cuda")
    rng = np.random.default_rng(2)
    A = np.zeros((7, ny + 1, n1))
    offs = [(-1, -1), (0, -1), (-1, 0), (0, 0), (1, 0), (0, 1), (1, 1)]
    for k, (dx, dy) in enumerate(offs):
        v = (1.0 + rng.random((ny + 1, n1))) if k == 3 else -(0.1 + 0.05 * rng.random((ny + 1, n1)))
        if dx == -1: v[:, 0] = 0
        if dx == 1: v[:, -1] = 0
        if dy == -1: v[0, :] = 0
        if dy == 1: v[-1, :] = 0
        A[k] = v
    At = torch.tensor(A.reshape(7, n), device=dev)
    lf, iu, cu = (torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3))
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    b = torch.tensor(rng.standard_normal(n), device=dev)
    x = torch.zeros(n, dtype=torch.float64, device=dev)
    r = torch.empty_like(x)
    tmp = torch.empty_like(x)
    lib, P, st = _capi.lib(), _capi.ptr, _capi.stream_ptr()
    _capi.check(lib.cm_stencil_line_factor(nx, ny, P(At), P(lf), P(iu), P(cu), P(err), st))
    assert int(err.item()) == 0
    _capi.check(lib.cm_stencil_zebra(nx, ny, 0, P(At), P(lf), P(iu), P(cu), P(b), P(x), 1, P(tmp), st))
    # colour 0 from a zero guess = plain tridiagonal solves of the even lines
    import scipy.linalg as sla
    X = x.cpu().numpy().reshape(ny + 1, n1)
    B = b.cpu().numpy().reshape(ny + 1, n1)
    for j in (0, 2, 36):
        ab = np.zeros((3, n1))
        ab[0, 1:] = A[4, j, :-1]
        ab[1] = A[3, j]
        ab[2, :-1] = A[2, j, 1:]
        ref = sla.solve_banded((1, 1), ab, B[j])
        assert np.max(np.abs(X[j] - ref)) < 1e-13 * np.max(np.abs(ref)) + 1e-14
    _capi.check(lib.cm_stencil_zebra(nx, ny, 1, P(At), P(lf), P(iu), P(cu), P(b), P(x), 0, P(tmp), st))
    _capi.check(lib.cm_stencil_residual(nx, ny, P(At), P(b), P(x), P(r), st))
    R = r.cpu().numpy().reshape(ny + 1, n1)
    assert np.max(np.abs(R[1::2])) < 1e-13 and np.max(np.abs(R[0::2])) > 1e-3


def test_pq_supg_convergence_rates_on_gpu():
    """scripts/run_pq_supg_convergence.py: expected orders of the P1 x P1 SUPG system (2 in L2(p), 1 in
    H1(p) and H1(q)) with mesh-independent GMRES iteration counts."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts",
                        "run_pq_supg_convergence.py")
    spec = importlib.util.spec_from_file_location("pq_supg_conv", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    r = mod.main(8, linear_solver="gmres")
    assert 1.9 < r["rp"][-1] < 2.1 and 0.95 < r["r1p"][-1] < 1.05 and 0.9 < r["rq"][-1] < 1.3
    assert max(max(l) for _, l in r["its"][3:]) <= 20


@pytest.mark.parametrize("kind,n", [("minus", 32), ("minus", 64), ("paper", 32), ("other", 32)])
def test_newton_c0ip_matches_oracle(kind, n):
    """C0IP q-equation (13-point interior-facet pattern, constant facet operator) solved on the GPU:
    …MixedPrimal_convergence_C0IP.py:141-147 (minus sign, strong q data at x=1), the "paper"
    variant test_for_paper_convergence.py:140-153 (plus sign, weak data, boundary penalty) and
    …_C0IP_otherBCs.py:55-60,131-134 (r2vec = (-1,0), plus sign, D1 = 2, D2 = 1.5; direct solve)."""
    from closestmolecule_b200.manufactured import p_exact, q_exact
    from util import np_qdata, oracle_ext
    mesh = cm.RectangleMesh(cm.Point(0, 0), cm.Point(1, 1), n, n)
    V = cm.FunctionSpace(mesh, cm.MixedElement([P1, P1]))
    mf = cm.MeshFunction("size_t", mesh, 1, 0)
    cm.CompiledSubDomain("near(x[0],0) && on_boundary").mark(mf, 30)
    cm.CompiledSubDomain("near(x[0],1) && on_boundary").mark(mf, 31)
    cm.CompiledSubDomain("(near(x[1],0) || near(x[1],1)) && on_boundary").mark(mf, 32)
    if kind == "minus":
        form = cm.PQPrimalForm.manufactured(V, "c0ip", D1=0.01, D2=0.015, stab=0.05, sign=-1.0,
                                            ext_terms=[(mf, 30, cm.EF_ADV)])
        bcs = [cm.DirichletBC(V.sub(0), p_exact, "on_boundary"), cm.DirichletBC(V.sub(1), q_exact, mf, 31)]
    elif kind == "other":
        form = cm.PQPrimalForm.manufactured(V, "c0ip", D1=2.0, D2=1.5, stab=7.5e-3, sign=+1.0, r2dir=-1.0,
                                            ext_terms=[(mf, 30, cm.EF_ADV)])
        assert (form.params.q_divr, form.params.q_ibp, form.params.r2x) == (-1.0, -1.0, -1.0)
        bcs = [cm.DirichletBC(V.sub(0), p_exact, "on_boundary"), cm.DirichletBC(V.sub(1), q_exact, mf, 31)]
    else:
        form = cm.PQPrimalForm.manufactured(
            V, "c0ip", D1=1.0, D2=1.5, stab=0.05, sign=+1.0, c0ip_h="facet", degree=6, pen_stab2=1.0,
            ext_terms=[(mf, 30, cm.EF_ADV), (mf, 32, cm.EF_ADV), (mf, 31, cm.EF_DATA),
                       (mf, "on_boundary", cm.EF_PENALTY)])
        bcs = [cm.DirichletBC(V.sub(0), p_exact, "on_boundary")]
    u = cm.Function(V)
    s = cm.NonlinearVariationalSolver(cm.NonlinearVariationalProblem(form, u, bcs))
    s.parameters["newton_solver"]["absolute_tolerance"] = 1e-8
    s.parameters["newton_solver"]["relative_tolerance"] = 1e-8
    s.parameters["newton_solver"]["krylov_solver"]["restart"] = 120
    if kind == "other":
        s.parameters["newton_solver"]["linear_solver"] = "mumps"
    it, conv = s.solve()
    om = oracle_mesh(mesh)
    ef, ek = oracle_ext(form, om)
    S = pq.PQSystem(om, oracle_params(form), ef, ek)
    dofs, vals = cm.DirichletBC.merge(bcs, 2 * mesh.num_vertices())
    f, qd = np_forcing(form), np_qdata(form)
    x, oit, oconv = newton_solve(lambda xx, wj: S.assemble(xx, wj, None, f, qd), np.zeros(S.N), S.rowptr, S.col,
                                 dofs.cpu().numpy().astype(np.int64), vals.cpu().numpy(), atol=1e-8, rtol=1e-8)
    assert conv and oconv and it == oit
    err = np.linalg.norm(u.vector().cpu().numpy() - x) / np.linalg.norm(x)
    assert err < 1e-8, (err, s.linear_iterations)


def test_transient_backward_euler_steps_match_oracle():
    """Two backward-Euler steps of the transient form (advection_diffusion_convergence.py:196-199: the
    (p - p_old)/dt term, assign(p_old, p_h) after every solve :225), SUPG q-equation, warm-started Newton."""
    import math
    from closestmolecule_b200.manufactured import forcing_torch, p_exact, q_exact
    n, dt = 48, 0.1
    mesh = cm.RectangleMesh(cm.Point(0, 0), cm.Point(1, 1), n, n // 2)
    V = cm.FunctionSpace(mesh, cm.MixedElement([P1, P1]))
    bd = cm.MeshFunction("size_t", mesh, 1, 0)
    cm.CompiledSubDomain("near(x[0],0) && on_boundary").mark(bd, 31)
    p_old = cm.interpolate(lambda x, y: 0.8 * p_exact(x, y), V.sub(0).collapse())
    D1, D2 = 1e-3, 1.5e-3
    gc = D2 * 4.0 * math.pi
    form = cm.PQPrimalForm(V, D1, D2, G=("power", gc), dt=dt, p_old=p_old,
                           forcing=forcing_torch(D1, D2, gc, q_react=1.0, x2_on_q=True),
                           q_react=1.0, q_x2p=0.0, q_x2q=1.0, supg_stab=1.0)
    bcs = [cm.DirichletBC(V.sub(0), p_exact, "on_boundary"), cm.DirichletBC(V.sub(1), q_exact, bd, 31)]
    u = cm.Function(V)
    s = cm.NonlinearVariationalSolver(cm.NonlinearVariationalProblem(form, u, bcs))
    s.parameters["newton_solver"]["absolute_tolerance"] = 1e-9
    s.parameters["newton_solver"]["relative_tolerance"] = 1e-9
    om = oracle_mesh(mesh)
    S = pq.PQSystem(om, oracle_params(form))
    dofs, vals = cm.DirichletBC.merge(bcs, 2 * mesh.num_vertices())
    dofs, vals = dofs.cpu().numpy().astype(np.int64), vals.cpu().numpy()
    f = np_forcing(form)
    xo = np.zeros(S.N)
    po = p_old.vector().cpu().numpy().copy()
    for step in range(2):
        it, conv = s.solve()
        xo, oit, oconv = newton_solve(lambda xx, wj: S.assemble(xx, wj, po, f), xo, S.rowptr, S.col, dofs, vals,
                                      atol=1e-9, rtol=1e-9)
        assert conv and oconv and it == oit
        err = np.linalg.norm(u.vector().cpu().numpy() - xo) / np.linalg.norm(xo)
        assert err < 1e-8, (step, err)
        p_old.vector().copy_(u.split()[0])      # assign(p_old, p_h)
        po = xo[0::2].copy()


def test_production_constants_fail_loudly():
    """The reference's production form (unstabilised Galerkin q-equation, advection-dominated p-block)
    needs a pivoting sparse LU (tests/test_gpu_direct.py): the Krylov path must raise,
    never return garbage."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts",
                        "run_production_instance.py")
    spec = importlib.util.spec_from_file_location("prod_inst", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mesh = cm.RectangleMesh(cm.Point(0, 0.1), cm.Point(5, 5), 64, 64)
    bdry = cm.structured_boundary_markers(mesh)
    with pytest.raises(RuntimeError):
        mod.solve_problem_instance(1, 0.1, 0.1, mesh, bdry, 0.1, steady_state=True, supg_stab=0.0,
                                   linear_solver="gmres")


def test_unstructured_xml_mesh_scalar_newton(tmp_path):
    """Mesh(xml) path (advection_diffusion_mixed.py:313): a mesh read from DOLFIN XML has no structured
    index, so assembly runs through the row-owner gather kernels and the solve through Jacobi-GMRES."""
    from closestmolecule_b200.manufactured import u_exact
    src = cm.RectangleMesh(cm.Point(0, 0), cm.Point(1, 1), 32, 32)
    cm.write_mesh(src, tmp_path / "m.xml")
    mesh = cm.read_mesh(tmp_path / "m.xml")
    assert not mesh.structured
    V = cm.FunctionSpace(mesh, "CG", 1)
    u = cm.Function(V)
    s = cm.NonlinearVariationalSolver(cm.NonlinearVariationalProblem(
        cm.ScalarADRForm.nonlinear_adr(V), u, [cm.DirichletBC(V, u_exact, "on_boundary")]))
    s.parameters["newton_solver"]["absolute_tolerance"] = 1e-7
    s.parameters["newton_solver"]["relative_tolerance"] = 1e-7
    it, conv = s.solve()
    ref = scalar.solve_fb(32)
    assert conv and np.linalg.norm(u.vector().cpu().numpy() - ref["u"]) / np.linalg.norm(ref["u"]) < 1e-8
    # the P-Q solver refuses unstructured meshes loudly
    W = cm.FunctionSpace(mesh, cm.MixedElement([P1, P1]))
    with pytest.raises(RuntimeError):
        cm.NonlinearVariationalSolver(cm.NonlinearVariationalProblem(
            cm.PQPrimalForm.manufactured(W, "supg"), cm.Function(W), [])).solve()


@pytest.mark.parametrize("meshkind,nx,ny", [("mapped", 96, 40), ("flat", 100, 60), ("unit", 72, 36)])
def test_newton_supg_on_production_like_geometry(meshkind, nx, ny):
    """SUPG manufactured problem on the boundary-fitted production-like domain (curved bottom
    boundary, advection_diffusion_mixed.py geometry) and on non-power-of-two grids: the Galerkin
    hierarchy is built from the assembled operator, so mapped coordinates need no special handling."""
    from closestmolecule_b200.manufactured import p_exact, q_exact
    if meshkind == "mapped":
        mesh = cm.MappedRectangleMesh(nx, ny, 2.0, 2.0, 0.1, 1.0)
    elif meshkind == "flat":
        mesh = cm.RectangleMesh(cm.Point(0, 0.1), cm.Point(2, 2), nx, ny)
    else:
        mesh = cm.RectangleMesh(cm.Point(0, 0), cm.Point(1, 1), nx, ny)
    V = cm.FunctionSpace(mesh, cm.MixedElement([P1, P1]))
    bd = cm.structured_boundary_markers(mesh)
    form = cm.PQPrimalForm.manufactured(V, "supg", D1=1e-2, D2=1.5e-2)
    bcs = [cm.DirichletBC(V.sub(0), p_exact, "on_boundary"), cm.DirichletBC(V.sub(1), q_exact, bd, 23)]
    u = cm.Function(V)
    s = cm.NonlinearVariationalSolver(cm.NonlinearVariationalProblem(form, u, bcs))
    s.parameters["newton_solver"]["absolute_tolerance"] = 1e-8
    s.parameters["newton_solver"]["relative_tolerance"] = 1e-8
    it, conv = s.solve()
    om = oracle_mesh(mesh)
    S = pq.PQSystem(om, oracle_params(form))
    dofs, vals = cm.DirichletBC.merge(bcs, 2 * mesh.num_vertices())
    f = np_forcing(form)
    x, oit, oconv = newton_solve(lambda xx, wj: S.assemble(xx, wj, None, f), np.zeros(S.N), S.rowptr, S.col,
                                 dofs.cpu().numpy().astype(np.int64), vals.cpu().numpy(), atol=1e-8, rtol=1e-8)
    assert conv and oconv and it == oit
    err = np.linalg.norm(u.vector().cpu().numpy() - x) / np.linalg.norm(x)
    assert err < 1e-8, (err, s.linear_iterations)
    assert max(s.linear_iterations) < 80, s.linear_iterations


def test_fused_krylov_vector_kernels_match_torch():
    """cm_dot_batch / cm_axpy_batch (the GMRES orthogonalisation kernels) against torch, including group
    boundaries (nvec = 1, 8, 9, 19) and a vector length that is not a multiple of anything."""
    from closestmolecule_b200 import _capi
    lib, P, st = _capi.lib(), _capi.ptr, _capi.stream_ptr()
    dev = torch.device("cuda")
    n, stride = 100003, 100008
    g = torch.Generator(device="cpu").manual_seed(1)
    scratch = torch.empty(int(lib.cm_vec_scratch_doubles()), dtype=torch.float64, device=dev)
    for nvec in (1, 8, 9, 19):
        V = torch.randn(nvec, stride, generator=g, dtype=torch.float64).to(dev)
        w = torch.randn(n, generator=g, dtype=torch.float64).to(dev)
        out = torch.empty(nvec, dtype=torch.float64, device=dev)
        _capi.check(lib.cm_dot_batch(n, P(V), stride, nvec, P(w), P(out), P(scratch), st))
        ref = V[:, :n] @ w
        assert float((out - ref).abs().max()) < 1e-11 * float(ref.abs().max() + 1)
        h = torch.randn(nvec, generator=g, dtype=torch.float64).to(dev)
        w2 = w.clone()
        nrm = torch.empty(1, dtype=torch.float64, device=dev)
        _capi.check(lib.cm_axpy_batch(n, P(V), stride, nvec, P(h), -1.0, P(w2), P(nrm), P(scratch), st))
        wref = w - h @ V[:, :n]
        assert float((w2 - wref).abs().max()) < 1e-12 * float(wref.abs().max())
        assert abs(float(nrm) - float(wref @ wref)) < 1e-11 * float(wref @ wref)
        out2 = torch.empty_like(out)
        _capi.check(lib.cm_dot_batch(n, P(V), stride, nvec, P(w), P(out2), P(scratch), st))
        assert torch.equal(out, out2)        # bit-reproducible


def test_laplace_convergence_driver_config1():
    """BASELINE config 1 (SphericalLaplace_convergence.py) through the GPU path: DoF column (n+1)^2, expected orders
    (README.md:118: h in H1, h^2 in L2), and the same errors as the oracle's restatement of the script."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts",
                        "run_laplace_convergence.py")
    spec = importlib.util.spec_from_file_location("lap_conv", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    r = mod.main(7)
    assert r["dofs"] == [(2 ** (k + 1) + 1) ** 2 for k in range(7)]
    assert 0.95 < r["r1"][-1] < 1.05 and 1.9 < r["r0"][-1] < 2.1
    ref = scalar.solve_fa(64)
    assert abs(r["e1"][5] / ref["e1"] - 1) < 1e-2 and abs(r["e0"][5] / ref["e0"] - 1) < 5e-2


def test_bicgstab_choice_solves_the_scalar_newton_problem():
    """linear_solver = 'bicgstab' (cm_bicgstab_solve: right-preconditioned BiCGStab + Jacobi on the node CSR) on the
    scalar nonlinear problem of README.md:109-115: same Newton solution as the direct path, real iterations."""
    from closestmolecule_b200.manufactured import u_exact
    n = 32
    sols = {}
    for name in ("bicgstab", "mumps"):
        mesh = cm.RectangleMesh(cm.Point(0, 0), cm.Point(1, 1), n, n)
        V = cm.FunctionSpace(mesh, "CG", 1)
        u = cm.Function(V)
        s = cm.NonlinearVariationalSolver(cm.NonlinearVariationalProblem(
            cm.ScalarADRForm.nonlinear_adr(V), u, [cm.DirichletBC(V, u_exact, "on_boundary")]))
        s.parameters["newton_solver"]["linear_solver"] = name
        s.parameters["newton_solver"]["absolute_tolerance"] = 1e-9
        s.parameters["newton_solver"]["relative_tolerance"] = 1e-9
        it, conv = s.solve()
        assert conv
        sols[name] = u.vector().cpu().numpy()
        if name == "bicgstab":
            assert type(s._linear).__name__ == "JacobiKrylovLinearSolver" and s._linear.method == "bicgstab"
            assert all(k > 0 for k in s.linear_iterations) and s._linear.last_residual < 1e-10
    assert np.linalg.norm(sols["bicgstab"] - sols["mumps"]) / np.linalg.norm(sols["mumps"]) < 1e-8
    e1, e0 = scalar.errors(oracle_mesh(mesh), sols["bicgstab"], 4)
    assert f"{e1:6.2e}" == "1.09e-01" and f"{e0:6.2e}" == "4.29e-04"
