"""Direct Newton linear solve on the GPU (cm_blu_*: banded LU with partial pivoting) -- the drop-in for
PETScLUSolver('mumps'|'umfpack') the reference configures at advection_diffusion_convergence.py:212,
advection_diffusion_mixed.py:223 and …MixedPrimal_convergence.py:136 -- against the oracle's SuperLU
Newton: linear solutions to 1e-10, Newton iterates within 1e-8 relative L2 (north-star tolerance) on
the systems the Krylov path cannot handle (unstabilised Galerkin q-equation = BASELINE config 2,
production constants = config 4, unstructured numbering)."""
import importlib.util
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import torch

import closestmolecule_b200 as cm
from closestmolecule_b200.manufactured import p_exact, q_exact
from closestmolecule_b200.solver import BandedLULinearSolver, LinearSolveError
from oracle import pq, scalar
from oracle.newton import newton_solve

from util import np_forcing, np_qdata, oracle_ext, oracle_mesh, oracle_params, random_state

pytestmark = pytest.mark.gpu
P1 = cm.FiniteElement("CG", "triangle", 1)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _galerkin_problem(mesh):
    """…MixedPrimal_convergence.py:74-131: p on the whole boundary, q on x = 1."""
    V = cm.FunctionSpace(mesh, cm.MixedElement([P1, P1]))
    bd = cm.MeshFunction("size_t", mesh, 1, 0)
    cm.CompiledSubDomain("near(x[0],1) && on_boundary").mark(bd, 31)
    form = cm.PQPrimalForm.manufactured(V, "galerkin")
    bcs = [cm.DirichletBC(V.sub(0), p_exact, "on_boundary"), cm.DirichletBC(V.sub(1), q_exact, bd, 31)]
    return V, form, bcs


def _oracle(mesh, form, bcs, atol, rtol, x0=None, ext=False):
    om = oracle_mesh(mesh)
    if ext:
        ef, ek = oracle_ext(form, om)
        S = pq.PQSystem(om, oracle_params(form), ef, ek)
    else:
        S = pq.PQSystem(om, oracle_params(form))
    dofs, vals = cm.DirichletBC.merge(bcs, 2 * mesh.num_vertices())
    f, qd = np_forcing(form), (np_qdata(form) if ext else None)
    x0 = np.zeros(S.N) if x0 is None else x0
    return newton_solve(lambda x, wj: S.assemble(x, wj, None, f, qd), x0, S.rowptr, S.col,
                        dofs.cpu().numpy().astype(np.int64), vals.cpu().numpy(), atol=atol, rtol=rtol)


def _shuffled(mesh, seed=3):
    """Same triangulation with randomly renumbered vertices (an 'unstructured' mesh: no (ix,iy) index)."""
    g = torch.Generator().manual_seed(seed)
    V = mesh.num_vertices()
    new = torch.randperm(V, generator=g).to(mesh.device)          # new[old]
    coords = torch.empty_like(mesh.coords)
    coords[new] = mesh.coords
    cells, _ = torch.sort(new[mesh.cells.long()], dim=1)
    return cm.Mesh(coords, cells.to(torch.int32), device=mesh.device)


@pytest.mark.parametrize("nx,ny", [(24, 24), (40, 17), (17, 40), (64, 64)])
def test_direct_solve_matches_superlu_on_galerkin_jacobian(nx, ny):
    """One linear solve with the zero-q-q-diagonal Jacobian; no refinement: plain GEPP accuracy."""
    mesh = cm.RectangleMesh(cm.Point(0, 0), cm.Point(1, 1), nx, ny)
    V, form, bcs = _galerkin_problem(mesh)
    s = cm.NonlinearVariationalSolver(cm.NonlinearVariationalProblem(form, cm.Function(V), bcs))
    asm = s.asm
    u = torch.tensor(random_state(mesh.num_vertices()), device=mesh.device)
    bc = asm._bc_tables(bcs)
    asm.assemble(u, True)
    asm.apply_bcs(bc, u, True)
    rp, col, vals = asm.csr()
    n = u.numel()
    A = sp.csr_matrix((vals.cpu().numpy(), col.cpu().numpy(), rp.cpu().numpy()), shape=(n, n))
    assert np.abs(A.diagonal()[1::2]).min() < 1e-6           # q-q diagonal ~ 0: no-pivot / Jacobi methods break down
    ref = spla.splu(A.tocsc()).solve(asm.b.cpu().numpy())
    lin = BandedLULinearSolver(asm, refinement_steps=0)
    assert lin.nbn >= min(nx, ny) + 2
    lin.setup()
    dx = torch.empty_like(u)
    assert lin.solve(asm.b, dx)
    # both factorisations carry an error of cond(A)*eps (cond ~ 1e6 here): compare at that level and
    # judge the solve itself by its residual
    err = np.linalg.norm(dx.cpu().numpy() - ref) / np.linalg.norm(ref)
    assert err < 1e-7 and lin.last_residual < 1e-12, (err, lin.last_residual)
    lin.refinement_steps = 1
    lin.solve(asm.b, dx)
    err = np.linalg.norm(dx.cpu().numpy() - ref) / np.linalg.norm(ref)
    assert err < 1e-7 and lin.last_residual < 1e-14, (err, lin.last_residual)


@pytest.mark.parametrize("n", [16, 48])
def test_newton_galerkin_matches_oracle(n):
    """BASELINE config 2: the nonlinear P-Q Newton solve of the unstabilised form with 'umfpack'."""
    mesh = cm.RectangleMesh(cm.Point(0, 0), cm.Point(1, 1), n, n)
    V, form, bcs = _galerkin_problem(mesh)
    u = cm.Function(V)
    s = cm.NonlinearVariationalSolver(cm.NonlinearVariationalProblem(form, u, bcs))
    s.parameters["newton_solver"].update(linear_solver="umfpack", absolute_tolerance=1e-7,
                                         relative_tolerance=1e-7)            # :132-138
    it, conv = s.solve()
    assert isinstance(s._linear, BandedLULinearSolver)
    x, oit, oconv = _oracle(mesh, form, bcs, 1e-7, 1e-7)
    assert conv and oconv and it == oit
    err = np.linalg.norm(u.vector().cpu().numpy() - x) / np.linalg.norm(x)
    assert err < 1e-8, err


@pytest.mark.parametrize("kind", ["flat", "curved"])
def test_production_instance_reference_form_matches_oracle(kind):
    """BASELINE config 4: solve_problem_instance with the reference's own (unstabilised) form, Gstar
    nonlinearity, production constants D1=2, D2=1.5 and 'mumps' (advection_diffusion_convergence.py:105-228)."""
    import math
    spec = importlib.util.spec_from_file_location("prod_inst", os.path.join(ROOT, "scripts",
                                                                            "run_production_instance.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sigma, gamma, n = 0.1, 1.0, 40
    if kind == "flat":
        mesh, sfn = cm.RectangleMesh(cm.Point(0, sigma), cm.Point(5, 5), n, n), None
    else:
        mesh = cm.MappedRectangleMesh(n, n, 5.0, 5.0, sigma, gamma)
        sfn = lambda x: sigma * torch.exp(-4 / 3 * math.pi * gamma * x ** 3)
    bdry = cm.structured_boundary_markers(mesh)
    u, log = mod.solve_problem_instance(1, 0.1, 0.1, mesh, bdry, sigma, steady_state=True, sigma_fn=sfn)   # c = np.array([1]) at :396
    s = mod.solve_problem_instance.last_solver
    assert isinstance(s._linear, BandedLULinearSolver)
    x, oit, oconv = _oracle(mesh, s.problem.form, s.problem.bcs, 1e-9, 1e-9)
    assert oconv and log[0][0] == oit
    err = np.linalg.norm(u.vector().cpu().numpy() - x) / np.linalg.norm(x)
    assert err < 1e-8, err


def test_unstructured_numbering_pq_newton():
    """P-Q Newton on a mesh without structured index (gather assembly + reverse Cuthill-McKee + LU):
    the case the structured multigrid solver refuses."""
    base = cm.RectangleMesh(cm.Point(0, 0), cm.Point(1, 1), 28, 20)
    mesh = _shuffled(base)
    assert not mesh.structured
    V, form, bcs = _galerkin_problem(mesh)
    u = cm.Function(V)
    s = cm.NonlinearVariationalSolver(cm.NonlinearVariationalProblem(form, u, bcs))
    s.parameters["newton_solver"].update(linear_solver="mumps", absolute_tolerance=1e-8,
                                         relative_tolerance=1e-8)
    it, conv = s.solve()
    assert s._linear.nbn < 80          # RCM found a band close to the 22 nodes of the structured one
    x, oit, oconv = _oracle(mesh, form, bcs, 1e-8, 1e-8)
    assert conv and oconv and it == oit
    err = np.linalg.norm(u.vector().cpu().numpy() - x) / np.linalg.norm(x)
    assert err < 1e-8, err


def test_scalar_default_lu_reproduces_readme_row():
    """bs = 1: SphericalAdvectionDiffusionReaction_convergence.py with its 'mumps' (README.md:111)."""
    from closestmolecule_b200.manufactured import u_exact
    mesh = cm.RectangleMesh(cm.Point(0, 0), cm.Point(1, 1), 16, 16)
    V = cm.FunctionSpace(mesh, "CG", 1)
    u = cm.Function(V)
    s = cm.NonlinearVariationalSolver(cm.NonlinearVariationalProblem(
        cm.ScalarADRForm.nonlinear_adr(V), u, [cm.DirichletBC(V, u_exact, "on_boundary")]))
    s.parameters["newton_solver"].update(linear_solver="mumps", absolute_tolerance=1e-7,
                                         relative_tolerance=1e-7)
    it, conv = s.solve()
    assert conv and isinstance(s._linear, BandedLULinearSolver) and s._linear.bs == 1
    ref = scalar.solve_fb(16)
    assert np.linalg.norm(u.vector().cpu().numpy() - ref["u"]) / np.linalg.norm(ref["u"]) < 1e-9


def test_c0ip_paper_variant_direct():
    """13-point (interior-facet) pattern: the band is twice as wide; test_for_paper_convergence.py:140-153."""
    n = 24
    mesh = cm.RectangleMesh(cm.Point(0, 0), cm.Point(1, 1), n, n)
    V = cm.FunctionSpace(mesh, cm.MixedElement([P1, P1]))
    mf = cm.MeshFunction("size_t", mesh, 1, 0)
    cm.CompiledSubDomain("near(x[0],0) && on_boundary").mark(mf, 30)
    cm.CompiledSubDomain("near(x[0],1) && on_boundary").mark(mf, 31)
    cm.CompiledSubDomain("(near(x[1],0) || near(x[1],1)) && on_boundary").mark(mf, 32)
    form = cm.PQPrimalForm.manufactured(
        V, "c0ip", D1=1.0, D2=1.5, stab=0.05, sign=+1.0, c0ip_h="facet", degree=6, pen_stab2=1.0,
        ext_terms=[(mf, 30, cm.EF_ADV), (mf, 32, cm.EF_ADV), (mf, 31, cm.EF_DATA),
                   (mf, "on_boundary", cm.EF_PENALTY)])
    bcs = [cm.DirichletBC(V.sub(0), p_exact, "on_boundary")]
    u = cm.Function(V)
    s = cm.NonlinearVariationalSolver(cm.NonlinearVariationalProblem(form, u, bcs))
    s.parameters["newton_solver"].update(linear_solver="mumps", absolute_tolerance=1e-8,
                                         relative_tolerance=1e-8)
    it, conv = s.solve()
    assert s._linear.nbn >= 2 * n
    x, oit, oconv = _oracle(mesh, form, bcs, 1e-8, 1e-8, ext=True)
    assert conv and oconv and it == oit
    assert np.linalg.norm(u.vector().cpu().numpy() - x) / np.linalg.norm(x) < 1e-8


def test_singular_matrix_and_memory_budget_fail_loudly():
    mesh = cm.RectangleMesh(cm.Point(0, 0), cm.Point(1, 1), 12, 12)
    V, form, bcs = _galerkin_problem(mesh)
    s = cm.NonlinearVariationalSolver(cm.NonlinearVariationalProblem(form, cm.Function(V), bcs))
    lin = BandedLULinearSolver(s.asm)
    s.asm.vals.zero_()
    with pytest.raises(RuntimeError, match="singular"):
        lin.setup()
    with pytest.raises(LinearSolveError, match="budget"):
        BandedLULinearSolver(s.asm, max_bytes=1 << 16)
    # the caller asked for an exact solve: when the factors do not fit, the Newton front end refuses ...
    s.parameters["newton_solver"]["lu_solver"]["max_bytes"] = 1 << 16
    with pytest.raises(LinearSolveError, match="allow_krylov_fallback"):
        s._prepare()
    # ... unless the Krylov path is allowed explicitly, and then it says so
    s.parameters["newton_solver"]["lu_solver"]["allow_krylov_fallback"] = True
    with pytest.warns(UserWarning, match="falls back to the Krylov path"):
        prm, lin2 = s._prepare()
    assert not isinstance(lin2, BandedLULinearSolver) and "budget" in s.lu_fallback_reason
    # the solver cache follows the parameters: asking for gmres afterwards builds the Krylov solver
    s.parameters["newton_solver"]["linear_solver"] = "gmres"
    prm, lin3 = s._prepare()
    assert lin3 is not lin2
