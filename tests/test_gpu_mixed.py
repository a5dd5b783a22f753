"""SURVEY §8f row 1 on the GPU: the RT1 x DG0 x CG1 'paper' system (test_for_paper_convergence.py, deg=0).
Element tensors + deterministic gather vs the oracle's assembly, Newton solution vs the oracle's, and the
recorded k=0 table (:200-207) reproduced through the GPU path."""
import importlib.util
import os

import numpy as np
import pytest
import scipy.sparse as sp
import torch

import closestmolecule_b200 as cm
from closestmolecule_b200.mixed import PaperMixedSystem
from oracle import paper_m0

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _system(n):
    mesh = cm.UnitSquareMesh(n, n)
    bd = cm.MeshFunction("size_t", mesh, 1, 0)
    cm.CompiledSubDomain("near(x[0],0) && on_boundary").mark(bd, 30)
    cm.CompiledSubDomain("near(x[0],1) && on_boundary").mark(bd, 31)
    cm.CompiledSubDomain("(near(x[1],0) || near(x[1],1)) && on_boundary").mark(bd, 32)
    return PaperMixedSystem(mesh, bd)


@pytest.mark.parametrize("n", [4, 12])
def test_assembly_matches_oracle(n):
    S, O = _system(n), paper_m0.PaperM0(n)
    rng = np.random.default_rng(1)
    un = np.concatenate([rng.standard_normal(O.E), 0.5 + rng.random(O.C), 1.0 + rng.random(O.V)])
    u = torch.tensor(un, device=S.mesh.device)
    vals, b = S.assemble(u, True)
    J, bo = O.assemble(un, True)
    got = sp.csr_matrix((vals.cpu().numpy(), S.col.cpu().numpy(), S.rowptr.cpu().numpy()), shape=(S.N, S.N))
    scale = abs(J).max()
    assert abs(got - J).max() < 1e-12 * scale, abs(got - J).max() / scale
    assert np.max(np.abs(b.cpu().numpy() - bo)) < 1e-12 * np.max(np.abs(bo))
    # residual-only pass gives the same vector
    _, b2 = S.assemble(u, False)
    assert torch.equal(b2, b)


def test_newton_solution_and_errors_match_oracle():
    n = 8
    S, O = _system(n), paper_m0.PaperM0(n)
    u = S.solve()
    uo = O.solve()
    assert S.newton_its == O.newton_its
    assert np.linalg.norm(u.cpu().numpy() - uo) / np.linalg.norm(uo) < 1e-8
    assert np.allclose(S.errors(u), O.errors(uo), rtol=1e-8)
    assert np.allclose(S.sigD().cpu().numpy(), O.sigD(), rtol=1e-9, atol=1e-11)


def test_recorded_k0_table_through_the_gpu_path():
    """scripts/run_paper_mixed_convergence.py prints the table recorded at test_for_paper_convergence.py:200-207:
    levels 3-6 digit for digit, the two coarsest to the last printed digit (analytic data instead of
    degree-5/6 interpolants) -- exactly as the oracle does (tests/test_oracle_golden.py)."""
    from test_oracle_golden import PAPER_K0
    spec = importlib.util.spec_from_file_location("paper_conv", os.path.join(ROOT, "scripts",
                                                                             "run_paper_mixed_convergence.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rows, its = mod.main(6)
    got = [[float(t) for t in r.replace("&", " ").split()] for r in rows]
    ref = [[float(t) for t in r.replace("&", " ").split()] for r in PAPER_K0[:6]]
    for k in range(2, 6):
        assert rows[k].split() == PAPER_K0[k].split(), (rows[k], PAPER_K0[k])
    for k in range(2):
        assert got[k][0] == ref[k][0] and got[k][1] == ref[k][1]
        for a, b in zip(got[k][2::2], ref[k][2::2]):
            assert abs(a - b) <= 0.011 * 10 ** np.floor(np.log10(b)), (a, b)


# ---- k = 1: RT2 x DG1 x CG2 ---------------------------------------------------------------------------------
def _system_k1(n):
    from closestmolecule_b200.mixed import PaperMixedSystemK1
    mesh = cm.UnitSquareMesh(n, n)
    bd = cm.MeshFunction("size_t", mesh, 1, 0)
    cm.CompiledSubDomain("near(x[0],0) && on_boundary").mark(bd, 30)
    cm.CompiledSubDomain("near(x[0],1) && on_boundary").mark(bd, 31)
    cm.CompiledSubDomain("(near(x[1],0) || near(x[1],1)) && on_boundary").mark(bd, 32)
    return PaperMixedSystemK1(mesh, bd)


@pytest.mark.parametrize("n", [3, 8])
def test_k1_assembly_matches_oracle(n):
    from oracle import paper_m1
    S, O = _system_k1(n), paper_m1.PaperM1(n)
    rng = np.random.default_rng(2)
    un = np.concatenate([rng.standard_normal(2 * O.E + 2 * O.C), 0.5 + rng.random(3 * O.C),
                         1.0 + rng.random(O.V + O.E)])
    u = torch.tensor(un, device=S.mesh.device)
    vals, b = S.assemble(u, True)
    J, bo = O.assemble(un, True)
    got = sp.csr_matrix((vals.cpu().numpy(), S.col.cpu().numpy(), S.rowptr.cpu().numpy()), shape=(S.N, S.N))
    scale = abs(J).max()
    assert abs(got - J).max() < 1e-12 * scale, abs(got - J).max() / scale
    assert np.max(np.abs(b.cpu().numpy() - bo)) < 1e-12 * np.max(np.abs(bo))


def test_k1_newton_solution_and_errors_match_oracle():
    from oracle import paper_m1
    n = 4
    S, O = _system_k1(n), paper_m1.PaperM1(n)
    u = S.solve()
    uo = O.solve()
    assert S.newton_its == O.newton_its
    assert np.linalg.norm(u.cpu().numpy() - uo) / np.linalg.norm(uo) < 1e-8
    assert np.allclose(S.errors(u), O.errors(uo), rtol=1e-8)


def test_recorded_k1_table_through_the_gpu_path():
    """The first recorded k=1 table (test_for_paper_convergence.py:211-217) through the GPU path: levels 3-5 digit
    for digit, the two coarsest as the oracle (tests/test_oracle_golden.py::test_paper_k1_table_rt2_dg1_cg2)."""
    from test_oracle_golden import PAPER_K1
    spec = importlib.util.spec_from_file_location("paper_conv", os.path.join(ROOT, "scripts",
                                                                             "run_paper_mixed_convergence.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    levels = 5
    rows, its = mod.main(levels, deg=1)
    got = [[float(t) for t in r.replace("&", " ").split()] for r in rows]
    ref = [[float(t) for t in r.replace("&", " ").split()] for r in PAPER_K1[:levels]]
    for k in range(2, levels):
        assert rows[k].split() == PAPER_K1[k].split(), (rows[k], PAPER_K1[k])
    for k in range(2):
        assert got[k][:6] == ref[k][:6], (rows[k], PAPER_K1[k])
        assert abs(got[k][6] - ref[k][6]) <= 0.011 * 10 ** np.floor(np.log10(ref[k][6])), (got[k], ref[k])


def test_k1_cuda_facet_operator_matches_host_tabulation():
    """cm_cg2_exterior_facets / cm_cg2_interior_facets against the torch restatement (_facet_tables), which the CPU
    test tests/test_mixed_tables.py compares with the oracle."""
    S = _system_k1(6)
    fr, fcn, fv, cvec = S._facet_tables(6)
    Kf = torch.zeros(S.nnz, dtype=torch.float64, device=S.mesh.device)
    Kf.index_add_(0, torch.searchsorted(S._key, fr * S.N + fcn), fv)
    assert float((S.Kf - Kf).abs().max()) < 1e-12 * float(Kf.abs().max())
    assert float((S.cvec - cvec).abs().max()) < 1e-12 * float(cvec.abs().max())
    # deterministic: two builds give the same bits
    S2 = _system_k1(6)
    assert torch.equal(S.Kf, S2.Kf) and torch.equal(S.cvec, S2.cvec)


# ---- BASELINE config 2's own script: ...MixedPrimal_convergence.py on RT2 x DG1 x CG2 (Galerkin q-equation) ----
def _system_galerkin(n):
    from closestmolecule_b200.mixed import PaperMixedSystemK1
    mesh = cm.RectangleMesh(cm.Point(0, 0), cm.Point(1, 1), n, n)
    bd = cm.MeshFunction("size_t", mesh, 1, 0)
    cm.CompiledSubDomain("near(x[0],1) && on_boundary").mark(bd, 31)
    cm.CompiledSubDomain("(near(x[0],0) || near(x[1],0) || near(x[1],1)) && on_boundary").mark(bd, 32)
    return PaperMixedSystemK1(mesh, bd, D1=1e-3, D2=1.5e-3, degree=4, variant="galerkin")


def test_mixed_primal_galerkin_assembly_and_newton_match_oracle():
    from oracle import paper_m1
    n = 6
    S, O = _system_galerkin(n), paper_m1.MixedPrimalM1(n)
    rng = np.random.default_rng(4)
    un = np.concatenate([rng.standard_normal(2 * O.E + 2 * O.C), 0.5 + rng.random(3 * O.C),
                         1.0 + rng.random(O.V + O.E)])
    u = torch.tensor(un, device=S.mesh.device)
    vals, b = S.assemble(u, True)
    J, bo = O.assemble(un, True)
    got = sp.csr_matrix((vals.cpu().numpy(), S.col.cpu().numpy(), S.rowptr.cpu().numpy()), shape=(S.N, S.N))
    assert abs(got - J).max() < 1e-12 * abs(J).max()
    assert np.max(np.abs(b.cpu().numpy() - bo)) < 1e-12 * np.max(np.abs(bo))
    us, uo = S.solve(), O.solve()
    assert S.newton_its == O.newton_its
    assert np.linalg.norm(us.cpu().numpy() - uo) / np.linalg.norm(uo) < 1e-8
    assert np.allclose(S.errors(us), O.errors(uo), rtol=1e-8)


def test_mixed_primal_convergence_driver_matches_oracle_table():
    """scripts/run_mixed_primal_convergence.py against the oracle's table of the same script (no table is recorded
    in the reference): identical rows, DoF column 3E + 5C + V, e_0(p) of second order."""
    from oracle import paper_m1
    spec = importlib.util.spec_from_file_location("mp_conv", os.path.join(ROOT, "scripts",
                                                                          "run_mixed_primal_convergence.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rows = mod.main(5)
    ref = paper_m1.mixed_primal_table(5)
    assert [r.split() for r in rows] == [r.split() for r in ref]
    assert 1.95 < float(rows[-1].replace("&", " ").split()[5]) < 2.05


def test_mixed_primal_supg_driver_matches_oracle_table():
    """...MixedPrimal_convergence_SUPG.py on RT2 x DG1 x CG2 (SUPG test function, hK = CellDiameter, q data on x = 0):
    assembly parity at a random state, then the driver's table against the oracle's; q converges with order ~2."""
    from closestmolecule_b200.mixed import PaperMixedSystemK1
    from oracle import paper_m1
    n = 5
    mesh = cm.RectangleMesh(cm.Point(0, 0), cm.Point(1, 1), n, n)
    bd = cm.MeshFunction("size_t", mesh, 1, 0)
    cm.CompiledSubDomain("near(x[0],0) && on_boundary").mark(bd, 31)
    S = PaperMixedSystemK1(mesh, bd, D1=1e-3, D2=1.5e-3, degree=4, variant="supg")
    O = paper_m1.MixedPrimalM1(n, supg=True)
    rng = np.random.default_rng(6)
    un = np.concatenate([rng.standard_normal(2 * O.E + 2 * O.C), 0.5 + rng.random(3 * O.C),
                         1.0 + rng.random(O.V + O.E)])
    vals, b = S.assemble(torch.tensor(un, device=mesh.device), True)
    J, bo = O.assemble(un, True)
    got = sp.csr_matrix((vals.cpu().numpy(), S.col.cpu().numpy(), S.rowptr.cpu().numpy()), shape=(S.N, S.N))
    assert abs(got - J).max() < 1e-12 * abs(J).max()
    assert np.max(np.abs(b.cpu().numpy() - bo)) < 1e-12 * np.max(np.abs(bo))
    spec = importlib.util.spec_from_file_location("mp_conv", os.path.join(ROOT, "scripts",
                                                                          "run_mixed_primal_convergence.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rows = mod.main(5, variant="supg")
    assert [r.split() for r in rows] == [r.split() for r in paper_m1.mixed_primal_table(5, supg=True)]
    assert float(rows[-1].replace("&", " ").split()[7]) > 1.9


def test_mixed_primal_c0ip_driver_matches_oracle_table():
    """...MixedPrimal_convergence_C0IP.py on RT2 x DG1 x CG2 (inflow term on x = 0, jump penalty with avg CellDiameter
    and the script's minus sign, natural p data everywhere, q data on x = 1, s.n = 0 on x = 0): assembly parity at a
    random state, the CUDA facet tables against the torch tabulation, then the driver's table against the oracle's."""
    from closestmolecule_b200.mixed import PaperMixedSystemK1
    from oracle import paper_m1
    n = 5
    mesh = cm.RectangleMesh(cm.Point(0, 0), cm.Point(1, 1), n, n)
    bd = cm.MeshFunction("size_t", mesh, 1, 0)
    cm.CompiledSubDomain("near(x[0],0) && on_boundary").mark(bd, 31)
    cm.CompiledSubDomain("near(x[0],1) && on_boundary").mark(bd, 32)
    cm.CompiledSubDomain("(near(x[1],0) || near(x[1],1)) && on_boundary").mark(bd, 33)
    S = PaperMixedSystemK1(mesh, bd, D1=0.01, D2=0.015, stab=-0.00075, stab2=0.0, degree=4, variant="c0ip")
    O = paper_m1.MixedPrimalC0IPM1(n)
    fr, fcn, fv, cvec = S._facet_tables(4)
    Kt = torch.zeros(S.nnz, dtype=torch.float64, device=mesh.device)
    Kt.index_add_(0, torch.searchsorted(S._key, fr * S.N + fcn), fv)
    assert float((Kt - S.Kf).abs().max()) < 1e-13 * float(Kt.abs().max())
    assert float((cvec - S.cvec).abs().max()) < 1e-13 * float(cvec.abs().max())
    rng = np.random.default_rng(8)
    un = np.concatenate([rng.standard_normal(2 * O.E + 2 * O.C), 0.5 + rng.random(3 * O.C),
                         1.0 + rng.random(O.V + O.E)])
    vals, b = S.assemble(torch.tensor(un, device=mesh.device), True)
    J, bo = O.assemble(un, True)
    got = sp.csr_matrix((vals.cpu().numpy(), S.col.cpu().numpy(), S.rowptr.cpu().numpy()), shape=(S.N, S.N))
    assert abs(got - J).max() < 1e-12 * abs(J).max()
    assert np.max(np.abs(b.cpu().numpy() - bo)) < 1e-12 * np.max(np.abs(bo))
    spec = importlib.util.spec_from_file_location("mp_conv", os.path.join(ROOT, "scripts",
                                                                          "run_mixed_primal_convergence.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rows = mod.main(4, variant="c0ip")
    assert [r.split() for r in rows] == [r.split() for r in paper_m1.mixed_primal_c0ip_table(4)]
    assert 1.9 < float(rows[-1].replace("&", " ").split()[5]) < 2.1


# ---- scalar CG2 pure-advection scripts (SUPG_pure_advection_convergence.py, C0IP_pure_advection_spherical_...) ----
def _pure(n, kind, stab=None):
    from closestmolecule_b200.mixed import PureAdvectionCG2
    mesh = cm.RectangleMesh(cm.Point(0, 0), cm.Point(1, 1), n, n)
    bd = cm.MeshFunction("size_t", mesh, 1, 0)
    cm.CompiledSubDomain("near(x[0],0) && on_boundary").mark(bd, 31)
    cm.CompiledSubDomain("near(x[0],1) && on_boundary").mark(bd, 32)
    cm.CompiledSubDomain("(near(x[1],0) || near(x[1],1)) && on_boundary").mark(bd, 33)
    return PureAdvectionCG2(mesh, bd, kind, stab)


@pytest.mark.parametrize("kind", ["supg", "c0ip"])
def test_pure_advection_matrix_and_solution_match_the_scalar_oracle(kind):
    from oracle import pure_advection as pa
    n = 8
    S, O = _pure(n, kind), pa.PureAdvection(n, kind)
    assert S.N == O.N
    vals, b = S.assemble()
    A, bo = O.assemble()
    got = sp.csr_matrix((vals.cpu().numpy(), S.col.cpu().numpy(), S.rowptr.cpu().numpy()), shape=(S.N, S.N))
    assert abs(got - A).max() < 1e-12 * abs(A).max()
    assert np.max(np.abs(b.cpu().numpy() - bo)) < 1e-12 * np.max(np.abs(bo))
    d, g = O.bcs()
    assert np.array_equal(S.bc_dofs.cpu().numpy(), d) and np.allclose(S.bc_vals.cpu().numpy(), g, rtol=0, atol=1e-15)
    q, qo = S.solve(), O.solve()
    # the C0IP script's system is poorly conditioned already at n = 8 (cond ~ 5e7)
    assert np.linalg.norm(q.cpu().numpy() - qo) / np.linalg.norm(qo) < (1e-10 if kind == "supg" else 1e-7)
    assert abs(S.error(q) - O.error(qo)) < 1e-7 * O.error(qo)


def test_pure_advection_drivers_match_oracle_tables():
    from oracle import pure_advection as pa
    spec = importlib.util.spec_from_file_location("pa_conv", os.path.join(ROOT, "scripts",
                                                                          "run_pure_advection_convergence.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rows = mod.main(5, "supg")
    assert [r.split() for r in rows] == [r.split() for r in pa.table("supg", 5)]
    assert 1.9 < float(rows[-1].replace("&", " ").split()[3]) < 2.1
    # the C0IP script as written: compared on the levels where the system is still well enough conditioned
    assert [r.split() for r in mod.main(3, "c0ip")] == [r.split() for r in pa.table("c0ip", 3)]
    # ... and on five levels with the stabilising sign
    rows, ref, prev = mod.main(5, "c0ip", -0.005), [], None
    for nk in range(5):
        m = pa.PureAdvection(2 ** (nk + 1), "c0ip", stab=-0.005)
        e, h = m.error(m.solve()), m.mesh.hmax()
        ref.append('{:6d} & {:.4f} & {:6.2e} & {:.3f} '.format(
            m.N, h, e, 0.0 if prev is None else np.log(e / prev[1]) / np.log(h / prev[0])))
        prev = (h, e)
    assert [r.split() for r in rows] == [r.split() for r in ref]
    assert float(rows[-1].replace("&", " ").split()[3]) > 1.5


def test_mixed_primal_c0ip_other_bcs_driver_matches_oracle_table():
    """..._C0IP_otherBCs.py on RT2 x DG1 x CG2 (r2vec = (-1,0) in the q-equation: qmode 2 of cm_mixed_m1_cells and
    kind bit 8 of cm_cg2_exterior_facets; s = project(sig_ex) on the left, q = q_ex on the right)."""
    from closestmolecule_b200.mixed import PaperMixedSystemK1
    from oracle import paper_m1
    n = 5
    mesh = cm.RectangleMesh(cm.Point(0, 0), cm.Point(1, 1), n, n)
    bd = cm.MeshFunction("size_t", mesh, 1, 0)
    cm.CompiledSubDomain("near(x[0],0) && on_boundary").mark(bd, 30)
    cm.CompiledSubDomain("near(x[0],1) && on_boundary").mark(bd, 31)
    cm.CompiledSubDomain("(near(x[1],0) || near(x[1],1)) && on_boundary").mark(bd, 32)
    S = PaperMixedSystemK1(mesh, bd, D1=2.0, D2=1.5, stab=0.0075, stab2=0.0, degree=4, variant="c0ip_other")
    O = paper_m1.MixedPrimalC0IPOtherM1(n)
    rng = np.random.default_rng(9)
    un = np.concatenate([rng.standard_normal(2 * O.E + 2 * O.C), 0.5 + rng.random(3 * O.C),
                         1.0 + rng.random(O.V + O.E)])
    vals, b = S.assemble(torch.tensor(un, device=mesh.device), True)
    J, bo = O.assemble(un, True)
    got = sp.csr_matrix((vals.cpu().numpy(), S.col.cpu().numpy(), S.rowptr.cpu().numpy()), shape=(S.N, S.N))
    assert abs(got - J).max() < 1e-12 * abs(J).max()
    assert np.max(np.abs(b.cpu().numpy() - bo)) < 1e-12 * np.max(np.abs(bo))
    spec = importlib.util.spec_from_file_location("mp_conv", os.path.join(ROOT, "scripts",
                                                                          "run_mixed_primal_convergence.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rows = mod.main(4, variant="c0ip_other")
    assert [r.split() for r in rows] == [r.split() for r in paper_m1.mixed_primal_c0ip_table(4, other_bcs=True)]
    assert 1.9 < float(rows[-1].replace("&", " ").split()[5]) < 2.1


def test_mixed_primal_c0ip_modified_driver_matches_oracle_table():
    """..._C0IP_otherBCs_modified.py (weak outflow data, FacetArea, degree 6, no boundary penalty, s = project(sig_ex)
    on the left only): the 'paper' layout with other constants and the full-gradient q error."""
    from oracle import paper_m1
    spec = importlib.util.spec_from_file_location("mp_conv", os.path.join(ROOT, "scripts",
                                                                          "run_mixed_primal_convergence.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rows = mod.main(4, variant="c0ip_modified")
    assert [r.split() for r in rows] == [r.split() for r in paper_m1.mixed_primal_c0ip_table(4, modified=True)]
