"""a12 on the GPU (cm_functional_error_norms, cm_project_flux_dg0, cm_functional_boundary_flux) against
the oracle restatement (oracle/functionals.py), and the reference-signature flux driver."""
import importlib.util
import math
import os

import numpy as np
import pytest
import torch

import closestmolecule_b200 as cm
from closestmolecule_b200 import functionals as fn
from closestmolecule_b200.manufactured import p_exact, q_exact
from oracle import fem, functionals as ofn

from util import oracle_mesh, random_state

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _meshes():
    yield cm.RectangleMesh(cm.Point(0, 0), cm.Point(1, 1), 37, 23)
    yield cm.MappedRectangleMesh(48, 40, 5.0, 5.0, 0.1, 1.0)


@pytest.mark.parametrize("weighted", [False, True])
def test_error_norms_match_oracle(weighted):
    gp = lambda x, y: (math.pi * torch.cos(math.pi * x) * torch.sin(math.pi * y),
                       math.pi * torch.sin(math.pi * x) * torch.cos(math.pi * y))
    gpn = lambda x, y: (np.pi * np.cos(np.pi * x) * np.sin(np.pi * y), np.pi * np.sin(np.pi * x) * np.cos(np.pi * y))
    pn = lambda x, y: np.sin(np.pi * x) * np.sin(np.pi * y)
    for mesh in _meshes():
        un = random_state(mesh.num_vertices(), seed=5)
        u = torch.tensor(un, device=mesh.device)
        om = oracle_mesh(mesh)
        for deg in (2, 4, 6):
            for xo in (False, True):
                e = fn.error_norms(mesh, u[0::2], p_exact, gp, degree=deg, weighted=weighted, x_derivative_only=xo)
                r = ofn.generic_error_norms(om, un[0::2], pn, gpn, deg, weighted, xo)
                assert np.allclose(e, r, rtol=1e-13), (deg, xo, e, r)
        # contiguous input gives the same bits as the strided view
        assert fn.error_norms(mesh, u[0::2].contiguous(), p_exact, gp, weighted=weighted) == \
            fn.error_norms(mesh, u[0::2], p_exact, gp, weighted=weighted)


def test_flux_projection_and_boundary_flux_match_oracle():
    for mesh in _meshes():
        un = random_state(mesh.num_vertices(), seed=7)
        u = torch.tensor(un, device=mesh.device)
        om = oracle_mesh(mesh)
        F = fn.project_flux_dg0(mesh, u, 2.0, 5.5)
        Fo = ofn.project_flux_dg0(om, un, 2.0, 5.5)
        assert np.max(np.abs(F.cpu().numpy() - Fo)) <= 1e-13 * np.max(np.abs(Fo))
        bd = cm.structured_boundary_markers(mesh)
        markers = bd.values.cpu().numpy()
        for mk in (21, 22, 23, 24):
            g = fn.boundary_flux(mesh, F, bd, mk)
            o = ofn.boundary_flux(om, Fo, markers, mk)
            assert np.allclose(g, o, rtol=1e-12, atol=1e-12 * max(1.0, abs(o[2]))), (mk, g, o)


def test_reference_signature_flux_driver():
    """solve_problem_instance(...) -> (flux, total_flux) and the results row of
    advection_diffusion_mixed.py:268 through the GPU path; the total bottom flux is compared with the oracle's
    post-processing of the oracle's own Newton solution and with the analytic estimate 4 pi D1 sigma c/(gamma+c)."""
    from oracle import pq
    from oracle.newton import newton_solve
    from util import oracle_params
    spec = importlib.util.spec_from_file_location("prod_inst", os.path.join(ROOT, "scripts", "run_production_instance.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sigma, gamma, c, D1, D2, n = 0.1, 1.0, 2.0, 2.0, 5.5, 40
    V1 = 4 * math.pi * (5.0 - sigma) ** 3 / 3
    mesh = cm.MappedRectangleMesh(n, n, 5.0, 5.0, sigma, gamma)
    bdry = cm.structured_boundary_markers(mesh)
    results = []
    flux, total = mod.solve_problem_instance_flux(c, 0.1, 0.1, mesh, bdry, sigma, gamma, results, V1, D1, D2,
                                                  "unused.xdmf", True)
    assert len(results) == 1 and len(results[0]) == 7 and results[0][5] == total
    assert abs(results[0][3] + results[0][4] - total) < 1e-12 * abs(total)
    s = mod.solve_problem_instance.last_solver
    om = oracle_mesh(mesh)
    S = pq.PQSystem(om, oracle_params(s.problem.form))
    dofs, vals = cm.DirichletBC.merge(s.problem.bcs, 2 * mesh.num_vertices())
    x, oit, oconv = newton_solve(lambda xx, wj: S.assemble(xx, wj, None, None), np.zeros(S.N), S.rowptr, S.col,
                                 dofs.cpu().numpy().astype(np.int64), vals.cpu().numpy(), atol=1e-9, rtol=1e-9)
    ref = V1 * ofn.boundary_flux(om, ofn.project_flux_dg0(om, x, D1, D2), bdry.values.cpu().numpy(), 24)[2]
    assert oconv and abs(total - ref) < 1e-7 * abs(ref), (total, ref)
    # physical sanity (coarse P1 mesh, DG0 flux, outward normal): order of magnitude of the analytic rate
    assert 0.05 < abs(total) / results[0][6] < 5.0, (total, results[0][6])
