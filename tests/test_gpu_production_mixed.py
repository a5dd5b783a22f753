"""The production discretisation DG2 x RT3 x CG3 of advection_diffusion_mixed.py:94-269 on the GPU
(closestmolecule_b200.mixed.ProductionMixedSystem / solve_problem_instance_mixed) against the oracle
(oracle/production_mixed.py).  No output of this script ships with the reference (parity unpinned): the checks are
GPU vs oracle and the analytic reaction rate the reference prints next to its own total flux (:264-267)."""
import math

import numpy as np
import pytest
import scipy.sparse as sp
import torch

import closestmolecule_b200 as cm
from closestmolecule_b200.mixed import ProductionMixedSystem, solve_problem_instance_mixed
from oracle.production_mixed import ProductionMixed

from util import oracle_mesh

pytestmark = pytest.mark.gpu
SIGMA, GAMMA, D1, D2 = 0.1, 1.0, 2.0, 5.5                      # correct_boundary_finite_element.py:91-103
V1 = 4 * math.pi * (5.0 - SIGMA) ** 3 / 3


def _setup(n, c, dt=None):
    mesh = cm.MappedRectangleMesh(n, n, 5.0, 5.0, SIGMA, GAMMA)
    bdry = cm.structured_boundary_markers(mesh)
    S = ProductionMixedSystem(mesh, bdry, c, SIGMA, GAMMA, V1, D1, D2, dt)
    O = ProductionMixed(oracle_mesh(mesh), bdry.values.cpu().numpy(), c, SIGMA, GAMMA, V1, D1, D2, dt)
    return mesh, bdry, S, O


@pytest.mark.parametrize("dt", [None, 0.1])
def test_assembly_matches_oracle(dt):
    mesh, bdry, S, O = _setup(5, 2.0, dt)
    assert S.N == O.N and np.array_equal(S.cell_dofs.cpu().numpy(), O.cell_dofs)
    rng = np.random.default_rng(3)
    un = np.concatenate([1e-3 * (1 + rng.random(O.np_)), 1e-3 * rng.standard_normal(O.ns),
                         1e-4 * (1 + rng.random(O.nq_))])
    po = 1e-3 * (1 + rng.random(O.N))
    u = torch.tensor(un, device=mesh.device)
    S.p_old.copy_(torch.tensor(po[:O.np_], device=mesh.device))
    vals, b = S.assemble(u, True)
    J, bo = O.assemble(un, True, po if dt else None)
    got = sp.csr_matrix((vals.cpu().numpy(), S.col.cpu().numpy(), S.rowptr.cpu().numpy()), shape=(S.N, S.N))
    assert abs(got - J).max() < 1e-11 * abs(J).max()
    assert np.max(np.abs(b.cpu().numpy() - bo)) < 1e-11 * np.max(np.abs(bo))


def test_steady_solve_flux_and_results_row():
    n, c = 8, 1.0
    mesh, bdry, S, O = _setup(n, c)
    results = []
    flux, total = solve_problem_instance_mixed(c, 0.1, 0.1, mesh, bdry, SIGMA, GAMMA, results, V1, D1, D2,
                                               None, True)
    Sg = solve_problem_instance_mixed.last_system
    uo = O.solve()
    assert O.converged and Sg.newton_its == O.newton_its
    assert np.linalg.norm(Sg.u.cpu().numpy() - uo) / np.linalg.norm(uo) < 1e-8
    fo = O.flux(uo)
    assert np.linalg.norm(flux.cpu().numpy() - fo) / np.linalg.norm(fo) < 1e-8
    r2, r1, tot = O.bottom_flux(fo)
    assert abs(total - tot) < 1e-8 * abs(tot)
    row = results[0]
    assert row[:3] == [SIGMA, GAMMA, c] and abs(row[3] - r1) < 1e-8 * abs(tot) and abs(row[4] - r2) < 1e-8 * abs(tot)
    assert row[5] == total and abs(row[6] - 4 * math.pi * D1 * SIGMA * c / (GAMMA + c)) < 1e-14


def test_total_flux_approaches_the_analytic_reaction_rate():
    """advection_diffusion_mixed.py:264-267 prints total_flux next to 4 pi D1 sigma c/(gamma+c): on a 24 x 24
    boundary-fitted mesh the two agree to a few per cent (the oracle gives 1.2664 vs 1.2566 at c = 1)."""
    n, c = 24, 1.0
    mesh = cm.MappedRectangleMesh(n, n, 5.0, 5.0, SIGMA, GAMMA)
    bdry = cm.structured_boundary_markers(mesh)
    results = []
    _, total = solve_problem_instance_mixed(c, 0.1, 0.1, mesh, bdry, SIGMA, GAMMA, results, V1, D1, D2, None, True)
    assert abs(total - 1.26643) < 2e-4, total
    assert abs(total / results[0][6] - 1.0) < 0.05


def test_boundary_correction_driver_runs_a_least_squares_step():
    """scripts/run_boundary_correction.py (correct_boundary_finite_element.py): two evaluations of the flux
    residuals on a coarse mesh; the residual vector has one entry per concentration and is finite."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bcorr", os.path.join(root, "scripts", "run_boundary_correction.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sol = mod.main(n=12, max_nfev=2, c_vals=(1, 2))   # (n = 8, c = 2) does not converge in the oracle either: Gstar branch flips
    assert sol.fun.shape == (2,) and np.all(np.isfinite(sol.fun)) and np.all(sol.x > 0)


def test_boundary_correction_approaches_the_reference_fit():
    """The fit of correct_boundary_finite_element.py on a graded 32 x 32 mesh: (sigma, gamma) lands next to the
    values the reference script starts from (sigma_init = 0.09867282, gamma_init = 1.20166547, :110-111, the
    result of its own earlier runs); measured here: (0.09640, 1.20210) at n = 32, (0.09848, 1.20363) at n = 48."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bcorr", os.path.join(root, "scripts", "run_boundary_correction.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    sol = mod.main(n=32, max_nfev=8)
    assert abs(sol.x[0] / 0.09867282 - 1.0) < 0.03 and abs(sol.x[1] / 1.20166547 - 1.0) < 0.005, sol.x
    assert np.max(np.abs(sol.fun)) < 0.05, sol.fun


def test_projections_reproduce_polynomials_and_postprocessing_writes_xdmf(tmp_path):
    """project() onto DG2 / CG3 / RT3 (advection_diffusion_mixed.py:238-254): polynomials of the space's degree are
    reproduced; the five post-processed fields have the documented meaning; the XDMF time series is well formed and
    its binary side file holds the vertex values."""
    import xml.etree.ElementTree as ET
    n, c = 6, 1.0
    mesh, bdry, S, O = _setup(n, c)
    x, y = S.Pq[..., 0], S.Pq[..., 1]
    f = 1.0 + 0.3 * x - 0.2 * y + 0.05 * x * y + 0.01 * y * y                     # P2: in DG2 and CG3
    for kind in ("dg2", "cg3"):
        co = S.project(kind, f)
        assert torch.max(torch.abs(S.evaluate(kind, co.flatten()) - f)) < 1e-11
        X = mesh.coords
        fv = 1.0 + 0.3 * X[:, 0] - 0.2 * X[:, 1] + 0.05 * X[:, 0] * X[:, 1] + 0.01 * X[:, 1] ** 2
        assert torch.max(torch.abs(S.vertex_values(kind, co.flatten()) - fv)) < 1e-11
    g = torch.stack([0.5 + 0.1 * x - 0.3 * y, -0.2 + 0.4 * x + 0.25 * y], -1)     # P1^2: in RT3
    cg = S.project("rt3", g)
    assert torch.max(torch.abs(S.evaluate("rt3", cg) - g)) < 1e-10
    # post-processing after a steady solve, written through the driver
    results = []
    out = str(tmp_path / "solution.xdmf")
    solve_problem_instance_mixed(c, 0.1, 0.1, mesh, bdry, SIGMA, GAMMA, results, V1, D1, D2, out, True)
    Sg = solve_problem_instance_mixed.last_system
    fields = Sg.postprocess()
    assert set(fields) == {"p", "q", "flux", "dif", "dif_flux"}
    # clipped field: q_h undershoots on this coarse mesh (that is why the script clips); the projection of
    # max(q_h, 0) is L2-orthogonal to the space: its moments against the CG3 basis equal those of the clipped field
    q_h = Sg.evaluate("cg3", Sg.u[Sg.q0:])
    assert float(q_h.min()) < 0
    qc = torch.clamp(q_h, min=0.0)
    a, Phi = Sg._qweights(), Sg._p3_values()
    mom = lambda fq: torch.zeros(Sg.V + 2 * Sg.E + Sg.C, dtype=torch.float64, device=fq.device).index_add_(
        0, (Sg.qd - Sg.q0).flatten(), torch.einsum("cq,qa,cq->ca", a, Phi, fq).flatten())
    m_proj, m_clip = mom(Sg.evaluate("cg3", fields["q"])), mom(qc)
    assert torch.max(torch.abs(m_proj - m_clip)) < 1e-10 * float(torch.max(torch.abs(m_clip)))
    assert float(Sg.evaluate("cg3", fields["q"]).min()) > float(q_h.min())
    # flux equals the driver's flux; dif = project(p - p_approx): with p_approx = 4 pi r2^2 interpolate(PInitial, DG2)
    # projected separately, the three DG2 fields add up (linearity of the projection)
    assert torch.equal(fields["flux"], Sg.flux())
    Xc = Sg.Xc
    nodes = torch.cat([Xc, 0.5 * (Xc[:, [1, 0, 0]] + Xc[:, [2, 2, 1]])], 1)
    xq = Sg.Pq[..., 0]
    pa_q = 4 * math.pi * xq * xq * Sg.evaluate("dg2", Sg.p_initial(nodes[..., 0], nodes[..., 1]).flatten())
    pa = Sg.project("dg2", pa_q).flatten()
    assert torch.max(torch.abs(fields["p"] - pa - fields["dif"])) < 1e-10 * float(torch.max(torch.abs(fields["p"])))
    root = ET.parse(out).getroot()
    grids = root.findall("./Domain/Grid/Grid")
    assert len(grids) == 1
    names = [a.get("Name") for a in grids[0].findall("Attribute")]
    assert names == ["p", "q", "flux", "dif", "dif_flux"]
    raw = np.fromfile(str(tmp_path / "solution.bin"), dtype=np.uint8)
    V = mesh.num_vertices()
    for a in grids[0].findall("Attribute"):
        item = a.find("DataItem")
        dims = [int(d) for d in item.get("Dimensions").split()]
        assert dims[0] == V and (len(dims) == 1 or dims[1] == 3)
        off = int(item.get("Seek"))
        vals = np.frombuffer(raw[off:off + 8 * int(np.prod(dims))].tobytes(), dtype="<f8")
        assert np.isfinite(vals).all()
        if a.get("Name") == "q":
            assert np.allclose(vals, Sg.vertex_values("cg3", fields["q"]).cpu().numpy())
