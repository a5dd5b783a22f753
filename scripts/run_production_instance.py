"""Port of solve_problem_instance (convergence/advection_diffusion_convergence.py:105-228, the flat /
curved-boundary production problem of advection_diffusion_mixed.py:94-269) onto closestmolecule_b200
for the CG1 x CG1 primal formulation: same mesh labels (right=21, top=22, left=23, bottom=24), same
Dirichlet list and order (:174-183), same data formulas (PInitial/QInitial :40-44,79-80), same Newton
parameters (:207-215), steady state or backward-Euler time loop (:196-204,219-226).

supg_stab=0 (default) is the reference's unstabilised Galerkin q-equation; it needs a pivoting sparse
LU, which linear_solver='mumps' provides on the GPU (banded LU with partial pivoting,
cm_blu_*) for meshes whose factors fit in HBM.  Beyond that size use linear_solver='gmres' together
with the SUPG test function of …MixedPrimal_convergence_SUPG.py:118-129 (supg_stab=1); note that the
multigrid-preconditioned Krylov path still stalls on the production constants.
"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from closestmolecule_b200 import *  # noqa: E402,F401,F403


def solve_problem_instance(concentration, t_final, dt, mesh, bdry, sigma, deg=1, steady_state=False,
                           D1=2.0, D2=1.5, r1_max=5.0, supg_stab=0.0, report=False, sigma_fn=None,
                           linear_solver='mumps', V1=None, production_gstar=False):
    assert deg == 1, "the GPU hot path is CG1 x CG1"
    right, top, left, bottom = 21, 22, 23, 24
    if V1 is None:
        V1 = 4 * math.pi * (r1_max - sigma) ** 3 / 3
    c = float(concentration)
    sig = sigma_fn if sigma_fn is not None else (lambda x: sigma)

    def p_initial(x, y):          # PInitial.eval (:40-44)
        return c / V1 * torch.exp(-4 / 3 * math.pi * c * x ** 3) * (1 - sig(x) / y)

    def p_right(x, y):            # PInitial(right=True)
        base = c / V1 * torch.exp(-4 / 3 * math.pi * c * x ** 3)
        return torch.where(y > 0, base * (1 - sig(x) / y), base)

    def q_initial(x, y):          # QInitial.eval (:79-80)
        return 1 / (4 * math.pi * V1) * torch.exp(-4 / 3 * math.pi * c * x ** 3)

    P0 = FiniteElement('CG', mesh.ufl_cell(), deg)
    P1 = FiniteElement('CG', mesh.ufl_cell(), deg)
    mixed_space = FunctionSpace(mesh, MixedElement([P0, P1]))
    if steady_state:
        t_final = 0
    t = 0
    u = Function(mixed_space)
    p_old = interpolate(p_initial, mixed_space.sub(0).collapse())
    bc = [DirichletBC(mixed_space.sub(0), p_right, bdry, right),
          DirichletBC(mixed_space.sub(0), p_initial, bdry, top),
          DirichletBC(mixed_space.sub(0), Constant(0.), bdry, bottom),
          DirichletBC(mixed_space.sub(1), q_initial, bdry, right)]
    gkind = "gstar_prod" if production_gstar else "gstar_conv"   # advection_diffusion_mixed.py:17-23 / …convergence.py:21-27
    FF = PQPrimalForm(mixed_space, D1, D2, G=(gkind, c), dt=None if steady_state else dt,
                      p_old=None if steady_state else p_old, supg_stab=supg_stab)
    problem = NonlinearVariationalProblem(FF, u, bcs=bc)
    solver = NonlinearVariationalSolver(problem)
    solver.parameters['nonlinear_solver'] = 'newton'
    solver.parameters['newton_solver']['linear_solver'] = linear_solver   # 'mumps' at :212
    solver.parameters['newton_solver']['absolute_tolerance'] = 1e-9
    solver.parameters['newton_solver']['relative_tolerance'] = 1e-9
    solver.parameters['newton_solver']['maximum_iterations'] = 10
    solver.parameters['newton_solver']['report'] = report
    log = []
    while t <= t_final:
        print("t=%.3f" % t)
        its, conv = solver.solve()
        log.append((its, list(solver.linear_iterations)))
        p_h, q_h = u.split()
        p_old.vector().copy_(p_h)          # assign(p_old, p_h)
        t += dt
    solve_problem_instance.last_solver = solver      # for the parity tests
    return u, log


def solve_problem_instance_flux(concentration, t_final, dt, mesh, bdry, sigma, gamma, results, V1, D1, D2,
                                output_filename=None, steady_state=False, **kw):
    """Signature, return value and `results` row of advection_diffusion_mixed.py:94-95,261-269 (the call
    correct_boundary_finite_element.py:80 makes) on the CG1 x CG1 primal formulation: the flux is the DG0
    projection of convergence/advection_diffusion_convergence.py:222-223 and the boundary functionals are
    V1*assemble(dot(flux,n)*4*pi*r1**2*ds(bottom)) with their r1 / r2 parts.  No XDMF file is written
    (output_filename is accepted and ignored)."""
    from closestmolecule_b200.functionals import boundary_flux, project_flux_dg0
    sfn = (lambda x: sigma * torch.exp(-4 / 3 * math.pi * gamma * x ** 3)) if gamma else None
    u, _ = solve_problem_instance(concentration, t_final, dt, mesh, bdry, sigma, steady_state=steady_state,
                                  D1=D1, D2=D2, V1=V1, sigma_fn=sfn, production_gstar=True, **kw)
    flux = project_flux_dg0(mesh, u.vector(), D1, D2)
    r2f, r1f, tot = boundary_flux(mesh, flux, bdry, 24)
    total_r2_flux, total_r1_flux, total_flux = V1 * r2f, V1 * r1f, V1 * tot
    total_approx_flux = 4 * math.pi * D1 * sigma * (concentration / (gamma + concentration))   # :264
    results.append([sigma, gamma, concentration, total_r1_flux, total_r2_flux, total_flux, total_approx_flux])
    return flux, total_flux


def finite_element_fluxes(parameters, c_vals, r1_max, r2_max, n, t_final, dt, D1, D2, steady_state, **kw):
    """correct_boundary_finite_element.py:59-86 with the FreeFem++ meshing replaced by the boundary-fitted
    structured mesh (MappedRectangleMesh): total bottom flux for every concentration on the mesh of
    (sigma, gamma) = parameters."""
    sigma, gamma = float(parameters[0]), float(parameters[1])
    V1 = 4 * math.pi * (r1_max - sigma) ** 3 / 3
    mesh = MappedRectangleMesh(n, n, r2_max, r1_max, sigma, gamma)
    bdry = structured_boundary_markers(mesh)
    out = []
    for c in c_vals:
        _, total_flux = solve_problem_instance_flux(c, t_final, dt, mesh, bdry, sigma, gamma, [], V1, D1, D2,
                                                    None, steady_state, r1_max=r1_max, **kw)
        out.append(total_flux)
    return out


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    sigma, gamma = 0.1, 1.0
    for kind in ("flat", "curved"):
        if kind == "flat":
            mesh = RectangleMesh(Point(0, sigma), Point(5, 5), n, n)
            sfn = None
        else:
            mesh = MappedRectangleMesh(n, n, 5.0, 5.0, sigma, gamma)
            sfn = lambda x: sigma * torch.exp(-4 / 3 * math.pi * gamma * x ** 3)
        bdry = structured_boundary_markers(mesh)
        for c in (1, 4):
            for steady in (True, False):
                try:
                    u, log = solve_problem_instance(c, 0.1, 0.1, mesh, bdry, sigma, steady_state=steady, sigma_fn=sfn)
                    print(kind, "n", n, "c", c, "steady" if steady else "transient", "Newton/linear its", log,
                          "max p %.3e" % float(u.split()[0].max()))
                except Exception as e:  # noqa: BLE001
                    print(kind, "n", n, "c", c, "steady" if steady else "transient", "FAILED:", str(e)[:150])
