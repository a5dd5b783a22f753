"""GPU run of the reference's scalar nonlinear advection-diffusion-reaction convergence study
(reference driver: convergence/SphericalAdvectionDiffusionReaction_convergence.py; its printed table
is recorded in README.md:109-115 of the reference and is reproduced digit for digit by this run).

Same problem data (u_ex = sin(pi x) sin(pi y), D2 = 1e-4, D1 = 1e-3, degree-4 quadrature, Newton with
absolute/relative tolerance 1e-7 from a zero initial guess), solved through closestmolecule_b200.
Needs a CUDA device."""
import math
import os
import sys
from dataclasses import dataclass

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import closestmolecule_b200 as cm  # noqa: E402
from closestmolecule_b200.functionals import error_norms  # noqa: E402
from closestmolecule_b200.manufactured import u_exact  # noqa: E402

ROW_FORMAT = '{:6d}  {:.4f} {:6.2e}  {:.3f}  {:6.2e}  {:.3f} '   # the reference's table row layout


@dataclass
class Level:
    dofs: int
    h: float
    err_h1: float
    err_l2: float
    rate_h1: float = 0.0
    rate_l2: float = 0.0

    def text(self):
        return ROW_FORMAT.format(self.dofs, self.h, self.err_h1, self.rate_h1, self.err_l2, self.rate_l2)


def exact_gradient(x, y):
    return (math.pi * torch.cos(math.pi * x) * torch.sin(math.pi * y),
            math.pi * torch.sin(math.pi * x) * torch.cos(math.pi * y))


def solve_level(n):
    mesh = cm.RectangleMesh(cm.Point(0, 0), cm.Point(1, 1), n, n)
    space = cm.FunctionSpace(mesh, 'CG', 1)
    uh = cm.Function(space)
    problem = cm.NonlinearVariationalProblem(cm.ScalarADRForm.nonlinear_adr(space, D1=1.e-3, D2=1.e-4), uh,
                                             bcs=cm.DirichletBC(space, u_exact, 'on_boundary'))
    newton = cm.NonlinearVariationalSolver(problem)
    newton.parameters['nonlinear_solver'] = 'newton'
    newton.parameters['newton_solver'].update(linear_solver='mumps', absolute_tolerance=1e-7,
                                              relative_tolerance=1e-7)
    newton.solve()
    l2, h1 = error_norms(mesh, uh.vector(), u_exact, exact_gradient, degree=4)
    return Level(space.dim(), mesh.hmax(), h1, l2)


def main(levels=7):
    table = []
    for k in range(levels):
        lev = solve_level(2 ** (k + 1))
        if table:
            prev = table[-1]
            ratio = math.log(lev.h / prev.h)
            lev.rate_h1 = math.log(lev.err_h1 / prev.err_h1) / ratio
            lev.rate_l2 = math.log(lev.err_l2 / prev.err_l2) / ratio
        table.append(lev)
    rows = [lev.text() for lev in table]
    print("  DoF      h    e_1(u)   r_1(u)   e_0(u)  r_0(u)")
    print("\n".join(rows))
    return rows


if __name__ == "__main__":
    main()
