"""Port of convergence/SphericalAdvectionDiffusionReaction_convergence.py onto closestmolecule_b200:
same loop, same solver parameters, same table format (compare with README.md:105-117 of the
reference).  Needs a CUDA device."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from closestmolecule_b200 import *  # noqa: E402,F401,F403
from closestmolecule_b200.functionals import error_norms  # noqa: E402
from closestmolecule_b200.manufactured import u_exact  # noqa: E402


def grad_u_exact(x, y):
    return (math.pi * torch.cos(math.pi * x) * torch.sin(math.pi * y),
            math.pi * torch.sin(math.pi * x) * torch.cos(math.pi * y))


def main(nkmax=7):
    hh, nn, eu, ru, e0, r0 = [], [], [], [0.0], [], [0.0]
    for nk in range(nkmax):
        print("....... Refinement level : nk = ", nk)
        nps = pow(2, nk + 1)
        mesh = RectangleMesh(Point(0, 0), Point(1, 1), nps, nps)
        hh.append(mesh.hmax())
        Vh = FunctionSpace(mesh, 'CG', 1)
        nn.append(Vh.dim())
        u = Function(Vh)
        bcU = DirichletBC(Vh, u_exact, 'on_boundary')
        FF = ScalarADRForm.nonlinear_adr(Vh, D1=1.e-3, D2=1.e-4)
        problem = NonlinearVariationalProblem(FF, u, bcs=bcU)
        solver = NonlinearVariationalSolver(problem)
        solver.parameters['nonlinear_solver'] = 'newton'
        solver.parameters['newton_solver']['linear_solver'] = 'mumps'
        solver.parameters['newton_solver']['absolute_tolerance'] = 1e-7
        solver.parameters['newton_solver']['relative_tolerance'] = 1e-7
        solver.solve()
        E0, E1 = error_norms(mesh, u.vector(), u_exact, grad_u_exact, degree=4)
        eu.append(E1)
        e0.append(E0)
        if nk > 0:
            ru.append(math.log(eu[nk] / eu[nk - 1]) / math.log(hh[nk] / hh[nk - 1]))
            r0.append(math.log(e0[nk] / e0[nk - 1]) / math.log(hh[nk] / hh[nk - 1]))
    print('====================================================')
    print('  DoF      h    e_1(u)   r_1(u)   e_0(u)  r_0(u)    ')
    print('====================================================')
    rows = []
    for nk in range(nkmax):
        rows.append('{:6d}  {:.4f} {:6.2e}  {:.3f}  {:6.2e}  {:.3f} '.format(nn[nk], hh[nk], eu[nk], ru[nk], e0[nk], r0[nk]))
        print(rows[-1])
    print('====================================================')
    return rows


if __name__ == "__main__":
    main()
