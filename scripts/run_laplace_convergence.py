"""BASELINE config 1: convergence/SphericalLaplace_convergence.py on the GPU path -- the linear scalar P1 problem
u - D Lap_r2r1(u) = f on the unit square (D = 1e-4, u_ex = sin(pi x) sin(pi y), :86-90), mesh ladder 2^(k+1)
(:103-104), solve(auv == Fv, u_h, bcU) with the default direct solver (:135), H1-seminorm / L2 errors and rates in
the reference's table layout (:160).  The reference records no table for this script; the expected orders are
h in H1 and h^2 in L2 (README.md:118)."""
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
        nps = pow(2, nk + 1)
        mesh = RectangleMesh(Point(0, 0), Point(1, 1), nps, nps)
        hh.append(mesh.hmax())
        Vh = FunctionSpace(mesh, 'CG', 1)
        nn.append(Vh.dim())
        bcU = DirichletBC(Vh, u_exact, 'on_boundary')
        u_h = Function(Vh)
        LinearVariationalSolver(ScalarADRForm.laplace(Vh, D=1e-4), u_h, [bcU]).solve()     # solve(auv == Fv, u_h, bcU)
        l2, h1 = error_norms(mesh, u_h.vector(), u_exact, grad_u_exact, degree=6)
        eu.append(h1)
        e0.append(l2)
        if nk > 0:
            lg = math.log(hh[nk] / hh[nk - 1])
            ru.append(math.log(eu[nk] / eu[nk - 1]) / lg)
            r0.append(math.log(e0[nk] / e0[nk - 1]) / lg)
    print('====================================================')
    print('  DoF      h    e_1(u)   r_1(u)   e_0(u)  r_0(u)    ')
    print('====================================================')
    rows = ['{:6d}  {:.4f} {:6.2e}  {:.3f}  {:6.2e}  {:.3f} '.format(nn[k], hh[k], eu[k], ru[k], e0[k], r0[k])
            for k in range(nkmax)]
    print("\n".join(rows))
    print('====================================================')
    return dict(h=hh, dofs=nn, e1=eu, r1=ru, e0=e0, r0=r0, rows=rows)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 7)
