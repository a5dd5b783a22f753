"""Manufactured-solution convergence study of the coupled P-Q Newton solve on the GPU path: the
CG1 x CG1 restatement of convergence/SphericalAdvectionDiffusionReaction_MixedPrimal_convergence_SUPG.py
(same exact solutions :48-49, constants :51-56, q Dirichlet data on the inlet x=0 :77,116, SUPG test
function :118-129, Newton tolerances :138-139, weighted error norms in the style of :147-149 and the
same table layout).  Needs a CUDA device."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from closestmolecule_b200 import *  # noqa: E402,F401,F403
from closestmolecule_b200.functionals import error_norms  # noqa: E402
from closestmolecule_b200.manufactured import p_exact, q_exact  # noqa: E402

pi = math.pi


def grad_p(x, y):
    return pi * torch.cos(pi * x) * torch.sin(pi * y), pi * torch.sin(pi * x) * torch.cos(pi * y)


def grad_q(x, y):
    return -pi * y * torch.sin(pi * x * y), -pi * x * torch.sin(pi * x * y)


def main(nkmax=8, report=False, linear_solver='umfpack'):
    hh, nn, ep, rp, e1p, r1p, eq, rq, its = [], [], [], [0.0], [], [0.0], [], [0.0], []
    for nk in range(nkmax):
        nps = pow(2, nk + 1)
        mesh = RectangleMesh(Point(0, 0), Point(1, 1), nps, nps)
        bdry = MeshFunction("size_t", mesh, 1)
        bdry.set_all(0)
        left = 31
        CompiledSubDomain("near(x[0],0) && on_boundary").mark(bdry, left)
        hh.append(mesh.hmax())
        P1 = FiniteElement('CG', mesh.ufl_cell(), 1)
        Vh = FunctionSpace(mesh, MixedElement([P1, P1]))
        nn.append(Vh.dim())
        u = Function(Vh)
        bcs = [DirichletBC(Vh.sub(0), p_exact, 'on_boundary'), DirichletBC(Vh.sub(1), q_exact, bdry, left)]
        FF = PQPrimalForm.manufactured(Vh, "supg", D1=1e-3, D2=1.5e-3, stab=1.0)
        solver = NonlinearVariationalSolver(NonlinearVariationalProblem(FF, u, bcs=bcs))
        solver.parameters['nonlinear_solver'] = 'newton'
        solver.parameters['newton_solver']['linear_solver'] = linear_solver   # 'umfpack' at …_SUPG.py:137
        solver.parameters['newton_solver']['absolute_tolerance'] = 1e-8
        solver.parameters['newton_solver']['relative_tolerance'] = 1e-8
        solver.parameters['newton_solver']['report'] = report
        n_it, _ = solver.solve()
        its.append((n_it, list(solver.linear_iterations)))
        p_h, q_h = u.split()
        e0, e1 = error_norms(mesh, p_h.contiguous(), p_exact, grad_p, degree=4, weighted=True)
        q0, q1 = error_norms(mesh, q_h.contiguous(), q_exact, grad_q, degree=4, weighted=True)
        ep.append(e0)
        e1p.append(e1)
        eq.append(math.sqrt(q0 ** 2 + q1 ** 2))
        if nk > 0:
            lg = math.log(hh[nk] / hh[nk - 1])
            rp.append(math.log(ep[nk] / ep[nk - 1]) / lg)
            r1p.append(math.log(e1p[nk] / e1p[nk - 1]) / lg)
            rq.append(math.log(eq[nk] / eq[nk - 1]) / lg)
    print('=================================================================================')
    print('  DoF      h      e_0(p)   r_0(p)   e_1(p)   r_1(p)   e_1(q)  r_1(q)   Newton (GMRES its)')
    print('=================================================================================')
    rows = []
    for nk in range(nkmax):
        rows.append('{:8d} & {:.4f} & {:6.2e} & {:.3f} & {:6.2e} & {:.3f} & {:6.2e} & {:.3f} & {} {}'.format(
            nn[nk], hh[nk], ep[nk], rp[nk], e1p[nk], r1p[nk], eq[nk], rq[nk], its[nk][0], its[nk][1]))
        print(rows[-1])
    print('=================================================================================')
    return dict(h=hh, ep=ep, e1p=e1p, eq=eq, rp=rp, r1p=r1p, rq=rq, its=its)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 8)
