"""Port of convergence/SUPG_pure_advection_convergence.py and convergence/C0IP_pure_advection_spherical_convergence.py
(scalar CG2 pure-advection problems in the weighted (r2, r1) measure) onto the GPU path: same mesh ladder (:66-70),
markers inlet 31 / outlet 32 / char 33 (:72-78), constants (:55 stab = 0.1; :53 stab = 0.005), forms, boundary
condition and table layout (:133-138).  The reference records no table for either; the rows are identical to
oracle/pure_advection.py's.  SUPG converges with order 2 in the weighted H1 norm; the C0IP script as written (positive
stab) is nearly singular from n = 16 on -- pass a negative stab (the sign ..._MixedPrimal_convergence_C0IP.py:147 uses)
to see the orders 1.6-2.

usage: run_pure_advection_convergence.py [nkmax] [supg|c0ip] [stab]"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from closestmolecule_b200 import *  # noqa: E402,F401,F403
from closestmolecule_b200.mixed import PureAdvectionCG2  # noqa: E402


def main(nkmax=6, kind="supg", stab=None):
    hh, nn, eq, rq = [], [], [], [0.0]
    for nk in range(nkmax):
        nps = pow(2, nk + 1)
        mesh = RectangleMesh(Point(0, 0), Point(1, 1), nps, nps)
        bdry = MeshFunction("size_t", mesh, 1)
        bdry.set_all(0)
        inlet, outlet, char = 31, 32, 33
        CompiledSubDomain("near(x[0],0) && on_boundary").mark(bdry, inlet)
        CompiledSubDomain("near(x[0],1) && on_boundary").mark(bdry, outlet)
        CompiledSubDomain("(near(x[1],0) || near(x[1],1)) && on_boundary").mark(bdry, char)
        hh.append(mesh.hmax())
        problem = PureAdvectionCG2(mesh, bdry, kind, stab)
        nn.append(problem.N)
        q_h = problem.solve()
        eq.append(problem.error(q_h))
        if nk > 0:
            rq.append(math.log(eq[nk] / eq[nk - 1]) / math.log(hh[nk] / hh[nk - 1]))
    print('====================================================')
    print('  DoF      h      e_1(q)  r_1(q)    ')
    print('====================================================')
    rows = ['{:6d} & {:.4f} & {:6.2e} & {:.3f} '.format(nn[k], hh[k], eq[k], rq[k]) for k in range(nkmax)]
    print("\n".join(rows))
    print('====================================================')
    return rows


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 6, sys.argv[2] if len(sys.argv) > 2 else "supg",
         float(sys.argv[3]) if len(sys.argv) > 3 else None)
