"""Port of convergence/test_for_paper_convergence.py (deg = 0: RT1 x DG0 x CG1, deg = 1: RT2 x DG1 x CG2) onto the
GPU path (closestmolecule_b200.mixed.PaperMixedSystem / PaperMixedSystemK1): same mesh ladder (:74,88), markers
(:90-94), data (:52-56,63-71), Newton tolerances (:161-164) and the same table layout (:195); the recorded tables
are at :200-207 (k=0) and :211-217 (k=1, stab = 0.1, stab2 = 2)."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from closestmolecule_b200 import *  # noqa: E402,F401,F403
from closestmolecule_b200.mixed import PaperMixedSystem, PaperMixedSystemK1  # noqa: E402


def main(nkmax=7, deg=0):
    hh, nn, es, ep, eq, rs, rp, rq, its = [], [], [], [], [], [0.0], [0.0], [0.0], []
    for nk in range(nkmax):
        nps = pow(2, nk + 1)
        mesh = UnitSquareMesh(nps, nps)
        bdry = MeshFunction("size_t", mesh, 1)
        bdry.set_all(0)
        left, right, rest = 30, 31, 32
        CompiledSubDomain("near(x[0],0) && on_boundary").mark(bdry, left)
        CompiledSubDomain("near(x[0],1) && on_boundary").mark(bdry, right)
        CompiledSubDomain("(near(x[1],0) || near(x[1],1)) && on_boundary").mark(bdry, rest)
        hh.append(mesh.hmax())
        if deg == 0:
            system = PaperMixedSystem(mesh, bdry, D1=1.0, D2=1.5, stab=0.05, stab2=1.0)      # :70-71
        else:
            system = PaperMixedSystemK1(mesh, bdry, D1=1.0, D2=1.5, stab=0.1, stab2=1.0 * (deg + 1))
        nn.append(system.N)
        u = system.solve(tol=1e-7)
        its.append(system.newton_its)
        e = system.errors(u)
        es.append(e[0]); ep.append(e[1]); eq.append(e[2])
        if nk > 0:
            lg = math.log(hh[nk] / hh[nk - 1])
            rs.append(math.log(es[nk] / es[nk - 1]) / lg)
            rp.append(math.log(ep[nk] / ep[nk - 1]) / lg)
            rq.append(math.log(eq[nk] / eq[nk - 1]) / lg)
    print('========================================================================')
    print('  DoF      h      e_div(s)   r_div(s)   e_0(p)   r_0(p)   e_1(q)  r_1(q)   ')
    print('========================================================================')
    rows = []
    for nk in range(nkmax):
        rows.append('{:6d} & {:.4f} & {:6.2e} & {:.3f} & {:6.2e} & {:.3f} & {:6.2e} & {:.3f} '.format(
            nn[nk], hh[nk], es[nk], rs[nk], ep[nk], rp[nk], eq[nk], rq[nk]))
        print(rows[-1])
    print('========================================================================')
    return rows, its


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 7, int(sys.argv[2]) if len(sys.argv) > 2 else 0)
