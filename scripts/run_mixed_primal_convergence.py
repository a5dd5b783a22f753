"""Port of convergence/SphericalAdvectionDiffusionReaction_MixedPrimal_convergence.py (BASELINE config 2's own script)
onto the GPU path with ITS OWN discretisation RT2 x DG1 x CG2 (deg = 1, :93-97): same mesh ladder (:71-72), markers
('left' = {x = 1} -> 31, rest -> 32, :74-80), constants (:51-56), forms (:120-128), boundary condition (:117), Newton
parameters (:132-138) and table layout (:166).  The reference records no table for it; the Galerkin q-equation gives
the orders ~1.5 / 2 / 1 for e_div(s) / e_0(p) / e_1(q) (the loss in q is why the SUPG / C0IP variants exist).
Variants 'supg' and 'c0ip' are the ..._SUPG.py / ..._C0IP.py siblings on the same elements.
The CG1 x CG1 restatement of the same problem is scripts/run_pq_supg_convergence.py."""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from closestmolecule_b200 import *  # noqa: E402,F401,F403
from closestmolecule_b200.mixed import PaperMixedSystemK1  # noqa: E402


def main(nkmax=6, variant="galerkin"):
    """variant 'galerkin': ...MixedPrimal_convergence.py; 'supg': ...MixedPrimal_convergence_SUPG.py (q data on
    {x = 0} :77, stab = 1 :56, Newton tolerances 1e-8 :138-139, nkmax = 5 :58); 'c0ip': ..._convergence_C0IP.py
    (markers left 31 / right 32 / rest 33 :78-92, D1 = .01, D2 = .015 :63-64, stab = 0.00075 entering with the minus
    sign :68,147, q = q_ex on the right and s.n = 0 on the left :132-135, Newton 1e-8 :160-161); 'c0ip_other':
    ..._convergence_C0IP_otherBCs.py (r2vec = (-1,0) in the q-equation :58, D1 = 2, D2 = 1.5, stab = +0.0075,
    s = project(sig_ex) on the left :121-122, q = q_ex on the right :123); 'c0ip_modified':
    ..._C0IP_otherBCs_modified.py (r2vec = (1,0), q data imposed weakly on the right :164, hK = FacetArea :111,
    quadrature degree 6 :38, UnitSquareMesh :98, tolerances 1e-7 :173-174)."""
    hh, nn, es, ep, eq, rs, rp, rq = [], [], [], [], [], [0.0], [0.0], [0.0]
    for nk in range(nkmax):
        nps = pow(2, nk + 1)
        mesh = RectangleMesh(Point(0, 0), Point(1, 1), nps, nps)
        bdry = MeshFunction("size_t", mesh, 1)
        bdry.set_all(0)
        full_grad = None
        if variant == "c0ip_modified":  # ..._C0IP_otherBCs_modified.py:43-47,98-111: the paper's layout of the facet
            # terms (weak outflow data, hK = FacetArea, degree 6) without the boundary penalty; left 30/right 31/rest 32
            mesh = UnitSquareMesh(nps, nps)
            bdry = MeshFunction("size_t", mesh, 1)
            bdry.set_all(0)
            CompiledSubDomain("near(x[0],0) && on_boundary").mark(bdry, 30)
            CompiledSubDomain("near(x[0],1) && on_boundary").mark(bdry, 31)
            CompiledSubDomain("(near(x[1],0) || near(x[1],1)) && on_boundary").mark(bdry, 32)
            system = PaperMixedSystemK1(mesh, bdry, D1=2.0, D2=1.5, stab=1e-3, stab2=0.0, degree=6, variant="paper")
            full_grad = True
        elif variant == "c0ip_other":     # ..._C0IP_otherBCs.py:55-60,78-82: left 30 / right 31 / rest 32, +stab
            CompiledSubDomain("near(x[0],0) && on_boundary").mark(bdry, 30)
            CompiledSubDomain("near(x[0],1) && on_boundary").mark(bdry, 31)
            CompiledSubDomain("(near(x[1],0) || near(x[1],1)) && on_boundary").mark(bdry, 32)
            system = PaperMixedSystemK1(mesh, bdry, D1=2.0, D2=1.5, stab=0.0075, stab2=0.0, degree=4,
                                        variant=variant)
        elif variant == "c0ip":
            CompiledSubDomain("near(x[0],0) && on_boundary").mark(bdry, 31)
            CompiledSubDomain("near(x[0],1) && on_boundary").mark(bdry, 32)
            CompiledSubDomain("(near(x[1],0) || near(x[1],1)) && on_boundary").mark(bdry, 33)
            system = PaperMixedSystemK1(mesh, bdry, D1=0.01, D2=0.015, stab=-0.00075, stab2=0.0, degree=4,
                                        variant=variant)
        else:
            left, rest = 31, 32
            xin, xout = (0, 1) if variant == "supg" else (1, 0)
            CompiledSubDomain(f"near(x[0],{xin}) && on_boundary").mark(bdry, left)
            CompiledSubDomain(f"(near(x[0],{xout}) || near(x[1],0) || near(x[1],1)) && on_boundary").mark(bdry, rest)
            system = PaperMixedSystemK1(mesh, bdry, D1=1e-3, D2=1.5e-3, degree=4, variant=variant)
        hh.append(mesh.hmax())
        nn.append(system.N)
        u = system.solve(tol=1e-7 if variant in ("galerkin", "c0ip_modified") else 1e-8)
        e = system.errors(u, full_grad)
        es.append(e[0]); ep.append(e[1]); eq.append(e[2])
        if nk > 0:
            lg = math.log(hh[nk] / hh[nk - 1])
            rs.append(math.log(es[nk] / es[nk - 1]) / lg)
            rp.append(math.log(ep[nk] / ep[nk - 1]) / lg)
            rq.append(math.log(eq[nk] / eq[nk - 1]) / lg)
    print('====================================================')
    print('  DoF      h     e_div(s)   r_div(s)   e_0(p)   r_0(p)   e_1(q)  r_1(q)    ')
    print('====================================================')
    rows = ['{:6d} & {:.4f} & {:6.2e} & {:.3f} & {:6.2e} & {:.3f} & {:6.2e} & {:.3f} '.format(
        nn[k], hh[k], es[k], rs[k], ep[k], rp[k], eq[k], rq[k]) for k in range(nkmax)]
    print("\n".join(rows))
    print('====================================================')
    return rows


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 6, sys.argv[2] if len(sys.argv) > 2 else "galerkin")
