"""CPU probe (numpy / scipy, uses the oracle): can x-line solves along r2 precondition the unstabilised Galerkin
q-equation (advection_diffusion_convergence.py:203; manufactured variant ...MixedPrimal_convergence.py:120-123)?
Field-split GMRES with an exact p-block solve and, for the q-block: the exact block, exact x-line solves as block
Jacobi, and 2 / 4 zebra sweeps of exact x-line solves.  Result (profiles/README.md, round 2): SUPG converges in
8 / 5 iterations with 2 / 4 zebra sweeps at every size; for the Galerkin q-block NO line-based variant converges
(300 iterations, residual stagnates at 1e-3 ... diverges), i.e. pivoted line solves alone do not give a Krylov
path for the reference's own Jacobian: the coupling between neighbouring lines is as strong as the line part.
Not collected by pytest (no test_ prefix); run: python tests/probes/galerkin_qblock_probe.py"""
import sys, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))))
import numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
from oracle import fem, pq
from oracle.newton import apply_bc_matrix_vec, apply_bc_residual, merge_bcs

def build(nx, ny, variant):
    mesh = fem.rectangle_mesh(0., 0., 1., 1., nx, ny)
    if variant == 'galerkin':
        prm = pq.PQParams(D1=1e-3, D2=1.5e-3, gkind=pq.G_POWER, gcoef=1.5e-3*4*np.pi, q_react=1.0)
        qside = 1.0
    elif variant == 'prod':   # no q reaction term at all
        prm = pq.PQParams(D1=1e-3, D2=1.5e-3, gkind=pq.G_POWER, gcoef=1.5e-3*4*np.pi, q_react=0.0)
        qside = 1.0
    else:
        prm = pq.PQParams(D1=1e-3, D2=1.5e-3, gkind=pq.G_POWER, gcoef=1.5e-3*4*np.pi, q_react=1.0, q_x2p=0.0, q_x2q=1.0, supg_stab=1.0)
        qside = 0.0
    S = pq.PQSystem(mesh, prm)
    X = mesh.coords
    fc = mesh.facets()
    bv = np.unique(fc.verts[fc.exterior])
    qv = np.nonzero(fem.near(X[:, 0], qside))[0]
    dofs, gv = merge_bcs([(2*bv, pq.p_exact(X[bv,0], X[bv,1])), (2*qv+1, pq.q_exact(X[qv,0], X[qv,1]))])
    u = np.empty(S.N); u[0::2] = pq.p_exact(X[:,0], X[:,1]); u[1::2] = pq.q_exact(X[:,0], X[:,1])
    u *= (1 + 1e-3*np.sin(7*np.repeat(X[:,0],2)))
    vals, b = S.assemble(u, True, None, None)
    apply_bc_matrix_vec(S.rowptr, S.col, vals, dofs)
    apply_bc_residual(b, u, dofs, gv)
    A = sp.csr_matrix((vals, S.col, S.rowptr), shape=(S.N, S.N))
    return mesh, A, b

def run(nx, ny, variant):
    mesh, A, b = build(nx, ny, variant)
    V = A.shape[0]//2
    n1 = nx+1
    P = sp.csr_matrix((np.ones(2*V), (np.arange(2*V), np.r_[np.arange(0,2*V,2), np.arange(1,2*V,2)])), shape=(2*V,2*V))  # split ordering
    perm = np.r_[np.arange(0,2*V,2), np.arange(1,2*V,2)]
    As = A[perm][:, perm].tocsr()
    App, Apq, Aqp, Aqq = As[:V,:V], As[:V,V:], As[V:,:V], As[V:,V:]
    # row scaling
    s = 1.0/np.abs(As).max(axis=1).toarray().ravel()
    Asc = sp.diags(s) @ As
    bs = s*b[perm]
    lup = spla.splu(App.tocsc())
    luq = spla.splu(Aqq.tocsc())
    # line part of Aqq: keep entries within the same grid row
    Aq = Aqq.tocoo()
    same = (Aq.row // n1) == (Aq.col // n1)
    T = sp.csr_matrix((Aq.data[same], (Aq.row[same], Aq.col[same])), shape=Aqq.shape)
    O = (Aqq - T).tocsr()
    luT = spla.splu(T.tocsc())
    rows = np.arange(V)//n1
    def zebra(rq, sweeps):
        x = np.zeros(V)
        for sw in range(sweeps):
            for c in (0,1):
                r = rq - O @ x
                xn = luT.solve(r)   # all lines (block diagonal) -> take colour c
                m = (rows % 2) == c
                x[m] = xn[m]
        return x
    res = {}
    for name, qsolve in (('exact', lambda r: luq.solve(r)), ('lineJac', lambda r: luT.solve(r)), ('zebra2', lambda r: zebra(r,2)), ('zebra4', lambda r: zebra(r,4))):
        def M(r):
            ru = r / s
            return np.r_[lup.solve(ru[:V]), qsolve(ru[V:])]
        its = [0]
        def cb(rk): its[0] += 1
        # right preconditioning via operator A M^-1
        Aop = spla.LinearOperator((2*V,2*V), matvec=lambda y: Asc @ M(y))
        y, info = spla.gmres(Aop, bs, rtol=1e-10, restart=100, maxiter=3, callback=cb, callback_type='pr_norm')
        x = M(y)
        rel = np.linalg.norm(Asc@x - bs)/np.linalg.norm(bs)
        res[name] = (its[0], info, float(rel))
    print(variant, nx, ny, res, flush=True)

for variant in ('supg', 'galerkin', 'prod'):
    for nx, ny in ((64,32),(128,64),(256,128)):
        run(nx, ny, variant)
