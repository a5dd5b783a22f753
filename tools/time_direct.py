"""Times the direct Newton linear solve (cm_blu_factor / cm_blu_solve) on the Galerkin P-Q Jacobian of
BASELINE config 2 (…MixedPrimal_convergence.py, up to 1 Mi cells) and, for the smaller meshes, SuperLU
on the same matrix on the host.  usage: python tools/time_direct.py 256 512 1024 [--cpu-max 512]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import closestmolecule_b200 as cm  # noqa: E402
from closestmolecule_b200.manufactured import p_exact, q_exact  # noqa: E402
from closestmolecule_b200.solver import BandedLULinearSolver  # noqa: E402


def run(nx, cpu):
    ny = nx // 2
    mesh = cm.RectangleMesh(cm.Point(0, 0), cm.Point(1, 1), nx, ny)
    P1 = cm.FiniteElement("CG", "triangle", 1)
    V = cm.FunctionSpace(mesh, cm.MixedElement([P1, P1]))
    bd = cm.MeshFunction("size_t", mesh, 1, 0)
    cm.CompiledSubDomain("near(x[0],1) && on_boundary").mark(bd, 31)
    form = cm.PQPrimalForm.manufactured(V, "galerkin")
    bcs = [cm.DirichletBC(V.sub(0), p_exact, "on_boundary"), cm.DirichletBC(V.sub(1), q_exact, bd, 31)]
    s = cm.NonlinearVariationalSolver(cm.NonlinearVariationalProblem(form, cm.Function(V), bcs))
    asm = s.asm
    X = mesh.coords
    u = torch.empty(2 * mesh.num_vertices(), dtype=torch.float64, device=mesh.device)
    u[0::2] = 0.9 * p_exact(X[:, 0], X[:, 1])
    u[1::2] = 1.1 * q_exact(X[:, 0], X[:, 1])
    bc = asm._bc_tables(bcs)
    asm.assemble(u, True)
    asm.apply_bcs(bc, u, True)
    lin = BandedLULinearSolver(asm, refinement_steps=1)
    dx = torch.empty_like(u)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    out = dict(nx=nx, ny=ny, cells=mesh.num_cells(), dofs=u.numel(), nbn=lin.nbn, nblk=-(-mesh.num_vertices() // lin.nbn),
               workspace_GiB=lin.nbytes / 2 ** 30)
    for rep in range(2):
        ev[0].record()
        lin.setup()
        ev[1].record()
        lin.solve(asm.b, dx)
        ev[2].record()
        torch.cuda.synchronize()
        out["factor_ms"], out["solve_refine_ms"] = ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2])
    nb = 2 * lin.nbn
    out["factor_tflops"] = 7.67 * nb ** 3 * out["nblk"] / out["factor_ms"] / 1e9
    out["rel_residual"] = lin.last_residual
    if cpu:
        import scipy.sparse as sp
        import scipy.sparse.linalg as spla
        rp, col, vals = asm.csr()
        n = u.numel()
        A = sp.csr_matrix((vals.cpu().numpy(), col.cpu().numpy(), rp.cpu().numpy()), shape=(n, n)).tocsc()
        b = asm.b.cpu().numpy()
        t0 = time.perf_counter()
        lu = spla.splu(A)
        t1 = time.perf_counter()
        ref = lu.solve(b)
        t2 = time.perf_counter()
        out["superlu_factor_ms"], out["superlu_solve_ms"] = 1e3 * (t1 - t0), 1e3 * (t2 - t1)
        out["rel_diff_vs_superlu"] = float(np.linalg.norm(dx.cpu().numpy() - ref) / np.linalg.norm(ref))
    print(json.dumps(out), flush=True)
    del lin, asm, s
    torch.cuda.empty_cache()


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("nx", type=int, nargs="+", help="mesh is nx x nx/2")
    ap.add_argument("--cpu-max", type=int, default=512, help="largest nx that is also solved by SuperLU on the host")
    a = ap.parse_args()
    for n in a.nx:
        run(n, n <= a.cpu_max)
