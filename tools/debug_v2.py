"""Diagnostics (GPU): cross the two GMRES drivers with the two preconditioner generations on one Jacobian."""
import os
import sys

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_mg import _supg_solver  # noqa: E402

KEYS = ("CM_SOLVER_V1", "CM_COARSE_NCTA", "CM_COARSE_N1", "CM_GRAPH", "CM_GMRES_HOST", "CM_PRECOND_V1")
for nx, ny in ((64, 32), (256, 128)):
    asm, make = _supg_solver(nx, ny)
    rp, col, vals = asm.csr()
    n = asm.b.numel()
    A = sp.csr_matrix((vals.cpu().numpy(), col.cpu().numpy(), rp.cpu().numpy()), shape=(n, n))
    ref = spla.splu(A.tocsc()).solve(asm.b.cpu().numpy())
    for name, env in (("v1", {"CM_SOLVER_V1": "1"}), ("v2", {}), ("v2 gmres_host", {"CM_GMRES_HOST": "1"}),
                      ("v2 precond_v1", {"CM_PRECOND_V1": "1"}),
                      ("v2 gmres_host unfused", {"CM_GMRES_HOST": "1", "CM_COARSE_N1": "0"}),
                      ("v2 gmres_host cta1", {"CM_GMRES_HOST": "1", "CM_COARSE_NCTA": "1"}),
                      ("v2 gmres_host precond_v1", {"CM_GMRES_HOST": "1", "CM_PRECOND_V1": "1"}),
                      ("v2 nograph", {"CM_GRAPH": "0"})):
        for k in KEYS:
            os.environ.pop(k, None)
        os.environ.update(env)
        lin = make()
        lin.setup()
        dx = torch.zeros_like(asm.b)
        try:
            ok = lin.solve(asm.b, dx, 1e-11, 0.0, 200)
        except Exception as e:
            print(nx, name, "EXC", e)
            continue
        err = np.linalg.norm(dx.cpu().numpy() - ref) / np.linalg.norm(ref)
        print(f"{nx:5d} {name:28s} ok={ok} its={lin.last_iterations:3d} resid={lin.last_residual:.2e} err_vs_lu={err:.2e} {lin.info()}")
