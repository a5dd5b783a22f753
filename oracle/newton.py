"""Oracle Newton loop, Dirichlet application and LU solve (test infrastructure).

Restates DOLFIN NewtonSolver::solve + DirichletBC::apply + PETScLUSolver [ext] as
configured at advection_diffusion_convergence.py:207-215,
SphericalAdvectionDiffusionReaction_convergence.py:102-110 and
…MixedPrimal_convergence.py:132-140.  SuperLU stands in for MUMPS/UMFPACK.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def apply_bc_matrix(rowptr, col, vals, bc_dofs):
    """bc.apply(A): zero the row, 1.0 on the diagonal, pattern kept (B.7)."""
    for d in np.asarray(bc_dofs):
        s, e = rowptr[d], rowptr[d + 1]
        vals[s:e] = 0.0
        k = s + np.searchsorted(col[s:e], d)
        vals[k] = 1.0


def apply_bc_matrix_vec(rowptr, col, vals, bc_dofs):
    """Vectorised variant of apply_bc_matrix for large meshes (same result)."""
    bc_dofs = np.asarray(bc_dofs, dtype=np.int64)
    if bc_dofs.size == 0:
        return
    from .fem import _ranges, csr_slots
    s = rowptr[bc_dofs].astype(np.int64)
    l = (rowptr[bc_dofs + 1] - rowptr[bc_dofs]).astype(np.int64)
    vals[_ranges(s, l)] = 0.0
    vals[csr_slots(rowptr, col, bc_dofs, bc_dofs)] = 1.0


def apply_bc_residual(b, x, bc_dofs, bc_vals):
    """bc.apply(b, x): b[dof] = x[dof] - g (B.7)."""
    b[bc_dofs] = x[bc_dofs] - bc_vals


def merge_bcs(bcs):
    """BCs are applied in list order; the last one wins on shared dofs (B.7;
    advection_diffusion_convergence.py:183).  bcs = [(dofs, values), ...] -> (dofs, values) unique."""
    d = {}
    for dofs, vals in bcs:
        for k, v in zip(np.asarray(dofs).tolist(), np.broadcast_to(vals, np.shape(dofs)).tolist()):
            d[k] = v
    keys = np.array(sorted(d), dtype=np.int64)
    return keys, np.array([d[k] for k in keys], dtype=np.float64)


def newton_solve(assemble, x0, rowptr, col, bc_dofs, bc_vals, atol=1e-10, rtol=1e-9, maxit=50,
                 relax=1.0, verbose=False, return_history=False):
    """assemble(x, want_jac) -> (vals or None, F).  Returns (x, iterations, converged[, history])."""
    x = x0.copy()
    n = x.size
    _, b = assemble(x, False)
    apply_bc_residual(b, x, bc_dofs, bc_vals)
    r0 = float(np.linalg.norm(b))
    r = r0
    hist = [r0]
    it = 0
    conv = r < atol or (r0 > 0 and r / r0 < rtol)
    while not conv and it < maxit:
        vals, _ = assemble(x, True)
        apply_bc_matrix_vec(rowptr, col, vals, bc_dofs)
        A = sp.csr_matrix((vals, col, rowptr), shape=(n, n))
        dx = spla.splu(A.tocsc()).solve(b)
        x -= relax * dx
        it += 1
        _, b = assemble(x, False)
        apply_bc_residual(b, x, bc_dofs, bc_vals)
        r = float(np.linalg.norm(b))
        hist.append(r)
        if verbose:
            print(f"Newton iteration {it}: r (abs) = {r:.3e} (tol = {atol:.3e}) "
                  f"r (rel) = {r / r0:.3e} (tol = {rtol:.3e})")
        conv = r < atol or r / r0 < rtol
    if return_history:
        return x, it, conv, hist
    return x, it, conv
