"""numpy restatement of the device-resident flexible GMRES of csrc/cm_gmres_dev.cu / cm_pqsolver.cu (CPU, no GPU):
right preconditioning on the row-equilibrated system  D A x = D b,  classical Gram-Schmidt on UNNORMALISED basis
vectors (the scale 1/||V_j|| is applied lazily), the next preconditioner input  D^-1 V_{j+1}  written unnormalised by
the update step (the preconditioner and the operator are linear: the operator step applies 1/||V_j||), Hessenberg
column + Givens rotations + residual estimate per iteration, x = Z y from the kept preconditioned vectors, restart
with the true residual.  The GPU kernels are compared with the CPU oracle through the Newton solves of
tests/test_gpu_solver.py; this file pins the ALGORITHM the kernels implement."""
import numpy as np


def fgmres_unnormalised(A, b, Minv, d, rtol, restart, maxit):
    """A: matrix, d: row scaling (D = diag(d)), Minv(r) ~ A^-1 r on the UNSCALED operator.  Returns x, iterations."""
    n = b.size
    As = d[:, None] * A
    bs = d * b
    x = np.zeros(n)
    w = bs.copy()                                   # r0 = b - A x0, x0 = 0
    its = 0
    target = None
    while True:
        m = restart
        V = np.zeros((m + 1, n))
        Z = np.zeros((m, n))
        vscale = np.zeros(m + 1)
        H = np.zeros((m + 1, m))
        cs, sn, g = np.zeros(m), np.zeros(m), np.zeros(m + 1)
        # start of a cycle (j = -1): V_0 = w unnormalised, ru = D^-1 V_0
        V[0] = w
        ru = V[0] / d
        hn = np.linalg.norm(V[0])
        vscale[0] = 1.0 / hn if hn > 0 else 0.0
        g[0] = hn
        if target is None:
            target = rtol * hn
        resid = hn
        j = 0
        while resid > target and its < maxit and j < m:
            zprime = Minv(ru)                       # preconditioner applied to the unnormalised vector
            xs = vscale[j]
            Z[j] = xs * zprime                      # operator step: normalisation applied to the kept vector ...
            w = As @ (xs * zprime)                  # ... and to the operator result
            dots = np.array([vscale[k] * (V[k] @ w) for k in range(j + 1)])
            vnew = w - sum(dots[k] * vscale[k] * V[k] for k in range(j + 1))
            V[j + 1] = vnew
            ru = vnew / d                           # fused into the update kernel
            hn = np.linalg.norm(vnew)
            vscale[j + 1] = 1.0 / hn if hn > 0 else 0.0
            Hc = np.zeros(m + 1)
            Hc[:j + 1] = dots
            Hc[j + 1] = hn
            for i in range(j):
                t = cs[i] * Hc[i] + sn[i] * Hc[i + 1]
                Hc[i + 1] = -sn[i] * Hc[i] + cs[i] * Hc[i + 1]
                Hc[i] = t
            dd = np.hypot(Hc[j], Hc[j + 1])
            c, s = (1.0, 0.0) if dd == 0 else (Hc[j] / dd, Hc[j + 1] / dd)
            cs[j], sn[j] = c, s
            Hc[j], Hc[j + 1] = dd, 0.0
            g[j + 1] = -s * g[j]
            g[j] = c * g[j]
            H[:, j] = Hc
            resid = abs(g[j + 1])
            its += 1
            j += 1
        y = np.zeros(j)
        for i in range(j - 1, -1, -1):
            y[i] = (g[i] - H[i, i + 1:j] @ y[i + 1:j]) / H[i, i]
        x = x + y @ Z[:j]                            # no further preconditioner application
        if resid <= target or its >= maxit:
            return x, its
        w = bs - As @ x                              # restart from the true residual


def _problem(n, seed):
    rng = np.random.default_rng(seed)
    A = np.diag(2.0 + rng.random(n)) + 0.3 * np.diag(rng.random(n - 1), 1) - 0.4 * np.diag(rng.random(n - 1), -1)
    A[::7] *= 1e-6                                   # badly scaled rows: what the equilibration is for
    b = A @ rng.standard_normal(n)
    d = 1.0 / np.max(np.abs(A), axis=1)
    return A, b, d


def test_flexible_gmres_with_unnormalised_basis_converges_to_the_direct_solution():
    A, b, d = _problem(60, 3)
    T = np.tril(A)                                   # an inexact "preconditioner": lower-triangular solve
    x, its = fgmres_unnormalised(A, b, lambda r: np.linalg.solve(T, r), d, 1e-12, 40, 200)
    ref = np.linalg.solve(A, b)
    assert its < 40
    assert np.linalg.norm(x - ref) / np.linalg.norm(ref) < 1e-9


def test_restart_path_and_scale_invariance():
    A, b, d = _problem(80, 5)
    Dinv = 1.0 / np.diag(A)
    x, its = fgmres_unnormalised(A, b, lambda r: Dinv * r, d, 1e-11, 5, 400)      # restart = 5: many cycles
    ref = np.linalg.solve(A, b)
    assert its > 5
    assert np.linalg.norm(x - ref) / np.linalg.norm(ref) < 1e-8
    # the iteration is invariant under a rescaling of the right-hand side (the lazy normalisation is exact)
    x2, its2 = fgmres_unnormalised(A, 1e6 * b, lambda r: Dinv * r, d, 1e-11, 5, 400)
    assert its2 == its
    assert np.linalg.norm(x2 / 1e6 - x) / np.linalg.norm(x) < 1e-10
