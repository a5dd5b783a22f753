"""numpy restatement, statement by statement, of csrc/cm_gmres.cu: bicgstab_solve (right-preconditioned BiCGStab
built from the fused multi-dot / multi-axpy primitives, with s kept in r between the half steps), run on a
Jacobi-preconditioned nonsymmetric convection-diffusion matrix: it converges to the direct solution; the exit at
the half step and the zero right-hand side are covered too.  Host logic only -- the CUDA kernels are exercised by
tests/test_gpu_solver.py."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla


def bicgstab(A, Minv, b, rtol, atol, maxit):
    n = b.size
    x, p, v = np.zeros(n), np.zeros(n), np.zeros(n)
    r, rh = b.copy(), b.copy()
    rho_new = rh @ r
    bnorm = np.sqrt(rho_new)
    target = max(rtol * bnorm, atol)
    rho = alpha = omega = 1.0
    its, resid = 0, bnorm
    converged = not (bnorm > target)
    while not converged and its < maxit:
        assert rho_new != 0.0
        beta = (rho_new / rho) * (alpha / omega)
        rho = rho_new
        p *= beta
        p += 1.0 * r + (-beta * omega) * v            # multi_axpy over (r, v)
        y = Minv(p)
        v = A @ y
        alpha = rho / (rh @ v)
        r -= alpha * v                                # s
        snorm2 = r @ r
        z = Minv(r)
        t = A @ z
        st, tt = r @ t, t @ t                         # multi_dot over (r, t) against t
        its += 1
        if np.sqrt(snorm2) <= target or tt == 0.0:
            x += alpha * y
            resid = np.sqrt(snorm2)
            converged = resid <= target
            break
        omega = st / tt
        x += alpha * y + omega * z                    # multi_axpy over (y, z)
        r -= omega * t
        resid = np.sqrt(r @ r)
        rho_new = rh @ r
        converged = resid <= target
    return x, its, resid / bnorm, converged


def test_bicgstab_restatement_converges_on_a_nonsymmetric_system():
    rng = np.random.default_rng(0)
    n = 400
    # 2D convection-diffusion 5-point matrix (nonsymmetric, diagonally dominant enough for Jacobi)
    m = 20
    I = sp.identity(m)
    T = sp.diags([-1.0 - 0.4, 2.2, -1.0 + 0.4], [-1, 0, 1], shape=(m, m))
    A = (sp.kron(I, T) + sp.kron(sp.diags([-1.0, 2.0, -1.0], [-1, 0, 1], shape=(m, m)), I)).tocsr()
    dinv = 1.0 / A.diagonal()
    b = rng.standard_normal(n)
    x, its, rel, ok = bicgstab(A, lambda r: dinv * r, b, 1e-11, 0.0, 400)
    assert ok and rel <= 1e-11
    xd = spla.spsolve(A.tocsc(), b)
    assert np.linalg.norm(x - xd) / np.linalg.norm(xd) < 1e-9
    assert np.linalg.norm(b - A @ x) / np.linalg.norm(b) < 1e-10       # recurrence residual = true residual
    assert 10 < its < 200


def test_half_step_exit_and_trivial_right_hand_side():
    A = sp.identity(5).tocsr() * 2.0
    b = np.arange(1.0, 6.0)
    x, its, rel, ok = bicgstab(A, lambda r: r, b, 1e-12, 0.0, 10)      # s = 0 after the first half step
    assert ok and its == 1 and np.allclose(x, b / 2.0)
    x, its, rel, ok = bicgstab(A, lambda r: r, np.zeros(5), 1e-12, 0.0, 10)
    assert ok and its == 0 and not x.any()
