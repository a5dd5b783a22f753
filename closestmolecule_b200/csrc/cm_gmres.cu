// Restarted, right-preconditioned GMRES with classical Gram-Schmidt batched into one fused
// multi-dot and one fused multi-axpy(+norm) kernel per iteration -> one host synchronisation per
// iteration (the Hessenberg column).  The operator and the preconditioner are callbacks so the
// same driver serves the generic CSR path and the structured field-split/multigrid path.
// Replaces PETScLUSolver::solve inside NewtonSolver [ext] (advection_diffusion_convergence.py:212).
#include <math.h>

#include <utility>
#include <vector>

#include "cm_common.cuh"
#include "cm_decl.cuh"

namespace cm {

// Pinned staging for the few scalars that cross the bus every iteration: allocated once per host thread and
// grown on demand (cudaMallocHost / cudaFreeHost per solve would cost a device synchronisation each).
static double* pinned_scalars(size_t n) {
  thread_local double* buf = nullptr;
  thread_local size_t cap = 0;
  if (n > cap) {
    if (buf) cudaFreeHost(buf);
    buf = nullptr;
    cap = 0;
    if (cudaMallocHost((void**)&buf, sizeof(double) * n) != cudaSuccess) {
      cudaGetLastError();
      buf = nullptr;
      return nullptr;
    }
    cap = n;
  }
  return buf;
}

size_t gmres_workspace_doubles(long long n, int restart) {
  const long long ne = (n + 1) & ~1LL;  // even stride keeps double2 loads aligned
  return (size_t)ne * (restart + 1)     // Krylov basis
         + (size_t)ne * 2               // w, z
         + (size_t)(restart + 4)        // h on the device
         + (size_t)148 * 4 * 8 + 64;    // reduction scratch
}

int gmres_solve(long long n, const LinOp& A, const LinOp& Minv, const double* b, double* x,
                double rtol, double atol, int maxit, int restart, int reorth, double* work,
                int* iters_out, double* resid_out, double* bnorm_out, cudaStream_t st,
                const VecSlice* sl, cm_dist* D) {
  // single GPU: the whole vector [0,n); multi GPU: this rank's slices + all-reduced dot products
  auto multi_dot = [&](long long, const double* Vb, long long vs, int nvec, const double* ww,
                       double* out, double* scr, cudaStream_t s_) -> int {
    int e = sl ? multi_dot_slice(*sl, Vb, vs, nvec, ww, out, scr, s_)
               : cm::multi_dot(n, Vb, vs, nvec, ww, out, scr, s_);
    if (e) return e;
    return dist_allreduce(D, out, nvec, s_);
  };
  auto multi_axpy = [&](long long, const double* Vb, long long vs, int nvec, const double* hh,
                        double sign, double* ww, double* n2, double* scr, cudaStream_t s_) -> int {
    int e = sl ? multi_axpy_slice(*sl, Vb, vs, nvec, hh, sign, ww, n2, scr, s_)
               : cm::multi_axpy(n, Vb, vs, nvec, hh, sign, ww, n2, scr, s_);
    if (e) return e;
    return n2 ? dist_allreduce(D, n2, 1, s_) : CM_OK;
  };
  auto vec_scale = [&](long long, const double* xx, const double*, double a, int inv, double* yy,
                       cudaStream_t s_) -> int {
    return sl ? vec_scale_slice(*sl, xx, a, inv, yy, s_) : cm::vec_scale(n, xx, nullptr, a, inv, yy, s_);
  };
  auto vec_zero = [&](double* yy) -> int {
    if (sl) return vec_scale_slice(*sl, nullptr, 0.0, 0, yy, st);
    return cudaMemsetAsync(yy, 0, sizeof(double) * n, st) == cudaSuccess ? CM_OK : CM_E_CUDA;
  };
  auto vec_copy = [&](double* yy, const double* xx) -> int {
    if (sl) return vec_scale_slice(*sl, xx, 1.0, 0, yy, st);
    return cudaMemcpyAsync(yy, xx, sizeof(double) * n, cudaMemcpyDeviceToDevice, st) == cudaSuccess
               ? CM_OK : CM_E_CUDA;
  };
  const long long ne = (n + 1) & ~1LL;
  double* V = work;
  double* w = V + (size_t)ne * (restart + 1);
  double* z = w + ne;
  double* hdev = z + ne;
  double* scratch = hdev + (restart + 4);
  scratch = (double*)(((uintptr_t)scratch + 15) & ~(uintptr_t)15);

  double* hhost = pinned_scalars((size_t)restart + 4);  // staging for the Hessenberg column / coefficients
  if (!hhost) {
    set_error("gmres: cudaMallocHost failed");
    return CM_E_CUDA;
  }

  const int m = restart;
  std::vector<double> H((size_t)(m + 1) * m, 0.0), cs(m), sn(m), g(m + 1), y(m);
  auto Hat = [&](int i, int j) -> double& { return H[(size_t)j * (m + 1) + i]; };

  // x0 = 0: r0 = b
  int rc = vec_zero(x);
  if (rc) return rc;
  if ((rc = vec_copy(w, b))) return rc;
  rc = multi_dot(n, w, ne, 1, w, hdev, scratch, st);
  if (rc) return rc;
  CM_CUDA(cudaMemcpyAsync(hhost, hdev, sizeof(double), cudaMemcpyDeviceToHost, st));
  CM_CUDA(cudaStreamSynchronize(st));
  const double bnorm = sqrt(hhost[0]);
  if (bnorm_out) *bnorm_out = bnorm;
  const double target = fmax(rtol * bnorm, atol);
  double beta = bnorm, resid = bnorm;
  int its = 0;
  bool converged = !(bnorm > target);
  if (!isfinite(bnorm)) {
    set_error("gmres: right-hand side norm is not finite");
    return CM_E_BREAKDOWN;
  }

  while (!converged && its < maxit) {
    // V0 = r / beta  (r is in w)
    rc = vec_scale(n, w, nullptr, beta, 1, V, st);
    if (rc) return rc;
    for (int i = 0; i <= m; ++i) g[i] = 0.0;
    g[0] = beta;
    int j = 0;
    for (; j < m && its < maxit; ++j) {
      double* vj = V + (size_t)ne * j;
      rc = Minv(vj, z);
      if (rc) return rc;
      rc = A(z, w);
      if (rc) return rc;
      rc = multi_dot(n, V, ne, j + 1, w, hdev, scratch, st);
      if (rc) return rc;
      rc = multi_axpy(n, V, ne, j + 1, hdev, -1.0, w, hdev + (j + 1), scratch, st);
      if (rc) return rc;
      CM_CUDA(cudaMemcpyAsync(hhost, hdev, sizeof(double) * (j + 2), cudaMemcpyDeviceToHost, st));
      CM_CUDA(cudaStreamSynchronize(st));
      for (int i = 0; i <= j; ++i) Hat(i, j) = hhost[i];
      if (reorth) {  // second classical Gram-Schmidt pass ("twice is enough")
        rc = multi_dot(n, V, ne, j + 1, w, hdev, scratch, st);
        if (rc) return rc;
        rc = multi_axpy(n, V, ne, j + 1, hdev, -1.0, w, hdev + (j + 1), scratch, st);
        if (rc) return rc;
        CM_CUDA(cudaMemcpyAsync(hhost, hdev, sizeof(double) * (j + 2), cudaMemcpyDeviceToHost, st));
        CM_CUDA(cudaStreamSynchronize(st));
        for (int i = 0; i <= j; ++i) Hat(i, j) += hhost[i];
      }
      const double hn = sqrt(fmax(hhost[j + 1], 0.0));
      if (!isfinite(hn)) {
        set_error("gmres: NaN/Inf in the Arnoldi process at iteration %d", its);
        return CM_E_BREAKDOWN;
      }
      Hat(j + 1, j) = hn;
      // previous rotations, then a new one
      for (int i = 0; i < j; ++i) {
        const double t = cs[i] * Hat(i, j) + sn[i] * Hat(i + 1, j);
        Hat(i + 1, j) = -sn[i] * Hat(i, j) + cs[i] * Hat(i + 1, j);
        Hat(i, j) = t;
      }
      const double d = hypot(Hat(j, j), Hat(j + 1, j));
      cs[j] = (d == 0.0) ? 1.0 : Hat(j, j) / d;
      sn[j] = (d == 0.0) ? 0.0 : Hat(j + 1, j) / d;
      Hat(j, j) = d;
      Hat(j + 1, j) = 0.0;
      g[j + 1] = -sn[j] * g[j];
      g[j] = cs[j] * g[j];
      resid = fabs(g[j + 1]);
      ++its;
      const bool lucky = hn <= 1e-300;
      if (resid <= target || lucky) {
        converged = true;
        ++j;
        break;
      }
      if (j + 1 < m || true) {  // next basis vector
        rc = vec_scale(n, w, nullptr, hn, 1, V + (size_t)ne * (j + 1), st);
        if (rc) return rc;
      }
    }
    // y = H^{-1} g (upper triangular, j columns); x += Minv(V y)
    const int k = j;
    for (int i = k - 1; i >= 0; --i) {
      double s = g[i];
      for (int l = i + 1; l < k; ++l) s -= Hat(i, l) * y[l];
      y[i] = s / Hat(i, i);
    }
    for (int i = 0; i < k; ++i) hhost[i] = y[i];
    CM_CUDA(cudaMemcpyAsync(hdev, hhost, sizeof(double) * k, cudaMemcpyHostToDevice, st));
    if ((rc = vec_zero(w))) return rc;
    rc = multi_axpy(n, V, ne, k, hdev, 1.0, w, nullptr, scratch, st);
    if (rc) return rc;
    rc = Minv(w, z);
    if (rc) return rc;
    hhost[m + 2] = 1.0;
    CM_CUDA(cudaMemcpyAsync(hdev + m + 2, hhost + m + 2, sizeof(double), cudaMemcpyHostToDevice, st));
    rc = multi_axpy(n, z, ne, 1, hdev + m + 2, 1.0, x, nullptr, scratch, st);
    if (rc) return rc;
    CM_CUDA(cudaStreamSynchronize(st));
    if (converged || its >= maxit) break;
    // restart: r = b - A x
    rc = A(x, z);
    if (rc) return rc;
    if ((rc = vec_copy(w, b))) return rc;
    hhost[m + 3] = 1.0;
    CM_CUDA(cudaMemcpyAsync(hdev + m + 3, hhost + m + 3, sizeof(double), cudaMemcpyHostToDevice, st));
    rc = multi_axpy(n, z, ne, 1, hdev + m + 3, -1.0, w, hdev, scratch, st);
    if (rc) return rc;
    CM_CUDA(cudaMemcpyAsync(hhost, hdev, sizeof(double), cudaMemcpyDeviceToHost, st));
    CM_CUDA(cudaStreamSynchronize(st));
    beta = sqrt(hhost[0]);
    resid = beta;
    if (beta <= target) converged = true;
  }
  if (iters_out) *iters_out = its;
  if (resid_out) *resid_out = resid;
  return converged ? CM_OK : 1;  // 1 = not converged within maxit (caller decides)
}

// ------------------------------------------------------------------------------------------------
// Right-preconditioned BiCGStab (van der Vorst) on the same fused vector kernels: the 'bicgstab' choice of
// parameters['newton_solver']['linear_solver'] in DOLFIN.  x0 = 0.  Seven work vectors laid out with one
// stride so that the fused multi-vector kernels can address pairs of them:
//   0 rhat | 1 r (holds s between the two half steps) | 2 p | 3 v | 4 t | 5 y = M^-1 p | 6 z = M^-1 s | 7 p'
// (p alternates between the slots 2 and 7 so that no kernel updates a vector in place through two pointers).
// Three host synchronisations per iteration (alpha; omega; rho and ||r||).  One iteration = two operator and two
// preconditioner applications.
size_t bicgstab_workspace_doubles(long long n) {
  const long long ne = (n + 1) & ~1LL;
  return (size_t)ne * 8 + 16 + (size_t)148 * 4 * 8 + 64;
}

int bicgstab_solve(long long n, const LinOp& A, const LinOp& Minv, const double* b, double* x, double rtol,
                   double atol, int maxit, double* work, int* iters_out, double* resid_out,
                   double* bnorm_out, cudaStream_t st) {
  const long long ne = (n + 1) & ~1LL;
  double* rh = work;
  double* r = rh + ne;
  double* p = r + ne;
  double* v = p + ne;
  double* t = v + ne;
  double* y = t + ne;
  double* z = y + ne;
  double* p2 = z + ne;
  double* hdev = p2 + ne;  // 16 scalars
  double* scratch = hdev + 16;
  scratch = (double*)(((uintptr_t)scratch + 15) & ~(uintptr_t)15);

  double* hhost = pinned_scalars(16);
  if (!hhost) {
    set_error("bicgstab: cudaMallocHost failed");
    return CM_E_CUDA;
  }
  auto put = [&](int slot, int cnt) -> int {  // hhost[slot..slot+cnt) -> hdev (same slots)
    CM_CUDA(cudaMemcpyAsync(hdev + slot, hhost + slot, sizeof(double) * cnt, cudaMemcpyHostToDevice, st));
    return CM_OK;
  };
  auto get = [&](int slot, int cnt) -> int {  // hdev -> hhost, then wait
    CM_CUDA(cudaMemcpyAsync(hhost + slot, hdev + slot, sizeof(double) * cnt, cudaMemcpyDeviceToHost, st));
    CM_CUDA(cudaStreamSynchronize(st));
    return CM_OK;
  };
  const size_t vb = sizeof(double) * (size_t)n;
  int rc;
  CM_CUDA(cudaMemsetAsync(x, 0, vb, st));
  CM_CUDA(cudaMemsetAsync(p, 0, vb, st));
  CM_CUDA(cudaMemsetAsync(v, 0, vb, st));
  CM_CUDA(cudaMemcpyAsync(r, b, vb, cudaMemcpyDeviceToDevice, st));
  CM_CUDA(cudaMemcpyAsync(rh, b, vb, cudaMemcpyDeviceToDevice, st));
  if ((rc = multi_dot(n, rh, ne, 1, r, hdev + 0, scratch, st))) return rc;  // rho = <rhat, r> = ||b||^2
  if ((rc = get(0, 1))) return rc;
  const double bnorm = sqrt(hhost[0]);
  if (bnorm_out) *bnorm_out = bnorm;
  if (!isfinite(bnorm)) {
    set_error("bicgstab: right-hand side norm is not finite");
    return CM_E_BREAKDOWN;
  }
  const double target = fmax(rtol * bnorm, atol);
  double rho_new = hhost[0], rho = 1.0, alpha = 1.0, omega = 1.0, resid = bnorm;
  int its = 0;
  bool converged = !(bnorm > target);
  while (!converged && its < maxit) {
    if (rho_new == 0.0 || !isfinite(rho_new)) {
      set_error("bicgstab: breakdown (<rhat, r> = %g) at iteration %d", rho_new, its);
      return CM_E_BREAKDOWN;
    }
    const double beta = (rho_new / rho) * (alpha / omega);
    rho = rho_new;
    // p = r + beta (p - omega v)
    if ((rc = vec_scale(n, p, nullptr, beta, 0, p2, st))) return rc;
    std::swap(p, p2);
    hhost[1] = 1.0;
    hhost[2] = -beta * omega;
    if ((rc = put(1, 2))) return rc;
    if ((rc = multi_axpy(n, r, 2 * ne, 2, hdev + 1, 1.0, p, nullptr, scratch, st))) return rc;  // V = (r, v)
    if ((rc = Minv(p, y))) return rc;
    if ((rc = A(y, v))) return rc;
    if ((rc = multi_dot(n, rh, ne, 1, v, hdev + 3, scratch, st))) return rc;
    if ((rc = get(3, 1))) return rc;
    if (hhost[3] == 0.0 || !isfinite(hhost[3])) {
      set_error("bicgstab: breakdown (<rhat, A M^-1 p> = %g) at iteration %d", hhost[3], its);
      return CM_E_BREAKDOWN;
    }
    alpha = rho / hhost[3];
    // s = r - alpha v (kept in r), ||s||^2 -> hdev[5]
    hhost[4] = alpha;
    if ((rc = put(4, 1))) return rc;
    if ((rc = multi_axpy(n, v, ne, 1, hdev + 4, -1.0, r, hdev + 5, scratch, st))) return rc;
    if ((rc = Minv(r, z))) return rc;
    if ((rc = A(z, t))) return rc;
    // <s, t>, <t, t>: V = (r, t) with stride 3 ne, against t
    if ((rc = multi_dot(n, r, 3 * ne, 2, t, hdev + 6, scratch, st))) return rc;
    if ((rc = get(5, 3))) return rc;
    ++its;
    const double snorm = sqrt(fmax(hhost[5], 0.0)), st_ = hhost[6], tt = hhost[7];
    if (snorm <= target || tt == 0.0) {  // converged at the half step: x += alpha y
      hhost[8] = alpha;
      if ((rc = put(8, 1))) return rc;
      if ((rc = multi_axpy(n, y, ne, 1, hdev + 8, 1.0, x, nullptr, scratch, st))) return rc;
      CM_CUDA(cudaStreamSynchronize(st));
      resid = snorm;
      converged = snorm <= target;
      if (!converged) {
        set_error("bicgstab: breakdown (A M^-1 s = 0 with ||s|| = %g) at iteration %d", snorm, its);
        return CM_E_BREAKDOWN;
      }
      break;
    }
    omega = st_ / tt;
    if (omega == 0.0 || !isfinite(omega)) {
      set_error("bicgstab: breakdown (omega = %g) at iteration %d", omega, its);
      return CM_E_BREAKDOWN;
    }
    // x += alpha y + omega z;  r = s - omega t, ||r||^2 -> hdev[11];  rho_new = <rhat, r> -> hdev[12]
    hhost[8] = alpha;
    hhost[9] = omega;
    hhost[10] = omega;
    if ((rc = put(8, 3))) return rc;
    if ((rc = multi_axpy(n, y, ne, 2, hdev + 8, 1.0, x, nullptr, scratch, st))) return rc;  // V = (y, z)
    if ((rc = multi_axpy(n, t, ne, 1, hdev + 10, -1.0, r, hdev + 11, scratch, st))) return rc;
    if ((rc = multi_dot(n, rh, ne, 1, r, hdev + 12, scratch, st))) return rc;
    if ((rc = get(11, 2))) return rc;
    resid = sqrt(fmax(hhost[11], 0.0));
    rho_new = hhost[12];
    if (!isfinite(resid)) {
      set_error("bicgstab: NaN/Inf residual at iteration %d", its);
      return CM_E_BREAKDOWN;
    }
    if (resid <= target) converged = true;
  }
  if (iters_out) *iters_out = its;
  if (resid_out) *resid_out = resid;
  return converged ? CM_OK : 1;
}

}  // namespace cm
