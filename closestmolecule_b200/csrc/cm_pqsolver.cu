// Newton linear solve J dx = F on structured meshes: GMRES on the row-equilibrated system,
// right-preconditioned by the block-lower-triangular field split  [[App, 0], [Aqp, Aqq]]:
//   p-block: one geometric-multigrid V(1,1) cycle with zebra x-line relaxation (Galerkin coarse
//            operators in stencil form, dense inverse on the coarsest grid);
//   q-block: two zebra x-line relaxation sweeps (the q-equation is transport along r2).
// All of it is caller-provided workspace + kernels; nothing is kept between calls except the
// workspace content written by cm_pqs_setup.  Replaces the sparse LU the reference requests with
// linear_solver='mumps'|'umfpack' (advection_diffusion_convergence.py:212,
// …MixedPrimal_convergence_SUPG.py:137).  The generic Jacobi-GMRES for CSR matrices is here too.
#include <vector>

#include "cm_common.cuh"
#include "cm_decl.cuh"

namespace cm {

constexpr int kMaxLevels = 24;
constexpr int kCoarseDenseMax = 256;

struct PQSLevel {
  int nx, ny;
  size_t n;
  double *A, *lf, *iu, *cu, *x, *b, *r;
};

struct PQSolver {
  int nx, ny, nv, nlev, restart;
  int q_sweeps = 2;  // zebra sweeps on the q-block per application
  bool dense_coarse;
  PQSLevel lev[kMaxLevels];
  double *PP, *PQ, *QP, *QQ, *sp, *sq;
  double *qlf, *qiu, *qcu;
  double *Ainv, *Mwork;
  double *bs, *xs, *tq, *ru;
  double* gm;
  int* err;
  size_t total_doubles;
};

static inline size_t even(size_t n) { return (n + 1) & ~(size_t)1; }

static int pqs_layout(PQSolver& S, int nx, int ny, int restart, double* base) {
  S.nx = nx;
  S.ny = ny;
  S.nv = (nx + 1) * (ny + 1);
  S.restart = restart;
  size_t off = 0;
  auto take = [&](size_t n) {
    double* p = base ? base + off : nullptr;
    off += even(n);
    return p;
  };
  const size_t V = (size_t)S.nv;
  S.PP = take(7 * V);
  S.PQ = take(7 * V);
  S.QP = take(7 * V);
  S.QQ = take(7 * V);
  S.sp = take(V);
  S.sq = take(V);
  S.qlf = take(V);
  S.qiu = take(V);
  S.qcu = take(V);
  S.bs = take(2 * V);
  S.xs = take(2 * V);
  S.tq = take(V);
  S.ru = take(2 * V);
  int lx = nx, ly = ny, l = 0;
  for (;;) {
    if (l >= kMaxLevels) return CM_E_UNSUPPORTED;
    PQSLevel& L = S.lev[l];
    L.nx = lx;
    L.ny = ly;
    L.n = (size_t)(lx + 1) * (ly + 1);
    L.A = (l == 0) ? S.PP : take(7 * L.n);
    L.lf = take(L.n);
    L.iu = take(L.n);
    L.cu = take(L.n);
    L.r = take(L.n);
    L.x = (l == 0) ? nullptr : take(L.n);
    L.b = (l == 0) ? nullptr : take(L.n);
    ++l;
    if ((lx & 1) || (ly & 1) || (lx < ly ? lx : ly) <= 4) break;
    lx >>= 1;
    ly >>= 1;
  }
  S.nlev = l;
  const size_t nc = S.lev[l - 1].n;
  S.dense_coarse = nc <= (size_t)kCoarseDenseMax;
  S.Ainv = take(S.dense_coarse ? nc * nc : 2);
  S.Mwork = take(S.dense_coarse ? 2 * nc * nc : 2);
  S.err = (int*)take(4);
  S.gm = take(gmres_workspace_doubles(2 * (long long)V, restart));
  S.total_doubles = off;
  return CM_OK;
}

static int check_err_flag(PQSolver& S, cudaStream_t st, const char* what) {
  int h = 0;
  CM_CUDA(cudaMemcpyAsync(&h, S.err, sizeof(int), cudaMemcpyDeviceToHost, st));
  CM_CUDA(cudaStreamSynchronize(st));
  if (h == 1) {
    set_error("%s: the Jacobian is not a 7-point stencil on the (nx,ny) grid (interior-facet "
              "pattern or unstructured mesh): the structured solver does not apply", what);
    return CM_E_UNSUPPORTED;
  }
  if (h == 2) {
    set_error("%s: zero pivot in an x-line factorisation (unstabilised Galerkin q-block?)", what);
    return CM_E_BREAKDOWN;
  }
  if (h == 3) {
    set_error("%s: singular coarsest-grid operator", what);
    return CM_E_BREAKDOWN;
  }
  return CM_OK;
}

static int zebra_sweep(const PQSLevel& L, const double* A, const double* lf, const double* iu,
                       const double* cu, const double* b, double* x, bool x_zero, cudaStream_t st) {
  int rc = st_zebra(L.nx, L.ny, 0, A, lf, iu, cu, b, x, x_zero ? 1 : 0, st);
  if (rc) return rc;
  return st_zebra(L.nx, L.ny, 1, A, lf, iu, cu, b, x, 0, st);
}

static int vcycle(PQSolver& S, int l, const double* b, double* x, cudaStream_t st) {
  PQSLevel& L = S.lev[l];
  int rc;
  if (l == S.nlev - 1) {
    if (S.dense_coarse) return st_coarse_apply((int)L.n, S.Ainv, b, x, st);
    rc = zebra_sweep(L, L.A, L.lf, L.iu, L.cu, b, x, true, st);
    for (int s = 1; s < 30 && !rc; ++s) rc = zebra_sweep(L, L.A, L.lf, L.iu, L.cu, b, x, false, st);
    return rc;
  }
  PQSLevel& C = S.lev[l + 1];
  if ((rc = zebra_sweep(L, L.A, L.lf, L.iu, L.cu, b, x, true, st))) return rc;
  if ((rc = st_residual(L.nx, L.ny, L.A, b, x, L.r, st))) return rc;
  if ((rc = st_restrict(C.nx, C.ny, L.r, C.b, st))) return rc;
  if ((rc = vcycle(S, l + 1, C.b, C.x, st))) return rc;
  if ((rc = st_prolong_add(C.nx, C.ny, C.x, x, st))) return rc;
  return zebra_sweep(L, L.A, L.lf, L.iu, L.cu, b, x, false, st);
}

size_t pqs_workspace_bytes(int nx, int ny, int restart) {
  PQSolver S;
  if (pqs_layout(S, nx, ny, restart, nullptr)) return 0;
  return S.total_doubles * sizeof(double) + 64;
}

static double* align_ws(void* ws) { return (double*)(((uintptr_t)ws + 63) & ~(uintptr_t)63); }

int pqs_setup(int nx, int ny, int restart, const int* nrowptr, const int* ncol, const double* vals,
              void* ws, size_t ws_bytes, cudaStream_t st) {
  PQSolver S;
  int rc = pqs_layout(S, nx, ny, restart, align_ws(ws));
  if (rc) return rc;
  if (S.total_doubles * sizeof(double) + 64 > ws_bytes) {
    set_error("pqs_setup: workspace too small");
    return CM_E_INVALID;
  }
  CM_CUDA(cudaMemsetAsync(S.err, 0, sizeof(int), st));
  if ((rc = st_extract_blocks(nx, ny, nrowptr, ncol, vals, S.PP, S.PQ, S.QP, S.QQ, S.sp, S.sq, S.err, st)))
    return rc;
  for (int l = 0; l < S.nlev; ++l) {
    PQSLevel& L = S.lev[l];
    if ((rc = st_line_factor(L.nx, L.ny, L.A, L.lf, L.iu, L.cu, S.err, st))) return rc;
    if (l + 1 < S.nlev && (rc = st_rap(S.lev[l + 1].nx, S.lev[l + 1].ny, L.A, S.lev[l + 1].A, st)))
      return rc;
  }
  PQSLevel& C = S.lev[S.nlev - 1];
  if (S.dense_coarse && (rc = st_coarse_invert(C.nx, C.ny, C.A, S.Mwork, S.Ainv, S.err, st))) return rc;
  if ((rc = st_line_factor(nx, ny, S.QQ, S.qlf, S.qiu, S.qcu, S.err, st))) return rc;
  return check_err_flag(S, st, "cm_pqs_setup");
}

int pqs_solve(int nx, int ny, int restart, const double* b, double* x, double rtol, double atol,
              int maxit, void* ws, size_t ws_bytes, int* iters, double* resid, cudaStream_t st) {
  PQSolver S;
  int rc = pqs_layout(S, nx, ny, restart, align_ws(ws));
  if (rc) return rc;
  if (S.total_doubles * sizeof(double) + 64 > ws_bytes) {
    set_error("pqs_solve: workspace too small");
    return CM_E_INVALID;
  }
  const int V = S.nv;
  if ((rc = st_split_scale(V, b, S.sp, S.sq, S.bs, st))) return rc;
  LinOp A = [&](const double* in, double* out) {
    return st_pq_spmv(nx, ny, S.PP, S.PQ, S.QP, S.QQ, S.sp, S.sq, in, out, st);
  };
  // M^{-1} = [[App,0],[Aqp,Aqq]]^{-1} (approximately) on the UNSCALED operator, composed with the
  // inverse row scaling: multigrid needs the true (variational) FE operator to be mesh independent
  LinOp Minv = [&](const double* r, double* z) {
    int e = st_unscale(V, r, S.sp, S.sq, S.ru, st);
    if (e) return e;
    if ((e = vcycle(S, 0, S.ru, z, st))) return e;
    if ((e = st_residual(nx, ny, S.QP, S.ru + V, z, S.tq, st))) return e;
    if ((e = zebra_sweep(S.lev[0], S.QQ, S.qlf, S.qiu, S.qcu, S.tq, z + V, true, st))) return e;
    for (int s = 1; s < S.q_sweeps && !e; ++s)
      e = zebra_sweep(S.lev[0], S.QQ, S.qlf, S.qiu, S.qcu, S.tq, z + V, false, st);
    return e;
  };
  double bnorm = 0.0;
  rc = gmres_solve(2LL * V, A, Minv, S.bs, S.xs, rtol, atol, maxit, restart, 0, S.gm, iters, resid,
                   &bnorm, st);
  if (rc < 0) return rc;
  if (resid && bnorm > 0.0) *resid /= bnorm;
  int rc2 = st_interleave(V, S.xs, x, st);
  if (rc2) return rc2;
  if (rc == 1) {
    set_error("cm_pqs_solve: GMRES did not reach rtol=%g in %d iterations (relative residual %g)",
              rtol, maxit, resid ? *resid : -1.0);
  }
  return rc;
}

// ------------------------------------------------------------------------------------------------
// generic path: Jacobi-preconditioned GMRES on the node-level CSR layouts (bs = 1 or 2)
// ------------------------------------------------------------------------------------------------
__global__ void jacobi_setup_kernel(const int bs, const int nv, const int* __restrict__ nrowptr,
                                    const int* __restrict__ diag_slot, const double* __restrict__ vals,
                                    double* __restrict__ dinv) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  const int rp = nrowptr[v], deg = nrowptr[v + 1] - rp, k = diag_slot[v];
  if (bs == 1) {
    const double d = vals[rp + k];
    dinv[v] = d != 0.0 ? 1.0 / d : 1.0;
  } else {
    // 2x2 node block [[pp,pq],[qp,qq]] inverted (identity when singular)
    const size_t base = 4 * (size_t)rp;
    const double a = vals[base + 2 * k], b = vals[base + 2 * k + 1];
    const double c = vals[base + 2 * deg + 2 * k], d = vals[base + 2 * deg + 2 * k + 1];
    const double det = a * d - b * c;
    const double scale = fmax(fmax(fabs(a), fabs(b)), fmax(fabs(c), fabs(d)));
    if (fabs(det) > 1e-14 * scale * scale && scale > 0.0) {
      const double id = 1.0 / det;
      dinv[4 * (size_t)v] = d * id; dinv[4 * (size_t)v + 1] = -b * id;
      dinv[4 * (size_t)v + 2] = -c * id; dinv[4 * (size_t)v + 3] = a * id;
    } else {
      dinv[4 * (size_t)v] = 1.0; dinv[4 * (size_t)v + 1] = 0.0;
      dinv[4 * (size_t)v + 2] = 0.0; dinv[4 * (size_t)v + 3] = 1.0;
    }
  }
}

__global__ void jacobi_apply_kernel(const int bs, const int nv, const double* __restrict__ dinv,
                                    const double* __restrict__ r, double* __restrict__ z) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  if (bs == 1) {
    z[v] = dinv[v] * r[v];
  } else {
    const double r0 = r[2 * (size_t)v], r1 = r[2 * (size_t)v + 1];
    const double* m = dinv + 4 * (size_t)v;
    z[2 * (size_t)v] = m[0] * r0 + m[1] * r1;
    z[2 * (size_t)v + 1] = m[2] * r0 + m[3] * r1;
  }
}

size_t krylov_workspace_bytes(int bs, int nv, int restart) {
  const long long n = (long long)bs * nv;
  return (gmres_workspace_doubles(n, restart) + (size_t)bs * bs * nv + 8) * sizeof(double) + 64;
}

int krylov_solve(int bs, int nv, const int* nrowptr, const int* ncol, const int* diag_slot,
                 const double* vals, const double* b, double* x, double rtol, double atol, int maxit,
                 int restart, void* ws, size_t ws_bytes, int* iters, double* resid,
                 cudaStream_t st) {
  if (krylov_workspace_bytes(bs, nv, restart) > ws_bytes) {
    set_error("krylov_solve: workspace too small");
    return CM_E_INVALID;
  }
  double* base = align_ws(ws);
  double* dinv = base;
  double* gm = base + even((size_t)bs * bs * nv);
  const int grid = (nv + 255) / 256;
  jacobi_setup_kernel<<<grid, 256, 0, st>>>(bs, nv, nrowptr, diag_slot, vals, dinv);
  CM_LAUNCH_CHECK();
  LinOp A = [&](const double* in, double* out) {
    return spmv(bs, nv, nrowptr, ncol, vals, in, out, 1.0, nullptr, st);
  };
  LinOp Minv = [&](const double* r, double* z) {
    jacobi_apply_kernel<<<grid, 256, 0, st>>>(bs, nv, dinv, r, z);
    ++g_launches;
    return cudaGetLastError() == cudaSuccess ? CM_OK : CM_E_CUDA;
  };
  double bnorm = 0.0;
  int rc = gmres_solve((long long)bs * nv, A, Minv, b, x, rtol, atol, maxit, restart, 0, gm, iters,
                       resid, &bnorm, st);
  if (resid && bnorm > 0.0) *resid /= bnorm;
  if (rc == 1)
    set_error("cm_krylov_solve: GMRES did not reach rtol=%g in %d iterations (relative residual %g)",
              rtol, maxit, resid ? *resid : -1.0);
  return rc;
}

}  // namespace cm
