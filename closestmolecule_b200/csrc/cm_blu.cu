// Direct Newton linear solve: banded LU with partial pivoting, organised as a block-tridiagonal
// factorisation of the (renumbered) Jacobian.  Replaces PETScLUSolver('mumps' | 'umfpack')
// (advection_diffusion_convergence.py:212, advection_diffusion_mixed.py:223,
// ...MixedPrimal_convergence.py:136) for the systems the multigrid-preconditioned Krylov path cannot
// handle: the unstabilised Galerkin q-equation (zero q-q diagonal, :203) and unstructured meshes.
//
// Nodes are renumbered by `perm` so that coupled nodes are at most nbn positions apart; blocks of nbn
// nodes (nb = bs*nbn dofs) then make the matrix block tridiagonal.  Step k factorises the 2nb x nb panel
// [S_k; A_{k+1,k}] with row pivoting over both block rows (= Gaussian elimination with partial
// pivoting of the band matrix, LAPACK dgbtrf's pivot set), applies the row permutation to the 2nb x 2nb
// block to its right, solves for U_{k,k+1}, U_{k,k+2} and forms the Schur complement of block row k+1.
// The dense panel LU, the triangular solves and the Schur update are library calls (cuSOLVER getrf,
// cuBLAS trsm/gemm: plain dense fp64 kernels); the band scatter from the DOLFIN-layout CSR, the
// pivot-to-permutation conversion, the fused row gather (laswp + relocation) and the solve-phase
// permutations are the kernels below.
#include <cublas_v2.h>
#include <cusolverDn.h>

#include <vector>

#include "cm_common.cuh"
#include "cm_decl.cuh"

namespace cm {

struct BluCtx {
  cusolverDnHandle_t sol = nullptr;
  cublasHandle_t blas = nullptr;
  double* getrf_work = nullptr;
  int getrf_lwork = 0;
};

struct BluLayout {
  int bs, nv, nbn, nb, nblk;
  size_t off_P, off_U, off_src, off_ipiv, off_info, off_orig, off_T0, off_T1, off_y, total;
};

static inline size_t up256(size_t x) { return (x + 255) / 256 * 256; }

static BluLayout blu_layout(int bs, int nv, int nbn) {
  BluLayout L;
  L.bs = bs;
  L.nv = nv;
  L.nbn = nbn;
  L.nb = bs * nbn;
  L.nblk = (nv + nbn - 1) / nbn;
  const size_t nb = (size_t)L.nb, nblk = (size_t)L.nblk;
  size_t o = 0;
  L.off_P = o;      o += up256(nblk * 2 * nb * nb * 8);   // panels [L11\U11; L21], ld 2nb
  L.off_U = o;      o += up256(nblk * 2 * nb * nb * 8);   // [U_{k,k+1} U_{k,k+2}], ld nb
  L.off_src = o;    o += up256(nblk * 2 * nb * 4);        // row permutation of every step
  L.off_ipiv = o;   o += up256(nblk * nb * 4);
  L.off_info = o;   o += up256((nblk + 1) * 4);           // getrf info per step + band-violation flag
  L.off_orig = o;   o += up256(2 * nb * nb * 8);          // [A_{k+1,k+1} A_{k+1,k+2}], ld nb
  L.off_T0 = o;     o += up256(nb * nb * 8);              // updated A_{k,k+1}, double buffered
  L.off_T1 = o;     o += up256(nb * nb * 8);
  L.off_y = o;      o += up256((nblk + 2) * nb * 8);      // permuted right-hand side / solution
  L.total = o;
  return L;
}

size_t blu_workspace_bytes(int bs, int nv, int nbn) { return blu_layout(bs, nv, nbn).total; }

// ---- kernels ---------------------------------------------------------------------------------

// Dense images of block row kb of the renumbered matrix: column block kb-1 -> dL, kb -> dD,
// kb+1 -> dU (column major, leading dimensions ldL/ldD/ldU; a NULL destination skips that block).
// Padding rows of the last block get a unit diagonal.  bs=2 reads the DOLFIN/PETSc value layout
// (scalar row 2v+c starts at 4*rowptr[v] + c*2*deg), bs=1 the node CSR.
__global__ void __launch_bounds__(128)
blu_scatter_kernel(const int bs, const int nv, const int nbn, const int kb,
                   const int* __restrict__ iperm, const int* __restrict__ perm,
                   const int* __restrict__ nrowptr, const int* __restrict__ ncol,
                   const double* __restrict__ vals, double* __restrict__ dL, const int ldL,
                   double* __restrict__ dD, const int ldD, double* __restrict__ dU, const int ldU,
                   int* __restrict__ band_flag) {
  const int lp = blockIdx.x * blockDim.x + threadIdx.x;
  if (lp >= nbn) return;
  const int p = kb * nbn + lp;
  if (p >= nv) {
    if (dD != nullptr)
      for (int c = 0; c < bs; ++c) dD[(size_t)(bs * lp + c) + (size_t)(bs * lp + c) * ldD] = 1.0;
    return;
  }
  const int v = iperm[p];
  const int rp = nrowptr[v];
  const int deg = nrowptr[v + 1] - rp;
  for (int j = 0; j < deg; ++j) {
    const int pw = perm[ncol[rp + j]];
    const int cb = pw / nbn;
    const int lc = pw - cb * nbn;
    double* dst;
    int ld;
    if (cb == kb) {
      dst = dD;
      ld = ldD;
    } else if (cb == kb - 1) {
      dst = dL;
      ld = ldL;
    } else if (cb == kb + 1) {
      dst = dU;
      ld = ldU;
    } else {
      atomicExch(band_flag, 1);  // the renumbering does not have the promised bandwidth
      continue;
    }
    if (dst == nullptr) continue;
    if (bs == 2) {
      const double* base = vals + 4 * (size_t)rp + 2 * j;
#pragma unroll
      for (int c = 0; c < 2; ++c)
#pragma unroll
        for (int d = 0; d < 2; ++d)
          dst[(size_t)(2 * lp + c) + (size_t)(2 * lc + d) * ld] = base[(size_t)c * 2 * deg + d];
    } else {
      dst[(size_t)lp + (size_t)lc * ld] = vals[rp + j];
    }
  }
}

// LAPACK pivot sequence (1-based, n swaps on m rows) -> src[r] = original row that ends up in row r
__global__ void __launch_bounds__(256)
blu_perm_kernel(const int m, const int n, const int* __restrict__ ipiv, int* __restrict__ src) {
  extern __shared__ int sh[];
  int* s = sh;
  int* piv = sh + m;
  for (int i = threadIdx.x; i < m; i += blockDim.x) s[i] = i;
  for (int i = threadIdx.x; i < n; i += blockDim.x) piv[i] = ipiv[i] - 1;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i < n; ++i) {
      const int j = piv[i];
      if (j != i && j >= 0 && j < m) {
        const int t = s[i];
        s[i] = s[j];
        s[j] = t;
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < m; i += blockDim.x) src[i] = s[i];
}

// Row interchanges of step k applied to the 2nb x 2nb block right of the panel, fused with the
// relocation of its two halves: rows [0,nb) -> U12 (ld nb), rows [nb,2nb) -> top of the next panel
// (columns [0,nb), ld 2nb) and the next T (columns [nb,2nb), ld nb).  Sources: rows of block k =
// [T | 0], rows of block k+1 = orig (ld nb).
__global__ void __launch_bounds__(256)
blu_gather_kernel(const int nb, const int* __restrict__ src, const double* __restrict__ T,
                  const double* __restrict__ orig, double* __restrict__ U12,
                  double* __restrict__ Pnext, double* __restrict__ Tnext) {
  const int c = blockIdx.y;  // column of the 2nb-wide block
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < 2 * nb; r += gridDim.x * blockDim.x) {
    const int s = src[r];
    double val;
    if (s < nb)
      val = (c < nb) ? T[(size_t)s + (size_t)c * nb] : 0.0;
    else
      val = orig[(size_t)(s - nb) + (size_t)c * nb];
    if (r < nb)
      U12[(size_t)r + (size_t)c * nb] = val;
    else if (c < nb)
      Pnext[(size_t)(r - nb) + (size_t)c * 2 * nb] = val;
    else
      Tnext[(size_t)(r - nb) + (size_t)(c - nb) * nb] = val;
  }
}

// y[bs*perm[v]+c] = b[bs*v+c]  (dir=0)   or   x[bs*v+c] = y[bs*perm[v]+c]  (dir=1)
__global__ void __launch_bounds__(256)
blu_pack_kernel(const int bs, const int nv, const int* __restrict__ perm, const double* __restrict__ in,
                double* __restrict__ out, const int dir) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= bs * nv) return;
  const int v = i / bs, c = i - v * bs;
  const int j = bs * perm[v] + c;
  if (dir == 0)
    out[j] = in[i];
  else
    out[i] = in[j];
}

// z[0..m) <- z[src[0..m)] in place (one CTA, staged through shared memory)
__global__ void __launch_bounds__(1024)
blu_vecperm_kernel(const int m, const int* __restrict__ src, double* __restrict__ z) {
  extern __shared__ double zs[];
  for (int i = threadIdx.x; i < m; i += blockDim.x) zs[i] = z[i];
  __syncthreads();
  for (int i = threadIdx.x; i < m; i += blockDim.x) z[i] = zs[src[i]];
}

// ---- host orchestration ------------------------------------------------------------------------

#define CM_SOLVER(call)                                                                     \
  do {                                                                                      \
    cusolverStatus_t s_ = (call);                                                           \
    if (s_ != CUSOLVER_STATUS_SUCCESS) {                                                    \
      cm::set_error("%s: cuSOLVER status %d at %s:%d", __func__, (int)s_, __FILE__, __LINE__); \
      return CM_E_CUDA;                                                                     \
    }                                                                                       \
  } while (0)
#define CM_BLAS(call)                                                                       \
  do {                                                                                      \
    cublasStatus_t s_ = (call);                                                             \
    if (s_ != CUBLAS_STATUS_SUCCESS) {                                                      \
      cm::set_error("%s: cuBLAS status %d at %s:%d", __func__, (int)s_, __FILE__, __LINE__);   \
      return CM_E_CUDA;                                                                     \
    }                                                                                       \
  } while (0)

void* blu_create() {
  BluCtx* c = new BluCtx();
  if (cusolverDnCreate(&c->sol) != CUSOLVER_STATUS_SUCCESS) {
    set_error("cm_blu_create: cusolverDnCreate failed");
    delete c;
    return nullptr;
  }
  if (cublasCreate(&c->blas) != CUBLAS_STATUS_SUCCESS) {
    set_error("cm_blu_create: cublasCreate failed");
    cusolverDnDestroy(c->sol);
    delete c;
    return nullptr;
  }
  cublasSetPointerMode(c->blas, CUBLAS_POINTER_MODE_HOST);
  return c;
}

int blu_destroy(void* ctx) {
  BluCtx* c = (BluCtx*)ctx;
  if (c == nullptr) return CM_OK;
  if (c->getrf_work) cudaFree(c->getrf_work);
  if (c->blas) cublasDestroy(c->blas);
  if (c->sol) cusolverDnDestroy(c->sol);
  delete c;
  return CM_OK;
}

static int blu_max_smem_optin() {
  int dev = 0, v = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  return v;
}

int blu_factor(void* ctx, int bs, int nv, int nbn, const int* perm, const int* iperm,
               const int* nrowptr, const int* ncol, const double* vals, void* ws, size_t ws_bytes,
               int* info_host, cudaStream_t st) {
  BluCtx* c = (BluCtx*)ctx;
  const BluLayout L = blu_layout(bs, nv, nbn);
  CM_CHECK_ARG(ws_bytes >= L.total, "workspace too small (cm_blu_workspace_bytes)");
  const int nb = L.nb, nblk = L.nblk;
  char* w = (char*)ws;
  double* P = (double*)(w + L.off_P);
  double* U = (double*)(w + L.off_U);
  int* src = (int*)(w + L.off_src);
  int* ipiv = (int*)(w + L.off_ipiv);
  int* info = (int*)(w + L.off_info);
  double* orig = (double*)(w + L.off_orig);
  double* T[2] = {(double*)(w + L.off_T0), (double*)(w + L.off_T1)};
  const size_t psz = (size_t)2 * nb * nb;  // doubles per panel and per U12 block

  CM_SOLVER(cusolverDnSetStream(c->sol, st));
  CM_BLAS(cublasSetStream(c->blas, st));
  int lwork = 0, lwork2 = 0;
  CM_SOLVER(cusolverDnDgetrf_bufferSize(c->sol, 2 * nb, nb, P, 2 * nb, &lwork));
  CM_SOLVER(cusolverDnDgetrf_bufferSize(c->sol, nb, nb, P, 2 * nb, &lwork2));
  if (lwork2 > lwork) lwork = lwork2;
  if (lwork > c->getrf_lwork) {
    if (c->getrf_work) cudaFree(c->getrf_work);
    c->getrf_work = nullptr;
    c->getrf_lwork = 0;
    CM_CUDA(cudaMalloc(&c->getrf_work, (size_t)lwork * 8));
    c->getrf_lwork = lwork;
  }
  const size_t perm_smem = (size_t)(2 * nb + nb) * 4;
  CM_CHECK_ARG((int)perm_smem <= blu_max_smem_optin(), "block size too large for the pivot kernel");
  if (perm_smem > 48 * 1024)
    CM_CUDA(cudaFuncSetAttribute(blu_perm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)perm_smem));

  CM_CUDA(cudaMemsetAsync(info, 0, (size_t)(nblk + 1) * 4, st));
  int* band_flag = info + nblk;
  const int sgrid = (nbn + 127) / 128;
  // block row 0: A_00 -> top of panel 0, A_01 -> T[0]
  CM_CUDA(cudaMemset2DAsync(P, (size_t)2 * nb * 8, 0, (size_t)nb * 8, nb, st));
  CM_CUDA(cudaMemsetAsync(T[0], 0, (size_t)nb * nb * 8, st));
  blu_scatter_kernel<<<sgrid, 128, 0, st>>>(bs, nv, nbn, 0, iperm, perm, nrowptr, ncol, vals, nullptr,
                                            0, P, 2 * nb, nblk > 1 ? T[0] : nullptr, nb, band_flag);
  CM_LAUNCH_CHECK();
  const double one = 1.0, mone = -1.0;
  for (int k = 0; k < nblk; ++k) {
    double* Pk = P + (size_t)k * psz;
    int* ipk = ipiv + (size_t)k * nb;
    int* srck = src + (size_t)k * 2 * nb;
    if (k == nblk - 1) {  // last block: plain nb x nb LU
      CM_SOLVER(cusolverDnDgetrf(c->sol, nb, nb, Pk, 2 * nb, c->getrf_work, ipk, info + k));
      blu_perm_kernel<<<1, 256, perm_smem, st>>>(nb, nb, ipk, srck);
      CM_LAUNCH_CHECK();
      break;
    }
    double* Uk = U + (size_t)k * psz;
    double* Pn = Pk + psz;
    double* Tc = T[k & 1];
    double* Tn = T[(k + 1) & 1];
    // block row k+1 of the original matrix: A_{k+1,k} -> bottom of the panel, the rest -> orig
    CM_CUDA(cudaMemset2DAsync(Pk + nb, (size_t)2 * nb * 8, 0, (size_t)nb * 8, nb, st));
    CM_CUDA(cudaMemsetAsync(orig, 0, psz * 8, st));
    blu_scatter_kernel<<<sgrid, 128, 0, st>>>(bs, nv, nbn, k + 1, iperm, perm, nrowptr, ncol, vals,
                                              Pk + nb, 2 * nb, orig, nb,
                                              (k + 2 < nblk) ? orig + (size_t)nb * nb : nullptr, nb,
                                              band_flag);
    CM_LAUNCH_CHECK();
    CM_SOLVER(cusolverDnDgetrf(c->sol, 2 * nb, nb, Pk, 2 * nb, c->getrf_work, ipk, info + k));
    blu_perm_kernel<<<1, 256, perm_smem, st>>>(2 * nb, nb, ipk, srck);
    CM_LAUNCH_CHECK();
    {
      dim3 grid((2 * nb + 255) / 256, 2 * nb);
      blu_gather_kernel<<<grid, 256, 0, st>>>(nb, srck, Tc, orig, Uk, Pn, Tn);
      CM_LAUNCH_CHECK();
    }
    CM_BLAS(cublasDtrsm(c->blas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N,
                        CUBLAS_DIAG_UNIT, nb, 2 * nb, &one, Pk, 2 * nb, Uk, nb));
    CM_BLAS(cublasDgemm(c->blas, CUBLAS_OP_N, CUBLAS_OP_N, nb, nb, nb, &mone, Pk + nb, 2 * nb, Uk, nb,
                        &one, Pn, 2 * nb));
    CM_BLAS(cublasDgemm(c->blas, CUBLAS_OP_N, CUBLAS_OP_N, nb, nb, nb, &mone, Pk + nb, 2 * nb,
                        Uk + (size_t)nb * nb, nb, &one, Tn, nb));
  }
  // one synchronisation per Jacobian: singularity / bandwidth report
  std::vector<int> h((size_t)nblk + 1);
  CM_CUDA(cudaMemcpyAsync(h.data(), info, (size_t)(nblk + 1) * 4, cudaMemcpyDeviceToHost, st));
  CM_CUDA(cudaStreamSynchronize(st));
  if (h[nblk] != 0) {
    set_error("cm_blu_factor: the renumbered pattern couples nodes more than nbn=%d positions apart", nbn);
    return CM_E_INVALID;
  }
  int bad = 0;
  for (int k = 0; k < nblk && bad == 0; ++k) {
    if (h[k] < 0) {
      set_error("cm_blu_factor: getrf argument %d rejected in block %d", -h[k], k);
      return CM_E_CUDA;
    }
    if (h[k] > 0) bad = k * nb + h[k];
  }
  if (info_host) *info_host = bad;
  if (bad != 0) {
    set_error("cm_blu_factor: exactly singular matrix (zero pivot at renumbered dof %d)", bad - 1);
    return CM_E_BREAKDOWN;
  }
  return CM_OK;
}

int blu_solve(void* ctx, int bs, int nv, int nbn, const int* perm, const double* b, double* x, void* ws,
              size_t ws_bytes, cudaStream_t st) {
  BluCtx* c = (BluCtx*)ctx;
  const BluLayout L = blu_layout(bs, nv, nbn);
  CM_CHECK_ARG(ws_bytes >= L.total, "workspace too small (cm_blu_workspace_bytes)");
  const int nb = L.nb, nblk = L.nblk;
  char* w = (char*)ws;
  const double* P = (const double*)(w + L.off_P);
  const double* U = (const double*)(w + L.off_U);
  const int* src = (const int*)(w + L.off_src);
  double* y = (double*)(w + L.off_y);
  const size_t psz = (size_t)2 * nb * nb;
  CM_BLAS(cublasSetStream(c->blas, st));
  const size_t vsm = (size_t)2 * nb * 8;
  CM_CHECK_ARG((int)vsm <= blu_max_smem_optin(), "block size too large for the permutation kernel");
  if (vsm > 48 * 1024)
    CM_CUDA(cudaFuncSetAttribute(blu_vecperm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)vsm));
  CM_CUDA(cudaMemsetAsync(y, 0, (size_t)(nblk + 2) * nb * 8, st));
  const int n = bs * nv;
  blu_pack_kernel<<<(n + 255) / 256, 256, 0, st>>>(bs, nv, perm, b, y, 0);
  CM_LAUNCH_CHECK();
  const double one = 1.0, mone = -1.0;
  // forward: P_k (row swaps), L11^{-1}, elimination of block row k+1
  for (int k = 0; k < nblk; ++k) {
    const double* Pk = P + (size_t)k * psz;
    double* yk = y + (size_t)k * nb;
    const int m = (k == nblk - 1) ? nb : 2 * nb;
    blu_vecperm_kernel<<<1, 1024, vsm, st>>>(m, src + (size_t)k * 2 * nb, yk);
    CM_LAUNCH_CHECK();
    CM_BLAS(cublasDtrsv(c->blas, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_UNIT, nb, Pk, 2 * nb,
                        yk, 1));
    if (k < nblk - 1)
      CM_BLAS(cublasDgemv(c->blas, CUBLAS_OP_N, nb, nb, &mone, Pk + nb, 2 * nb, yk, 1, &one, yk + nb, 1));
  }
  // backward: y_k = U11^{-1} (y_k - [U_{k,k+1} U_{k,k+2}] [y_{k+1}; y_{k+2}])
  for (int k = nblk - 1; k >= 0; --k) {
    const double* Pk = P + (size_t)k * psz;
    double* yk = y + (size_t)k * nb;
    if (k < nblk - 1)
      CM_BLAS(cublasDgemv(c->blas, CUBLAS_OP_N, nb, 2 * nb, &mone, U + (size_t)k * psz, nb, yk + nb, 1,
                          &one, yk, 1));
    CM_BLAS(cublasDtrsv(c->blas, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, nb, Pk,
                        2 * nb, yk, 1));
  }
  blu_pack_kernel<<<(n + 255) / 256, 256, 0, st>>>(bs, nv, perm, y, x, 1);
  CM_LAUNCH_CHECK();
  return CM_OK;
}

}  // namespace cm
