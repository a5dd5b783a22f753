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
// For four or more blocks the elimination runs from both ends at once ("twisted" factorisation): chain A
// eliminates the blocks 0..mA-1 top-down, chain B -- the same algorithm on the node-reversed matrix, on a
// second stream with its own library handles -- the blocks nblk-1..mA+2 bottom-up; the two leftover block
// rows meet in a dense 2nb x 2nb system for the blocks mA, mA+1.  Each chain only draws pivots from its own
// rows, and the columns it eliminates vanish outside those rows, so this is as robust as the one-sided
// sweep (a nonsingular matrix never produces a zero pivot) at half the sequential depth.
// The dense panel LU, the triangular solves and the Schur update are library calls (cuSOLVER getrf,
// cuBLAS trsm/gemm: plain dense fp64 kernels); the band scatter from the DOLFIN-layout CSR, the
// pivot-to-permutation conversion, the fused row gather (laswp + relocation) and the solve-phase
// permutations are the kernels below.
#include <cublas_v2.h>
#include <cusolverDn.h>

#include <string>
#include <thread>
#include <vector>

#include "cm_common.cuh"
#include "cm_decl.cuh"

namespace cm {

int blu_destroy(void* ctx);

struct BluCtx {
  struct Side {  // library handles and getrf scratch of one chain (the chains run concurrently)
    cusolverDnHandle_t sol = nullptr;
    cublasHandle_t blas = nullptr;
    double* getrf_work = nullptr;
    int getrf_lwork = 0;
  } side[2];
  cudaStream_t stB = nullptr;  // stream of chain B
  cudaEvent_t fork = nullptr, join = nullptr;
};

struct BluLayout {
  int bs, nv, nbn, nb, nblk;
  bool twisted;  // two-sided elimination (nblk >= 4)
  int mA, mB;    // blocks eliminated by chain A (top-down) and chain B (bottom-up); mA + mB = nblk - 2
  size_t off_P, off_U, off_src, off_ipiv, off_info, off_scratch, off_mid, off_midpiv, off_z, off_y, off_ord,
      total;
};

static inline size_t up256(size_t x) { return (x + 255) / 256 * 256; }

static BluLayout blu_layout(int bs, int nv, int nbn) {
  BluLayout L;
  L.bs = bs;
  L.nv = nv;
  L.nbn = nbn;
  L.nb = bs * nbn;
  L.nblk = (nv + nbn - 1) / nbn;
  L.twisted = L.nblk >= 4;
  L.mA = L.twisted ? (L.nblk - 2) / 2 : 0;
  L.mB = L.twisted ? L.nblk - 2 - L.mA : 0;
  const size_t nb = (size_t)L.nb, nblk = (size_t)L.nblk;
  const size_t psz = 2 * nb * nb;
  size_t o = 0;
  L.off_P = o;       o += up256(nblk * psz * 8);            // panels [L11 / U11; L21] of both chains, ld 2nb
  L.off_U = o;       o += up256(nblk * psz * 8);            // [U_{k,k+1} U_{k,k+2}], ld nb
  L.off_src = o;     o += up256(nblk * 2 * nb * 4);         // row permutation of every step
  L.off_ipiv = o;    o += up256(nblk * nb * 4);
  L.off_info = o;    o += up256((nblk + 2) * 4);            // getrf info per step, middle system, band flag
  L.off_scratch = o; o += up256(4 * psz * 8);               // per chain: orig (psz), T0, T1 (psz/2 each)
  L.off_mid = o;     o += up256(2 * psz * 8);               // LU of the 2nb x 2nb middle system
  L.off_midpiv = o;  o += up256(4 * nb * 4);                // its pivots and row permutation
  L.off_z = o;       o += up256(2 * nb * 8);
  L.off_y = o;       o += up256(2 * (nblk + 2) * nb * 8);   // permuted right-hand side / solution per chain
  L.off_ord = o;     o += up256((2 * nblk * (size_t)nbn + (size_t)nv) * 4);  // ipermA, ipermB (padded), permB
  L.total = o;
  return L;
}

size_t blu_workspace_bytes(int bs, int nv, int nbn) { return blu_layout(bs, nv, nbn).total; }

// ---- kernels ---------------------------------------------------------------------------------

// Dense images of block row kb of the renumbered matrix: column block kb-1 -> dL, kb -> dD,
// kb+1 -> dU (column major, leading dimensions ldL/ldD/ldU; a NULL destination skips that block).
// Padding positions (iperm < 0) get a unit diagonal.  bs=2 reads the DOLFIN/PETSc value layout
// (scalar row 2v+c starts at 4*rowptr[v] + c*2*deg), bs=1 the node CSR.
__global__ void __launch_bounds__(128)
blu_scatter_kernel(const int bs, const int nbn, const int kb,
                   const int* __restrict__ iperm, const int* __restrict__ perm,
                   const int* __restrict__ nrowptr, const int* __restrict__ ncol,
                   const double* __restrict__ vals, double* __restrict__ dL, const int ldL,
                   double* __restrict__ dD, const int ldD, double* __restrict__ dU, const int ldU,
                   int* __restrict__ band_flag) {
  const int lp = blockIdx.x * blockDim.x + threadIdx.x;
  if (lp >= nbn) return;
  const int p = kb * nbn + lp;
  const int v = iperm[p];  // padded inverse permutation: -1 marks a padding position
  if (v < 0) {
    if (dD != nullptr)
      for (int c = 0; c < bs; ++c) dD[(size_t)(bs * lp + c) + (size_t)(bs * lp + c) * ldD] = 1.0;
    return;
  }
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

// padded inverse permutations of both chains and the permutation of chain B (node order reversed):
//   ipA[p] = iperm[p] (p < nv) or -1,  ipB[p] = ipA[npad-1-p],  permB[v] = npad-1-perm[v]
__global__ void __launch_bounds__(256)
blu_orderings_kernel(const int nv, const int npad, const int* __restrict__ perm,
                     const int* __restrict__ iperm, int* __restrict__ ipA, int* __restrict__ ipB,
                     int* __restrict__ permB) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < npad) {
    const int a = i < nv ? iperm[i] : -1;
    ipA[i] = a;
    ipB[npad - 1 - i] = a;
  }
  if (i < nv) permB[i] = npad - 1 - perm[i];
}

// node reversal inside a block, component order kept
__device__ __forceinline__ int blu_rev(const int i, const int bs, const int nbn) {
  return bs * (nbn - 1 - i / bs) + i % bs;
}

// middle system for the unknowns [x_mA (chain-A order); x'_mB (chain-B order)]:
//   M = [[S_A, T_A rev], [T_B rev, S_B]]   (column major, ld 2nb; S_* have ld 2nb, T_* ld nb)
__global__ void __launch_bounds__(256)
blu_middle_kernel(const int nb, const int bs, const int nbn, const double* __restrict__ SA,
                  const double* __restrict__ TA, const double* __restrict__ SB,
                  const double* __restrict__ TB, double* __restrict__ M) {
  const int c = blockIdx.y;
  for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < 2 * nb; r += gridDim.x * blockDim.x) {
    double v;
    if (r < nb)
      v = (c < nb) ? SA[(size_t)r + (size_t)c * 2 * nb] : TA[(size_t)r + (size_t)blu_rev(c - nb, bs, nbn) * nb];
    else
      v = (c < nb) ? TB[(size_t)(r - nb) + (size_t)blu_rev(c, bs, nbn) * nb]
                   : SB[(size_t)(r - nb) + (size_t)(c - nb) * 2 * nb];
    M[(size_t)r + (size_t)c * 2 * nb] = v;
  }
}

// z = [yA_mid; yB_mid]  (dir = 0)   or   yA_mid = z_A, yA_next = rev(z_B), yB_mid = z_B, yB_next = rev(z_A)  (dir = 1)
__global__ void __launch_bounds__(256)
blu_middle_vec_kernel(const int nb, const int bs, const int nbn, double* __restrict__ yA_mid,
                      double* __restrict__ yB_mid, double* __restrict__ z, const int dir) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nb) return;
  if (dir == 0) {
    z[i] = yA_mid[i];
    z[nb + i] = yB_mid[i];
  } else {
    const int j = blu_rev(i, bs, nbn);
    yA_mid[i] = z[i];
    yA_mid[nb + j] = z[nb + i];
    yB_mid[i] = z[nb + i];
    yB_mid[nb + j] = z[i];
  }
}

// x[bs*v+c] = yA[bs*perm[v]+c] when the node lies in the blocks 0..last_A of chain A, else yB[bs*permB[v]+c]
__global__ void __launch_bounds__(256)
blu_unpack2_kernel(const int bs, const int nv, const int nbn, const int last_A,
                   const int* __restrict__ perm, const int* __restrict__ permB,
                   const double* __restrict__ yA, const double* __restrict__ yB, double* __restrict__ x) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= bs * nv) return;
  const int v = i / bs, c = i - v * bs;
  const int pa = perm[v];
  x[i] = (pa / nbn <= last_A) ? yA[bs * pa + c] : yB[bs * permB[v] + c];
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
  bool ok = true;
  for (int i = 0; i < 2 && ok; ++i) {
    ok = cusolverDnCreate(&c->side[i].sol) == CUSOLVER_STATUS_SUCCESS &&
         cublasCreate(&c->side[i].blas) == CUBLAS_STATUS_SUCCESS;
    if (ok) cublasSetPointerMode(c->side[i].blas, CUBLAS_POINTER_MODE_HOST);
  }
  ok = ok && cudaStreamCreateWithFlags(&c->stB, cudaStreamNonBlocking) == cudaSuccess &&
       cudaEventCreateWithFlags(&c->fork, cudaEventDisableTiming) == cudaSuccess &&
       cudaEventCreateWithFlags(&c->join, cudaEventDisableTiming) == cudaSuccess;
  if (!ok) {
    set_error("cm_blu_create: cuSOLVER / cuBLAS / stream creation failed");
    blu_destroy(c);
    return nullptr;
  }
  return c;
}

int blu_destroy(void* ctx) {
  BluCtx* c = (BluCtx*)ctx;
  if (c == nullptr) return CM_OK;
  for (int i = 0; i < 2; ++i) {
    if (c->side[i].getrf_work) cudaFree(c->side[i].getrf_work);
    if (c->side[i].blas) cublasDestroy(c->side[i].blas);
    if (c->side[i].sol) cusolverDnDestroy(c->side[i].sol);
  }
  if (c->fork) cudaEventDestroy(c->fork);
  if (c->join) cudaEventDestroy(c->join);
  if (c->stB) cudaStreamDestroy(c->stB);
  delete c;
  return CM_OK;
}

static int blu_max_smem_optin() {
  int dev = 0, v = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  return v;
}

// views of one chain inside the workspace
struct BluChain {
  double *P, *U, *orig, *T[2], *y;
  int *src, *ipiv, *info;
  const int *perm, *iperm;  // iperm padded to nblk*nbn entries (-1 = padding)
  int nsteps;               // full elimination steps
  bool final_block;         // one-sided sweep: plain LU of the last block after the steps
};

static void blu_chains(const BluLayout& L, char* w, const int* perm, BluChain& A, BluChain& B) {
  const size_t psz = (size_t)2 * L.nb * L.nb;
  double* P = (double*)(w + L.off_P);
  double* U = (double*)(w + L.off_U);
  int* src = (int*)(w + L.off_src);
  int* ipiv = (int*)(w + L.off_ipiv);
  int* info = (int*)(w + L.off_info);
  double* scr = (double*)(w + L.off_scratch);
  double* y = (double*)(w + L.off_y);
  int* ord = (int*)(w + L.off_ord);
  const size_t npad = (size_t)L.nblk * L.nbn, ylen = (size_t)(L.nblk + 2) * L.nb;
  A.P = P; A.U = U; A.src = src; A.ipiv = ipiv; A.info = info;
  A.orig = scr; A.T[0] = scr + psz; A.T[1] = scr + psz + psz / 2; A.y = y;
  A.perm = perm; A.iperm = ord;
  A.nsteps = L.twisted ? L.mA : L.nblk - 1;
  A.final_block = !L.twisted;
  const size_t pa = (size_t)A.nsteps + 1;  // panels / steps used by chain A (the last holds its leftover S)
  B.P = P + pa * psz; B.U = U + pa * psz; B.src = src + pa * 2 * L.nb; B.ipiv = ipiv + pa * L.nb;
  B.info = info + pa;
  B.orig = scr + 2 * psz; B.T[0] = scr + 3 * psz; B.T[1] = scr + 3 * psz + psz / 2; B.y = y + ylen;
  B.perm = ord + 2 * npad; B.iperm = ord + npad;
  B.nsteps = L.twisted ? L.mB : 0;
  B.final_block = false;
}

static int blu_prepare_side(BluCtx::Side& h, int nb, const double* P, cudaStream_t st) {
  CM_SOLVER(cusolverDnSetStream(h.sol, st));
  CM_BLAS(cublasSetStream(h.blas, st));
  int lw[3] = {0, 0, 0};
  CM_SOLVER(cusolverDnDgetrf_bufferSize(h.sol, 2 * nb, nb, const_cast<double*>(P), 2 * nb, &lw[0]));
  CM_SOLVER(cusolverDnDgetrf_bufferSize(h.sol, nb, nb, const_cast<double*>(P), 2 * nb, &lw[1]));
  CM_SOLVER(cusolverDnDgetrf_bufferSize(h.sol, 2 * nb, 2 * nb, const_cast<double*>(P), 2 * nb, &lw[2]));
  int lwork = lw[0] > lw[1] ? lw[0] : lw[1];
  if (lw[2] > lwork) lwork = lw[2];
  if (lwork > h.getrf_lwork) {
    if (h.getrf_work) cudaFree(h.getrf_work);
    h.getrf_work = nullptr;
    h.getrf_lwork = 0;
    CM_CUDA(cudaMalloc(&h.getrf_work, (size_t)lwork * 8));
    h.getrf_lwork = lwork;
  }
  return CM_OK;
}

const char* last_error_text();

// Runs `body` (the enqueue loop of chain B) on a worker thread bound to the caller's device, so that the
// two chains are issued concurrently: the panel LU is a long sequence of small launches and one host thread
// cannot feed two streams at full rate.  Errors of the worker are re-raised on the calling thread.
template <class F>
struct BluWorker {
  std::thread th;
  int rc = CM_OK;
  std::string msg;
  explicit BluWorker(F body) {
    int dev = 0;
    cudaGetDevice(&dev);
    th = std::thread([this, dev, body]() {
      cudaSetDevice(dev);
      rc = body();
      if (rc != CM_OK) msg = last_error_text();
    });
  }
  int join() {
    if (th.joinable()) th.join();
    if (rc != CM_OK) set_error("%s", msg.c_str());
    return rc;
  }
  ~BluWorker() {
    if (th.joinable()) th.join();
  }
};

// steps 0..nsteps-1 of one chain; afterwards the leftover block row [S | T] sits in the top half of panel
// `nsteps` and in T[nsteps & 1] (or, for the one-sided sweep, the last block is factorised in place)
static int blu_chain_factor(BluCtx::Side& h, const BluChain& C, int bs, int nbn, int nblk,
                            const int* nrowptr, const int* ncol, const double* vals, int* band_flag,
                            size_t perm_smem, cudaStream_t st) {
  const int nb = bs * nbn;
  const size_t psz = (size_t)2 * nb * nb;
  const int sgrid = (nbn + 127) / 128;
  const double one = 1.0, mone = -1.0;
  // block row 0: A_00 -> top of panel 0, A_01 -> T[0]
  CM_CUDA(cudaMemset2DAsync(C.P, (size_t)2 * nb * 8, 0, (size_t)nb * 8, nb, st));
  CM_CUDA(cudaMemsetAsync(C.T[0], 0, (size_t)nb * nb * 8, st));
  blu_scatter_kernel<<<sgrid, 128, 0, st>>>(bs, nbn, 0, C.iperm, C.perm, nrowptr, ncol, vals, nullptr, 0,
                                            C.P, 2 * nb, nblk > 1 ? C.T[0] : nullptr, nb, band_flag);
  CM_LAUNCH_CHECK();
  for (int k = 0; k < C.nsteps; ++k) {
    double* Pk = C.P + (size_t)k * psz;
    double* Pn = Pk + psz;
    double* Uk = C.U + (size_t)k * psz;
    int* ipk = C.ipiv + (size_t)k * nb;
    int* srck = C.src + (size_t)k * 2 * nb;
    double* Tc = C.T[k & 1];
    double* Tn = C.T[(k + 1) & 1];
    // block row k+1 of the original matrix: A_{k+1,k} -> bottom of the panel, the rest -> orig
    CM_CUDA(cudaMemset2DAsync(Pk + nb, (size_t)2 * nb * 8, 0, (size_t)nb * 8, nb, st));
    CM_CUDA(cudaMemsetAsync(C.orig, 0, psz * 8, st));
    blu_scatter_kernel<<<sgrid, 128, 0, st>>>(bs, nbn, k + 1, C.iperm, C.perm, nrowptr, ncol, vals, Pk + nb,
                                              2 * nb, C.orig, nb,
                                              (k + 2 < nblk) ? C.orig + (size_t)nb * nb : nullptr, nb,
                                              band_flag);
    CM_LAUNCH_CHECK();
    CM_SOLVER(cusolverDnDgetrf(h.sol, 2 * nb, nb, Pk, 2 * nb, h.getrf_work, ipk, C.info + k));
    blu_perm_kernel<<<1, 256, perm_smem, st>>>(2 * nb, nb, ipk, srck);
    CM_LAUNCH_CHECK();
    {
      dim3 grid((2 * nb + 255) / 256, 2 * nb);
      blu_gather_kernel<<<grid, 256, 0, st>>>(nb, srck, Tc, C.orig, Uk, Pn, Tn);
      CM_LAUNCH_CHECK();
    }
    CM_BLAS(cublasDtrsm(h.blas, CUBLAS_SIDE_LEFT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_UNIT, nb,
                        2 * nb, &one, Pk, 2 * nb, Uk, nb));
    CM_BLAS(cublasDgemm(h.blas, CUBLAS_OP_N, CUBLAS_OP_N, nb, nb, nb, &mone, Pk + nb, 2 * nb, Uk, nb, &one,
                        Pn, 2 * nb));
    CM_BLAS(cublasDgemm(h.blas, CUBLAS_OP_N, CUBLAS_OP_N, nb, nb, nb, &mone, Pk + nb, 2 * nb,
                        Uk + (size_t)nb * nb, nb, &one, Tn, nb));
  }
  if (C.final_block) {  // one-sided sweep: plain nb x nb LU of the last block
    const int k = C.nsteps;
    double* Pk = C.P + (size_t)k * psz;
    CM_SOLVER(cusolverDnDgetrf(h.sol, nb, nb, Pk, 2 * nb, h.getrf_work, C.ipiv + (size_t)k * nb, C.info + k));
    blu_perm_kernel<<<1, 256, perm_smem, st>>>(nb, nb, C.ipiv + (size_t)k * nb, C.src + (size_t)k * 2 * nb);
    CM_LAUNCH_CHECK();
  }
  return CM_OK;
}

int blu_factor(void* ctx, int bs, int nv, int nbn, const int* perm, const int* iperm,
               const int* nrowptr, const int* ncol, const double* vals, void* ws, size_t ws_bytes,
               int* info_host, cudaStream_t st) {
  BluCtx* c = (BluCtx*)ctx;
  const BluLayout L = blu_layout(bs, nv, nbn);
  CM_CHECK_ARG(ws_bytes >= L.total, "workspace too small (cm_blu_workspace_bytes)");
  const int nb = L.nb, nblk = L.nblk;
  char* w = (char*)ws;
  BluChain A, B;
  blu_chains(L, w, perm, A, B);
  int* info = (int*)(w + L.off_info);
  int* band_flag = info + nblk + 1;
  const int npad = nblk * nbn;

  int rc;
  if ((rc = blu_prepare_side(c->side[0], nb, A.P, st))) return rc;
  if (L.twisted && (rc = blu_prepare_side(c->side[1], nb, B.P, c->stB))) return rc;
  const size_t perm_smem = (size_t)(4 * nb) * 4;  // up to 2nb rows and 2nb pivots (middle system)
  CM_CHECK_ARG((int)perm_smem <= blu_max_smem_optin(), "block size too large for the pivot kernel");
  if (perm_smem > 48 * 1024)
    CM_CUDA(cudaFuncSetAttribute(blu_perm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)perm_smem));
  CM_CUDA(cudaMemsetAsync(info, 0, (size_t)(nblk + 2) * 4, st));
  {
    int* ord = (int*)(w + L.off_ord);
    blu_orderings_kernel<<<(npad + 255) / 256, 256, 0, st>>>(nv, npad, perm, iperm, ord, ord + npad,
                                                            ord + 2 * (size_t)npad);
    CM_LAUNCH_CHECK();
  }
  if (L.twisted) {
    CM_CUDA(cudaEventRecord(c->fork, st));
    CM_CUDA(cudaStreamWaitEvent(c->stB, c->fork, 0));
    auto bodyB = [&]() -> int {
      int e = blu_chain_factor(c->side[1], B, bs, nbn, nblk, nrowptr, ncol, vals, band_flag, perm_smem, c->stB);
      if (e) return e;
      CM_CUDA(cudaEventRecord(c->join, c->stB));
      return CM_OK;
    };
    BluWorker<decltype(bodyB)> worker(bodyB);
    rc = blu_chain_factor(c->side[0], A, bs, nbn, nblk, nrowptr, ncol, vals, band_flag, perm_smem, st);
    const int rcB = worker.join();
    if (rc) return rc;
    if (rcB) return rcB;
  } else if ((rc = blu_chain_factor(c->side[0], A, bs, nbn, nblk, nrowptr, ncol, vals, band_flag, perm_smem,
                                    st))) {
    return rc;
  }
  if (L.twisted) {
    CM_CUDA(cudaStreamWaitEvent(st, c->join, 0));
    const size_t psz = (size_t)2 * nb * nb;
    double* M = (double*)(w + L.off_mid);
    int* ipm = (int*)(w + L.off_midpiv);
    dim3 grid((2 * nb + 255) / 256, 2 * nb);
    blu_middle_kernel<<<grid, 256, 0, st>>>(nb, bs, nbn, A.P + (size_t)A.nsteps * psz, A.T[A.nsteps & 1],
                                            B.P + (size_t)B.nsteps * psz, B.T[B.nsteps & 1], M);
    CM_LAUNCH_CHECK();
    CM_SOLVER(cusolverDnDgetrf(c->side[0].sol, 2 * nb, 2 * nb, M, 2 * nb, c->side[0].getrf_work, ipm,
                               info + nblk));
    blu_perm_kernel<<<1, 256, perm_smem, st>>>(2 * nb, 2 * nb, ipm, ipm + 2 * nb);
    CM_LAUNCH_CHECK();
  }
  // one synchronisation per Jacobian: singularity / bandwidth report
  std::vector<int> h((size_t)nblk + 2);
  CM_CUDA(cudaMemcpyAsync(h.data(), info, (size_t)(nblk + 2) * 4, cudaMemcpyDeviceToHost, st));
  CM_CUDA(cudaStreamSynchronize(st));
  if (h[nblk + 1] != 0) {
    set_error("cm_blu_factor: the renumbered pattern couples nodes more than nbn=%d positions apart", nbn);
    return CM_E_INVALID;
  }
  int bad = 0;
  for (int k = 0; k <= nblk && bad == 0; ++k) {
    if (h[k] < 0) {
      set_error("cm_blu_factor: getrf argument %d rejected in step %d", -h[k], k);
      return CM_E_CUDA;
    }
    if (h[k] > 0) bad = k * nb + h[k];
  }
  if (info_host) *info_host = bad;
  if (bad != 0) {
    set_error("cm_blu_factor: exactly singular matrix (zero pivot in elimination step %d, column %d)",
              (bad - 1) / nb, (bad - 1) % nb);
    return CM_E_BREAKDOWN;
  }
  return CM_OK;
}

// forward substitution of one chain: P_k (row swaps), L11^{-1}, elimination of block row k+1
static int blu_chain_forward(BluCtx::Side& h, const BluChain& C, int nb, size_t vsm, cudaStream_t st) {
  const size_t psz = (size_t)2 * nb * nb;
  const double one = 1.0, mone = -1.0;
  const int nfac = C.nsteps + (C.final_block ? 1 : 0);
  for (int k = 0; k < nfac; ++k) {
    const double* Pk = C.P + (size_t)k * psz;
    double* yk = C.y + (size_t)k * nb;
    const bool last = k == C.nsteps;  // the plain LU of the last block
    blu_vecperm_kernel<<<1, 1024, vsm, st>>>(last ? nb : 2 * nb, C.src + (size_t)k * 2 * nb, yk);
    CM_LAUNCH_CHECK();
    CM_BLAS(cublasDtrsv(h.blas, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_UNIT, nb, Pk, 2 * nb, yk, 1));
    if (!last)
      CM_BLAS(cublasDgemv(h.blas, CUBLAS_OP_N, nb, nb, &mone, Pk + nb, 2 * nb, yk, 1, &one, yk + nb, 1));
  }
  return CM_OK;
}

// backward substitution: y_k = U11^{-1} (y_k - [U_{k,k+1} U_{k,k+2}] [y_{k+1}; y_{k+2}])
static int blu_chain_backward(BluCtx::Side& h, const BluChain& C, int nb, cudaStream_t st) {
  const size_t psz = (size_t)2 * nb * nb;
  const double one = 1.0, mone = -1.0;
  const int nfac = C.nsteps + (C.final_block ? 1 : 0);
  for (int k = nfac - 1; k >= 0; --k) {
    const double* Pk = C.P + (size_t)k * psz;
    double* yk = C.y + (size_t)k * nb;
    if (k < C.nsteps)
      CM_BLAS(cublasDgemv(h.blas, CUBLAS_OP_N, nb, 2 * nb, &mone, C.U + (size_t)k * psz, nb, yk + nb, 1, &one,
                          yk, 1));
    CM_BLAS(cublasDtrsv(h.blas, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, nb, Pk, 2 * nb, yk,
                        1));
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
  BluChain A, B;
  blu_chains(L, w, perm, A, B);
  CM_BLAS(cublasSetStream(c->side[0].blas, st));
  if (L.twisted) CM_BLAS(cublasSetStream(c->side[1].blas, c->stB));
  const size_t vsm = (size_t)2 * nb * 8;
  CM_CHECK_ARG((int)vsm <= blu_max_smem_optin(), "block size too large for the permutation kernel");
  if (vsm > 48 * 1024)
    CM_CUDA(cudaFuncSetAttribute(blu_vecperm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)vsm));
  const size_t ylen = (size_t)(nblk + 2) * nb;
  const int n = bs * nv;
  int rc;
  CM_CUDA(cudaMemsetAsync(A.y, 0, (L.twisted ? 2 : 1) * ylen * 8, st));
  blu_pack_kernel<<<(n + 255) / 256, 256, 0, st>>>(bs, nv, A.perm, b, A.y, 0);
  CM_LAUNCH_CHECK();
  if (!L.twisted) {
    if ((rc = blu_chain_forward(c->side[0], A, nb, vsm, st))) return rc;
    if ((rc = blu_chain_backward(c->side[0], A, nb, st))) return rc;
    blu_pack_kernel<<<(n + 255) / 256, 256, 0, st>>>(bs, nv, A.perm, A.y, x, 1);
    CM_LAUNCH_CHECK();
    return CM_OK;
  }
  blu_pack_kernel<<<(n + 255) / 256, 256, 0, st>>>(bs, nv, B.perm, b, B.y, 0);
  CM_LAUNCH_CHECK();
  CM_CUDA(cudaEventRecord(c->fork, st));
  CM_CUDA(cudaStreamWaitEvent(c->stB, c->fork, 0));
  {
    auto bodyB = [&]() -> int {
      int e = blu_chain_forward(c->side[1], B, nb, vsm, c->stB);
      if (e) return e;
      CM_CUDA(cudaEventRecord(c->join, c->stB));
      return CM_OK;
    };
    BluWorker<decltype(bodyB)> worker(bodyB);
    rc = blu_chain_forward(c->side[0], A, nb, vsm, st);
    const int rcB = worker.join();
    if (rc) return rc;
    if (rcB) return rcB;
  }
  CM_CUDA(cudaStreamWaitEvent(st, c->join, 0));
  {  // middle system
    const double* M = (const double*)(w + L.off_mid);
    const int* srcm = (const int*)(w + L.off_midpiv) + 2 * nb;
    double* z = (double*)(w + L.off_z);
    double* yAm = A.y + (size_t)A.nsteps * nb;
    double* yBm = B.y + (size_t)B.nsteps * nb;
    blu_middle_vec_kernel<<<(nb + 255) / 256, 256, 0, st>>>(nb, bs, nbn, yAm, yBm, z, 0);
    CM_LAUNCH_CHECK();
    const size_t vsm2 = (size_t)2 * nb * 8;
    blu_vecperm_kernel<<<1, 1024, vsm2, st>>>(2 * nb, srcm, z);
    CM_LAUNCH_CHECK();
    CM_BLAS(cublasDtrsv(c->side[0].blas, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, CUBLAS_DIAG_UNIT, 2 * nb, M,
                        2 * nb, z, 1));
    CM_BLAS(cublasDtrsv(c->side[0].blas, CUBLAS_FILL_MODE_UPPER, CUBLAS_OP_N, CUBLAS_DIAG_NON_UNIT, 2 * nb, M,
                        2 * nb, z, 1));
    blu_middle_vec_kernel<<<(nb + 255) / 256, 256, 0, st>>>(nb, bs, nbn, yAm, yBm, z, 1);
    CM_LAUNCH_CHECK();
  }
  CM_CUDA(cudaEventRecord(c->fork, st));
  CM_CUDA(cudaStreamWaitEvent(c->stB, c->fork, 0));
  {
    auto bodyB = [&]() -> int {
      int e = blu_chain_backward(c->side[1], B, nb, c->stB);
      if (e) return e;
      CM_CUDA(cudaEventRecord(c->join, c->stB));
      return CM_OK;
    };
    BluWorker<decltype(bodyB)> worker(bodyB);
    rc = blu_chain_backward(c->side[0], A, nb, st);
    const int rcB = worker.join();
    if (rc) return rc;
    if (rcB) return rcB;
  }
  CM_CUDA(cudaStreamWaitEvent(st, c->join, 0));
  blu_unpack2_kernel<<<(n + 255) / 256, 256, 0, st>>>(bs, nv, nbn, A.nsteps, A.perm, B.perm, A.y, B.y, x);
  CM_LAUNCH_CHECK();
  return CM_OK;
}

}  // namespace cm
