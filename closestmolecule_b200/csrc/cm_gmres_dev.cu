// Device-resident restarted GMRES (right preconditioning, classical Gram-Schmidt) for the structured P-Q
// solver: the Hessenberg column, the Givens rotations, the residual estimate and the iteration counter live
// on the GPU, so that ONE GMRES iteration is a fixed sequence of kernels without changing arguments (it is
// captured once as a CUDA graph and replayed, cm_pqsolver.cu):
//   dot kernel      h_i = <v_i, w> (i <= j) in one pass over w; per-CTA partials, the last CTA to finish
//                   adds them up in a fixed order (deterministic, no second launch)
//   [all-reduce]    multi-GPU: one push-and-sum launch for the j+1 numbers (cm_dist.cu)
//   update kernel   w' = w - sum h_i v_i stored as the NEXT basis vector without normalising it, and
//                   ||w'||^2 accumulated in the same pass (exact norm of the orthogonalised vector: with a
//                   multigrid preconditioner A M^-1 = I + small, so ||w'|| << ||w|| and the Pythagorean
//                   shortcut <w,w> - sum h_i^2 loses every digit)
//   [all-reduce]    multi-GPU: one number
//   givens kernel   one thread: h_{j+1,j} = ||w'||, the scale 1/h_{j+1,j} of the new basis vector (applied
//                   lazily by every kernel that reads the basis: no extra pass to normalise it), Givens
//                   rotations, residual estimate, convergence flag, j <- j+1
// The host reads a 64-byte status block per iteration (convergence test) -- nothing else crosses the bus.
// Replaces PETScLUSolver::solve inside NewtonSolver [ext] (advection_diffusion_convergence.py:212).
#include <math.h>

#include "cm_common.cuh"
#include "cm_decl.cuh"

namespace cm {

constexpr int kGB = 256;
constexpr int kGGrid = 148 * 4;
constexpr int kGK = 8;   // dot products per pass over w

size_t gmres_dev_doubles(long long n, int m) {
  const long long ne = (n + 1) & ~1LL;
  const size_t ngroups = (size_t)(m + 2 + kGK - 1) / kGK;
  return (size_t)ne * (m + 1) + (size_t)ne * 2        // basis, w, z
         + (size_t)ne * m                             // preconditioned basis Z (flexible variant: x = Z y)
         + 3 * (size_t)(m + 2) + 4                    // dots, coef, vscale, nrm
         + (size_t)(m + 1) * m + 2 * (size_t)m + (size_t)(m + 1) + (size_t)m   // H, cs, sn, g, y
         + kGmresStatus + 8                           // status, tickets
         + ngroups * kGGrid * kGK + 64;               // reduction scratch
}

void gmres_dev_layout(GmresDev& G, long long n, int m, double* work) {
  const long long ne = (n + 1) & ~1LL;
  G.n = n;
  G.ne = ne;
  G.m = m;
  double* p = work;
  G.V = p; p += (size_t)ne * (m + 1);
  G.w = p; p += ne;
  G.z = p; p += ne;
  G.Z = p; p += (size_t)ne * m;
  G.dots = p; p += m + 2;
  G.coef = p; p += m + 2;
  G.vscale = p; p += m + 2;
  G.nrm = p; p += 4;
  G.H = p; p += (size_t)(m + 1) * m;
  G.cs = p; p += m;
  G.sn = p; p += m;
  G.g = p; p += m + 1;
  G.y = p; p += m;
  G.st = p; p += kGmresStatus;
  G.tickets = reinterpret_cast<unsigned*>(p); p += 8;
  p = reinterpret_cast<double*>((reinterpret_cast<uintptr_t>(p) + 15) & ~(uintptr_t)15);
  G.partial = p;
}

__device__ __forceinline__ double gd_block_sum(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = (l < (blockDim.x >> 5)) ? sh[l] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
  }
  __syncthreads();
  return r;  // valid in thread 0
}

// the vector is [0, n) (sl.len == 0) or the union of two index ranges (strip-partitioned rank)
__device__ __forceinline__ long long gd_index(long long i, const VecSlice s) {
  return i < s.len ? s.off0 + i : s.off1 + (i - s.len);
}

// one pass over w for K consecutive basis vectors (fully unrolled: K independent loads in flight per thread)
template <int K, bool SLICE>
__device__ __forceinline__ void gd_dot_group(const GmresDev& G, const VecSlice& sl, const double* __restrict__ Vb,
                                             const double* __restrict__ w, double* acc) {
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = 0.0;
  if (SLICE) {
    // the two index ranges of a strip, each as an optional leading element, 16-byte aligned pairs, optional tail
    // (the q-part starts at an odd offset when the vertex count is odd)
    for (int part = 0; part < 2; ++part) {
      const long long s0 = part ? sl.off1 : sl.off0;
      const long long lead = s0 & 1;
      const long long npair = (sl.len - lead) >> 1;
      for (long long i = blockIdx.x * (long long)kGB + threadIdx.x; i < npair; i += (long long)gridDim.x * kGB) {
        const long long q = s0 + lead + 2 * i;
        const double2 ww = *reinterpret_cast<const double2*>(w + q);
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const double2 vv = *reinterpret_cast<const double2*>(Vb + (size_t)k * G.ne + q);
          acc[k] += vv.x * ww.x + vv.y * ww.y;
        }
      }
      if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (lead) {
#pragma unroll
          for (int k = 0; k < K; ++k) acc[k] += Vb[(size_t)k * G.ne + s0] * w[s0];
        }
        if ((sl.len - lead) & 1) {
          const long long q = s0 + sl.len - 1;
#pragma unroll
          for (int k = 0; k < K; ++k) acc[k] += Vb[(size_t)k * G.ne + q] * w[q];
        }
      }
    }
  } else {
    const long long n2 = G.n >> 1;
    const double2* w2 = reinterpret_cast<const double2*>(w);
    for (long long i = blockIdx.x * (long long)kGB + threadIdx.x; i < n2; i += (long long)gridDim.x * kGB) {
      const double2 ww = w2[i];
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const double2 vv = reinterpret_cast<const double2*>(Vb + (size_t)k * G.ne)[i];
        acc[k] += vv.x * ww.x + vv.y * ww.y;
      }
    }
  }
}

// dots[k] = vscale[k] <V_k, w> for k <= j (V_k holds the unnormalised basis vector);  j = st[0]
template <bool SLICE>
__global__ void __launch_bounds__(kGB)
gd_dot_kernel(const GmresDev G, const VecSlice sl) {
  __shared__ double sh[32];
  __shared__ bool s_last;
  const int j = (int)G.st[kStJ];
  const int nvec = j + 1;   // V_0..V_j
  const int ngroups = (nvec + kGK - 1) / kGK;
  const double* __restrict__ w = G.w;
  for (int grp = 0; grp < ngroups; ++grp) {
    const int k0 = grp * kGK;
    const int nk = min(kGK, nvec - k0);
    const double* Vb = G.V + (size_t)k0 * G.ne;
    double acc[kGK];
#pragma unroll
    for (int k = 0; k < kGK; ++k) acc[k] = 0.0;
    switch (nk) {   // uniform over the grid
      case 1: gd_dot_group<1, SLICE>(G, sl, Vb, w, acc); break;
      case 2: gd_dot_group<2, SLICE>(G, sl, Vb, w, acc); break;
      case 3: gd_dot_group<3, SLICE>(G, sl, Vb, w, acc); break;
      case 4: gd_dot_group<4, SLICE>(G, sl, Vb, w, acc); break;
      case 5: gd_dot_group<5, SLICE>(G, sl, Vb, w, acc); break;
      case 6: gd_dot_group<6, SLICE>(G, sl, Vb, w, acc); break;
      case 7: gd_dot_group<7, SLICE>(G, sl, Vb, w, acc); break;
      default: gd_dot_group<8, SLICE>(G, sl, Vb, w, acc); break;
    }
#pragma unroll
    for (int k = 0; k < kGK; ++k) {
      if (k >= nk) break;
      const double s = gd_block_sum(acc[k], sh);
      if (threadIdx.x == 0) G.partial[((size_t)grp * gridDim.x + blockIdx.x) * kGK + k] = s;
    }
  }
  // the last CTA to finish adds the partials up in CTA order
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(G.tickets + 0, 1u);
    s_last = t == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int kk = 0; kk < nvec; ++kk) {
    const int grp = kk / kGK, k = kk - grp * kGK;
    double v = 0.0;
    for (int b = threadIdx.x; b < (int)gridDim.x; b += kGB)
      v += *reinterpret_cast<const volatile double*>(G.partial + ((size_t)grp * gridDim.x + b) * kGK + k);
    const double s = gd_block_sum(v, sh);
    if (threadIdx.x == 0) G.dots[kk] = s * G.vscale[kk];
  }
  if (threadIdx.x == 0) G.tickets[0] = 0u;
}

// V_{j+1} = w - sum_{k<=j} h_k vscale_k V_k (unnormalised), nrm[0] = ||V_{j+1}||^2 over this rank's part
// ru != nullptr: also ru = D^-1 V_{j+1} (D = diag(sp | sq), the row equilibration; split layout [p | q] with
// nhalf entries each): the input of the next preconditioner application, still unnormalised -- the
// preconditioner is linear, the operator kernel multiplies its result by vscale_{j+1} (cm_stencil.cu)
template <bool SLICE>
__global__ void __launch_bounds__(kGB)
gd_update_kernel(const GmresDev G, const VecSlice sl, const double* __restrict__ sp, const double* __restrict__ sq,
                 double* __restrict__ ru, const long long nhalf) {
  extern __shared__ double s_h[];   // m + 2
  __shared__ double sh[32];
  __shared__ bool s_last;
  const int j = (int)G.st[kStJ];
  for (int k = threadIdx.x; k <= j; k += kGB) s_h[k] = G.dots[k] * G.vscale[k];
  __syncthreads();
  double nrm = 0.0;
  if (j + 1 <= G.m) {
    double* __restrict__ vout = G.V + (size_t)(j + 1) * G.ne;
    const double* w = G.w;
    if (SLICE) {
      auto one = [&](const long long q) {   // a single element (unaligned ends of a range)
        double a = w[q];
        for (int k = 0; k <= j; ++k) a -= s_h[k] * G.V[(size_t)k * G.ne + q];
        vout[q] = a;
        if (ru != nullptr) ru[q] = a / (q < nhalf ? sp[q] : sq[q - nhalf]);
        nrm += a * a;
      };
      for (int part = 0; part < 2; ++part) {
        const long long s0 = part ? sl.off1 : sl.off0;
        const long long lead = s0 & 1;
        const long long npair = (sl.len - lead) >> 1;
        const double* sc = part ? sq - nhalf : sp;   // scaling of this range, indexed like the vector
        for (long long i = blockIdx.x * (long long)kGB + threadIdx.x; i < npair; i += (long long)gridDim.x * kGB) {
          const long long q = s0 + lead + 2 * i;
          double2 a = *reinterpret_cast<const double2*>(w + q);
          const double* vq = G.V + q;
          int k = 0;
          for (; k + 4 <= j + 1; k += 4) {   // four independent 16-byte loads in flight
            const double2 v0 = *reinterpret_cast<const double2*>(vq + (size_t)k * G.ne);
            const double2 v1 = *reinterpret_cast<const double2*>(vq + (size_t)(k + 1) * G.ne);
            const double2 v2 = *reinterpret_cast<const double2*>(vq + (size_t)(k + 2) * G.ne);
            const double2 v3 = *reinterpret_cast<const double2*>(vq + (size_t)(k + 3) * G.ne);
            a.x -= s_h[k] * v0.x + s_h[k + 1] * v1.x + s_h[k + 2] * v2.x + s_h[k + 3] * v3.x;
            a.y -= s_h[k] * v0.y + s_h[k + 1] * v1.y + s_h[k + 2] * v2.y + s_h[k + 3] * v3.y;
          }
          for (; k <= j; ++k) {
            const double2 vv = *reinterpret_cast<const double2*>(vq + (size_t)k * G.ne);
            a.x -= s_h[k] * vv.x;
            a.y -= s_h[k] * vv.y;
          }
          *reinterpret_cast<double2*>(vout + q) = a;
          if (ru != nullptr) {
            double2 o;
            o.x = a.x / sc[q];
            o.y = a.y / sc[q + 1];
            *reinterpret_cast<double2*>(ru + q) = o;
          }
          nrm += a.x * a.x + a.y * a.y;
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) {
          if (lead) one(s0);
          if ((sl.len - lead) & 1) one(s0 + sl.len - 1);
        }
      }
    } else {
      const long long n2 = G.n >> 1;
      const double2* w2 = reinterpret_cast<const double2*>(w);
      for (long long i = blockIdx.x * (long long)kGB + threadIdx.x; i < n2; i += (long long)gridDim.x * kGB) {
        double2 a = w2[i];
        const double2* vq = reinterpret_cast<const double2*>(G.V) + i;
        const size_t ne2 = (size_t)(G.ne >> 1);
        int k = 0;
        for (; k + 4 <= j + 1; k += 4) {   // four independent 16-byte loads in flight
          const double2 v0 = vq[(size_t)k * ne2], v1 = vq[(size_t)(k + 1) * ne2];
          const double2 v2 = vq[(size_t)(k + 2) * ne2], v3 = vq[(size_t)(k + 3) * ne2];
          a.x -= s_h[k] * v0.x + s_h[k + 1] * v1.x + s_h[k + 2] * v2.x + s_h[k + 3] * v3.x;
          a.y -= s_h[k] * v0.y + s_h[k + 1] * v1.y + s_h[k + 2] * v2.y + s_h[k + 3] * v3.y;
        }
        for (; k <= j; ++k) {
          const double2 vv = vq[(size_t)k * ne2];
          a.x -= s_h[k] * vv.x;
          a.y -= s_h[k] * vv.y;
        }
        reinterpret_cast<double2*>(vout)[i] = a;
        if (ru != nullptr) {
          const long long e0 = 2 * i, e1 = 2 * i + 1;
          double2 o;
          o.x = a.x / (e0 < nhalf ? sp[e0] : sq[e0 - nhalf]);
          o.y = a.y / (e1 < nhalf ? sp[e1] : sq[e1 - nhalf]);
          reinterpret_cast<double2*>(ru)[i] = o;
        }
        nrm += a.x * a.x + a.y * a.y;
      }
    }
  }
  const double bs = gd_block_sum(nrm, sh);
  if (threadIdx.x == 0) G.partial[blockIdx.x] = bs;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(G.tickets + 1, 1u);
    s_last = t == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  double v = 0.0;
  for (int b = threadIdx.x; b < (int)gridDim.x; b += kGB) v += *reinterpret_cast<const volatile double*>(G.partial + b);
  const double tot = gd_block_sum(v, sh);
  if (threadIdx.x == 0) {
    G.nrm[0] = tot;
    G.tickets[1] = 0u;
  }
}

// the GMRES cycle as a WHILE node of a CUDA graph: another iteration?
__device__ __forceinline__ void gd_set_continue(const GmresDev& G, const unsigned long long handle) {
  if (handle == 0ull) return;
  const double* st = G.st;
  const bool cont = st[kStConv] == 0.0 && st[kStFlag] == 0.0 && st[kStIts] < st[kStMaxit] && st[kStJ] < (double)G.m;
  cudaGraphSetConditional((cudaGraphConditionalHandle)handle, cont ? 1u : 0u);
}

__global__ void gd_cond_kernel(const GmresDev G, const unsigned long long handle) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  gd_set_continue(G, handle);
}

// scalar part of an iteration, one thread: nrm[0] = ||V_{j+1}||^2 (summed over the ranks), dots = column j
__device__ __forceinline__ void gd_givens_body(const GmresDev& G, const double rtol, const double atol) {
  double* st = G.st;
  const int j = (int)st[kStJ], m = G.m;
  const double hn2 = G.nrm[0];
  const double hn = sqrt(fmax(hn2, 0.0));
  if (!isfinite(hn2)) {
    st[kStFlag] = 2.0;   // NaN / Inf in the Arnoldi process
    st[kStConv] = 0.0;
    st[kStJ] = (double)(j + 1);
    return;
  }
  if (j + 1 <= m) G.vscale[j + 1] = hn > 0.0 ? 1.0 / hn : 0.0;
  if (j < 0) {   // start of a cycle: w held the residual, v_0 = w / beta
    if (st[kStBnorm] < 0.0) {
      st[kStBnorm] = hn;
      st[kStTarget] = fmax(rtol * hn, atol);
    }
    G.g[0] = hn;
    st[kStResid] = hn;
    st[kStConv] = (hn <= st[kStTarget]) ? 1.0 : 0.0;
    st[kStJ] = 0.0;
    st[kStK] = 0.0;
    return;
  }
  double* Hc = G.H + (size_t)j * (m + 1);
  for (int i = 0; i <= j; ++i) Hc[i] = G.dots[i];
  Hc[j + 1] = hn;
  for (int i = 0; i < j; ++i) {
    const double t = G.cs[i] * Hc[i] + G.sn[i] * Hc[i + 1];
    Hc[i + 1] = -G.sn[i] * Hc[i] + G.cs[i] * Hc[i + 1];
    Hc[i] = t;
  }
  const double d = hypot(Hc[j], Hc[j + 1]);
  const double c = (d == 0.0) ? 1.0 : Hc[j] / d, s = (d == 0.0) ? 0.0 : Hc[j + 1] / d;
  G.cs[j] = c;
  G.sn[j] = s;
  Hc[j] = d;
  Hc[j + 1] = 0.0;
  G.g[j + 1] = -s * G.g[j];
  G.g[j] = c * G.g[j];
  const double resid = fabs(G.g[j + 1]);
  st[kStResid] = resid;
  st[kStIts] += 1.0;
  st[kStJ] = (double)(j + 1);
  st[kStK] = (double)(j + 1);
  if (resid <= st[kStTarget] || hn <= 1e-300) st[kStConv] = 1.0;   // converged, or lucky breakdown
}

__global__ void gd_givens_kernel(const GmresDev G, const double rtol, const double atol, const unsigned long long handle) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  gd_givens_body(G, rtol, atol);
  gd_set_continue(G, handle);
}

// start of a solve / of a restart cycle: j = -1 (the next dot + update pair normalises the residual in w)
__global__ void gd_begin_kernel(const GmresDev G, const int first, const int maxit) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double* st = G.st;
  st[kStJ] = -1.0;
  st[kStFlag] = 0.0;
  st[kStConv] = 0.0;
  st[kStK] = 0.0;
  if (first) {
    st[kStMaxit] = (double)maxit;
    st[kStIts] = 0.0;
    st[kStBnorm] = -1.0;
    st[kStTarget] = 0.0;
    st[kStResid] = 0.0;
    G.tickets[0] = 0u;
    G.tickets[1] = 0u;
  }
}

// y = R^{-1} g for the k = st[kStK] columns of the cycle -> coef
__global__ void gd_solve_y_kernel(const GmresDev G) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int k = (int)G.st[kStK], m = G.m;
  for (int i = k - 1; i >= 0; --i) {
    double s = G.g[i];
    for (int l = i + 1; l < k; ++l) s -= G.H[(size_t)l * (m + 1) + i] * G.coef[l];
    G.coef[i] = s / G.H[(size_t)i * (m + 1) + i];
  }
}

// out (+)= sum_{k < K} coef_k B_k, K = st[kStK];  B = Z (the preconditioned basis: x += Z y) or, with
// use_vscale, the unnormalised Krylov basis V
template <bool SLICE>
__global__ void __launch_bounds__(kGB)
gd_combine_kernel(const GmresDev G, const VecSlice sl, const double* __restrict__ B, const int use_vscale,
                  const int accumulate, double* __restrict__ out) {
  extern __shared__ double s_h[];
  const int K = (int)G.st[kStK];
  for (int k = threadIdx.x; k < K; k += kGB) s_h[k] = G.coef[k] * (use_vscale ? G.vscale[k] : 1.0);
  __syncthreads();
  const long long n = SLICE ? 2 * sl.len : G.n;
  for (long long i = blockIdx.x * (long long)kGB + threadIdx.x; i < n; i += (long long)gridDim.x * kGB) {
    const long long q = SLICE ? gd_index(i, sl) : i;
    double a = accumulate ? out[q] : 0.0;
    int k = 0;
    for (; k + 4 <= K; k += 4)
      a += s_h[k] * B[(size_t)k * G.ne + q] + s_h[k + 1] * B[(size_t)(k + 1) * G.ne + q] +
           s_h[k + 2] * B[(size_t)(k + 2) * G.ne + q] + s_h[k + 3] * B[(size_t)(k + 3) * G.ne + q];
    for (; k < K; ++k) a += s_h[k] * B[(size_t)k * G.ne + q];
    out[q] = a;
  }
}

// y = alpha*x + beta*y on the vector (x may be null: y = beta*y; beta == 0 never reads y)
template <bool SLICE>
__global__ void __launch_bounds__(kGB)
gd_axpby_kernel(const long long ntot, const VecSlice sl, const double alpha, const double* __restrict__ x,
                const double beta, double* __restrict__ y) {
  const long long n = SLICE ? 2 * sl.len : ntot;
  for (long long i = blockIdx.x * (long long)kGB + threadIdx.x; i < n; i += (long long)gridDim.x * kGB) {
    const long long q = SLICE ? gd_index(i, sl) : i;
    double v = (beta != 0.0) ? beta * y[q] : 0.0;
    if (x != nullptr) v += alpha * x[q];
    y[q] = v;
  }
}

static int gd_grid(long long nelem) {
  long long g = (nelem + kGB - 1) / kGB;
  if (g < 1) g = 1;
  return (int)(g < kGGrid ? g : kGGrid);
}
static long long gd_len(const GmresDev& G, const VecSlice* sl) { return sl ? 2 * sl->len : G.n; }

int gmres_dev_begin(const GmresDev& G, int first, cudaStream_t st, int maxit) {
  gd_begin_kernel<<<1, 32, 0, st>>>(G, first, maxit);
  CM_LAUNCH_CHECK();
  return CM_OK;
}

// bytes: every basis vector in use + w once; the iteration index is on the device, the caller passes the
// value it knows (jhost) for the byte accounting only.  jhost < 0 (start of a cycle): nothing to do.
// CTAs of the dot kernel that are resident at once (78 registers: 3 per SM, not the 4 of kGGrid): a grid-stride
// loop over more CTAs than fit runs a second, mostly empty wave (ncu: 31 % warps active, 60 % DRAM throughput)
template <bool SLICE>
static int gd_dot_resident_grid() {
  static std::atomic<int> cached{0};
  int g = cached.load();
  if (g > 0) return g;
  int per_sm = 0, dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gd_dot_kernel<SLICE>, kGB, 0) != cudaSuccess || per_sm < 1) {
    cudaGetLastError();
    per_sm = 2;
  }
  g = per_sm * sms;
  if (g > kGGrid) g = kGGrid;
  cached.store(g);
  return g;
}

// once per process, outside any stream capture (cm_pqs_create)
void gmres_dev_init() {
  gd_dot_resident_grid<true>();
  gd_dot_resident_grid<false>();
}

int gmres_dev_dot(const GmresDev& G, const VecSlice* sl, int jhost, cudaStream_t st) {
  if (jhost < 0) return CM_OK;
  const long long len = gd_len(G, sl);
  int grid = gd_grid(sl ? len : len / 2);
  const int resident = sl ? gd_dot_resident_grid<true>() : gd_dot_resident_grid<false>();
  if (grid > resident) grid = resident;
  if (sl) gd_dot_kernel<true><<<grid, kGB, 0, st>>>(G, *sl);
  else gd_dot_kernel<false><<<grid, kGB, 0, st>>>(G, VecSlice{0, 0, 0});
  CM_BYTES((double)len * 8.0 * (jhost + 2));
  CM_LAUNCH_CHECK();
  return CM_OK;
}

int gmres_dev_update(const GmresDev& G, const VecSlice* sl, int jhost, cudaStream_t st, const double* sp,
                     const double* sq, double* ru) {
  const long long len = gd_len(G, sl);
  const int grid = gd_grid(sl ? len : len / 2);
  const size_t smem = (size_t)(G.m + 2) * sizeof(double);
  if (ru != nullptr && (G.n & 1)) {
    set_error("gmres_dev_update: the fused unscaling needs an even vector length");
    return CM_E_INVALID;
  }
  if (sl) gd_update_kernel<true><<<grid, kGB, smem, st>>>(G, *sl, sp, sq, ru, G.n / 2);
  else gd_update_kernel<false><<<grid, kGB, smem, st>>>(G, VecSlice{0, 0, 0}, sp, sq, ru, G.n / 2);
  CM_BYTES((double)len * 8.0 * (jhost + 3 + (ru ? 2 : 0)));
  CM_LAUNCH_CHECK();
  return CM_OK;
}

int gmres_dev_givens(const GmresDev& G, double rtol, double atol, cudaStream_t st, unsigned long long cond_handle) {
  gd_givens_kernel<<<1, 32, 0, st>>>(G, rtol, atol, cond_handle);
  CM_LAUNCH_CHECK();
  return CM_OK;
}

int gmres_dev_add_cond_node(cudaGraph_t graph, cudaGraphNode_t* node, const GmresDev& G, unsigned long long cond_handle) {
  GmresDev Gc = G;
  unsigned long long h = cond_handle;
  void* args[2] = {&Gc, &h};
  cudaKernelNodeParams kp{};
  kp.func = (void*)gd_cond_kernel;
  kp.gridDim = dim3(1, 1, 1);
  kp.blockDim = dim3(32, 1, 1);
  kp.sharedMemBytes = 0;
  kp.kernelParams = args;
  kp.extra = nullptr;
  CM_CUDA(cudaGraphAddKernelNode(node, graph, nullptr, 0, &kp));
  return CM_OK;
}

int gmres_dev_solve_y(const GmresDev& G, cudaStream_t st) {
  gd_solve_y_kernel<<<1, 32, 0, st>>>(G);
  CM_LAUNCH_CHECK();
  return CM_OK;
}

// x += Z y (the solution update of the flexible variant: no further preconditioner application)
int gmres_dev_combine(const GmresDev& G, const VecSlice* sl, double* out, int khost, cudaStream_t st) {
  const long long len = gd_len(G, sl);
  const int grid = gd_grid(len);
  const size_t smem = (size_t)(G.m + 2) * sizeof(double);
  if (sl) gd_combine_kernel<true><<<grid, kGB, smem, st>>>(G, *sl, G.Z, 0, 1, out);
  else gd_combine_kernel<false><<<grid, kGB, smem, st>>>(G, VecSlice{0, 0, 0}, G.Z, 0, 1, out);
  CM_BYTES((double)len * 8.0 * (khost + 2));
  CM_LAUNCH_CHECK();
  return CM_OK;
}

int gmres_dev_axpby(const GmresDev& G, const VecSlice* sl, double alpha, const double* x, double beta, double* y,
                    cudaStream_t st) {
  const long long len = gd_len(G, sl);
  const int grid = gd_grid(len);
  if (sl) gd_axpby_kernel<true><<<grid, kGB, 0, st>>>(G.n, *sl, alpha, x, beta, y);
  else gd_axpby_kernel<false><<<grid, kGB, 0, st>>>(G.n, VecSlice{0, 0, 0}, alpha, x, beta, y);
  CM_BYTES((double)len * 8.0 * ((x ? 1 : 0) + (beta != 0.0 ? 2 : 1)));
  CM_LAUNCH_CHECK();
  return CM_OK;
}

}  // namespace cm
