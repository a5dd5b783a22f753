// Shared-memory resident V-cycle for the coarsest levels of the multigrid hierarchy: ONE CTA, every phase
// separated by __syncthreads, EVERYTHING the cycle touches resident in shared memory for its whole duration:
// x and b of every level, the four off-line diagonals and the three line factors of every level, the dense
// inverse of the coarsest operator.  The arrays arrive by cp.async (global -> shared without registers) in two
// groups: the first level's right-hand side and operator, then the rest, which streams in behind the first
// level's pre-smoothing.  After that no phase waits for L2 again: a phase costs the shared-memory latency of one
// warp-wide line solve (~0.4 us) instead of the 3-4 us of a kernel boundary, a cluster barrier with global
// round trips, or an L2 refill of the operator (the first version of this kernel reloaded the operator at every
// level change and took 70 us per cycle, in-situ timeline profiles/r2_trace_n1_mid.csv; this one ~4x less).
// Two identities keep the tridiagonal diagonals out of shared memory and remove two phases per level:
//   * right after a zebra sweep the residual on an even row is  -(A - T)(x_odd_new - x_odd_old)  (the line solve
//     of the colour-0 half sweep made T x_even = b_even - (A - T) x_odd_old exact), so the restricted residual
//     needs the four off-line diagonals and the change of the odd rows only;
//   * the coarse-grid correction is only ever read through the odd rows next to the first colour-0 sweep of
//     the up leg, so x_odd + P e is formed on the fly there and never stored.
// Replaces the smoother / coarse solve part of the sparse LU the reference requests per Newton step
// (advection_diffusion_convergence.py:212, ...MixedPrimal_convergence_SUPG.py:137).
#include <stdlib.h>

#include "cm_common.cuh"
#include "cm_decl.cuh"
#include "cm_mg.cuh"

namespace cm {

namespace {

constexpr int kSmT = 512;   // 16 warps, 128 registers (more threads cap the registers at 96 and spill)
constexpr int kSmE = 3;     // unknowns per lane: lines of up to 96 unknowns
constexpr size_t kSmMaxBytes = 227 * 1024 - 768;   // dynamic shared memory this kernel may ask for

struct SmLevel {   // one level in shared memory
  int nx, ny, n;
  double *x, *b;
  const double *a0, *a1, *a5, *a6, *lf, *iu, *cu;
};

struct AffS {
  double a, b;
};
__device__ __forceinline__ AffS affs_compose(const AffS later, const AffS earlier) {
  return AffS{later.a * earlier.a, later.a * earlier.b + later.b};
}
// value entering this lane's segment (exclusive scan over the lanes of one warp, rev: from the last lane down)
__device__ __forceinline__ double affs_warp_scan(AffS m, const bool rev) {
  const int lane = threadIdx.x & 31;
  const int pos = rev ? 31 - lane : lane;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double pa = rev ? __shfl_down_sync(0xffffffffu, m.a, o) : __shfl_up_sync(0xffffffffu, m.a, o);
    const double pb = rev ? __shfl_down_sync(0xffffffffu, m.b, o) : __shfl_up_sync(0xffffffffu, m.b, o);
    if (pos >= o) m = affs_compose(m, AffS{pa, pb});
  }
  double eb = rev ? __shfl_down_sync(0xffffffffu, m.b, 1) : __shfl_up_sync(0xffffffffu, m.b, 1);
  if (pos == 0) eb = 0.0;
  return eb;
}

__device__ __forceinline__ void cp_async8(double* sdst, const double* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(sdst)), "l"(gsrc)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void sm_copy(double* dst, const double* src, const int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) cp_async8(dst + i, src + i);
}

// MODE 0: zero guess;  1: with couplings;  3: with couplings to x_odd + P e (e = correction of the next level)
// delta != nullptr: store x_new - x_old of the line (x_old = 0 when old_zero) at delta[(j >> 1) * n1 + i]
// (MODE is a run-time argument and the sweep is NOT inlined: the cycle calls it from ~12 places, and one
// straight-line copy per call site made 290 KB of code that a single CTA walks through once -- instruction
// fetches from L2 dominated the first version of this rewrite)
__device__ __forceinline__ void sm_line_solve(const int MODE, const SmLevel& L, const int j, const double* e, const int m1,
                                              double* delta, const bool old_zero) {
  const int n1 = L.nx + 1, nx = L.nx, ny = L.ny, lane = threadIdx.x & 31;
  const int base = j * n1;
  const int E = (n1 + 31) >> 5;
  const int i0 = lane * E;
  double r[kSmE], lv[kSmE], cv[kSmE], iv[kSmE];
  const int lo = (j > 0 ? j - 1 : 0) * n1, up = (j < ny ? j + 1 : ny) * n1;
  const double* x = L.x;
#pragma unroll
  for (int k = 0; k < kSmE; ++k) {
    const int i = i0 + k;
    if (k < E && i < n1) {
      double s = L.b[base + i];
      if (MODE != 0) {
        const int im = i > 0 ? i - 1 : 0, ip = i < nx ? i + 1 : nx;
        if (j > 0) {
          double x0 = x[lo + im], x1 = x[lo + i];
          if (MODE == 3) {
            const int c = ((j >> 1) - 1) * m1;
            x0 += 0.5 * (e[c + (im >> 1)] + e[c + (im >> 1) + m1 + (im & 1)]);
            x1 += 0.5 * (e[c + (i >> 1)] + e[c + (i >> 1) + m1 + (i & 1)]);
          }
          s -= L.a0[base + i] * x0 + L.a1[base + i] * x1;
        }
        if (j < ny) {
          double x0 = x[up + i], x1 = x[up + ip];
          if (MODE == 3) {
            const int c = (j >> 1) * m1;
            x0 += 0.5 * (e[c + (i >> 1)] + e[c + (i >> 1) + m1 + (i & 1)]);
            x1 += 0.5 * (e[c + (ip >> 1)] + e[c + (ip >> 1) + m1 + (ip & 1)]);
          }
          s -= L.a5[base + i] * x0 + L.a6[base + i] * x1;
        }
      }
      r[k] = s;
      lv[k] = L.lf[base + i];
      cv[k] = L.cu[base + i];
      iv[k] = L.iu[base + i];
    }
  }
  AffS m{1.0, 0.0};
#pragma unroll
  for (int k = 0; k < kSmE; ++k)
    if (k < E && i0 + k < n1) { m.b = r[k] - lv[k] * m.b; m.a = -lv[k] * m.a; }
  double z = affs_warp_scan(m, false);
#pragma unroll
  for (int k = 0; k < kSmE; ++k)
    if (k < E && i0 + k < n1) { z = r[k] - lv[k] * z; r[k] = z * iv[k]; }
  m = AffS{1.0, 0.0};
#pragma unroll
  for (int k = kSmE - 1; k >= 0; --k)
    if (k < E && i0 + k < n1) { m.b = r[k] - cv[k] * m.b; m.a = -cv[k] * m.a; }
  double xn = affs_warp_scan(m, true);
#pragma unroll
  for (int k = kSmE - 1; k >= 0; --k)
    if (k < E && i0 + k < n1) {
      xn = r[k] - cv[k] * xn;
      if (delta != nullptr) delta[(j >> 1) * n1 + i0 + k] = old_zero ? xn : xn - L.x[base + i0 + k];
      L.x[base + i0 + k] = xn;
    }
}

__device__ __noinline__ void sm_half_sweep(const int MODE, const SmLevel* Lp, const int colour, const double* e,
                                          const int m1, double* delta, const bool old_zero) {
  const SmLevel L = *Lp;
  const int nw = blockDim.x >> 5, w = threadIdx.x >> 5;
  for (int j = colour + 2 * w; j <= L.ny; j += 2 * nw) sm_line_solve(MODE, L, j, e, m1, delta, old_zero);
  __syncthreads();
}

// optional phase stamps (%globaltimer, thread 0) for tools/time_smem_cycle.py
__device__ unsigned long long g_smem_dbg[64];
__device__ int g_smem_dbg_n;
#define SMT()                                                                           \
  do {                                                                                  \
    if (C.dbg && threadIdx.x == 0 && g_smem_dbg_n < 64) g_smem_dbg[g_smem_dbg_n++] = global_ns(); \
  } while (0)

__global__ void __launch_bounds__(kSmT, 1)
coarse_smem_kernel(const CoarseCycle C) {
  extern __shared__ __align__(16) double sm[];
  __shared__ SmLevel lev[kCoarseMaxLevels];
  __shared__ double* s_delta;
  __shared__ double* s_ainv;
  const int last = C.nlev - 1;
  const bool dense = C.Ainv != nullptr;
  if (C.dbg && threadIdx.x == 0) g_smem_dbg_n = 0;
  SMT();   // 0: start
  if (threadIdx.x == 0) {
    double* p = sm;
    for (int k = 0; k <= last; ++k) {
      SmLevel& L = lev[k];
      L.nx = C.lev[k].nx; L.ny = C.lev[k].ny;
      L.n = (L.nx + 1) * (L.ny + 1);
      L.x = p; p += L.n;
      L.b = p; p += L.n;
      if (k < last || !dense) {
        L.a0 = p; L.a1 = p + L.n; L.a5 = p + 2 * L.n; L.a6 = p + 3 * L.n;
        L.lf = p + 4 * L.n; L.iu = p + 5 * L.n; L.cu = p + 6 * L.n;
        p += 7 * L.n;
      }
    }
    s_delta = p;
    p += (C.lev[0].ny / 2) * (C.lev[0].nx + 1);
    s_ainv = (dense && C.ainv_smem) ? p : nullptr;
  }
  __syncthreads();
  // ---- everything the cycle reads, in two cp.async groups: level 0 first
  auto fetch_op = [&](const int k) {
    const CoarseLevel& G = C.lev[k];
    const SmLevel& L = lev[k];
    const size_t nv = (size_t)L.n;
    sm_copy(const_cast<double*>(L.a0), G.A, L.n);
    sm_copy(const_cast<double*>(L.a1), G.A + nv, L.n);
    sm_copy(const_cast<double*>(L.a5), G.A + 5 * nv, L.n);
    sm_copy(const_cast<double*>(L.a6), G.A + 6 * nv, L.n);
    sm_copy(const_cast<double*>(L.lf), G.lf, L.n);
    sm_copy(const_cast<double*>(L.iu), G.iu, L.n);
    sm_copy(const_cast<double*>(L.cu), G.cu, L.n);
  };
  sm_copy(lev[0].b, C.lev[0].b, lev[0].n);   // written by the previous kernel of the stream
  if (last > 0 || !dense) fetch_op(0);
  cp_async_commit();
  for (int k = 1; k <= last; ++k)
    if (k < last || !dense) fetch_op(k);
  if (s_ainv != nullptr) sm_copy(s_ainv, C.Ainv, lev[last].n * lev[last].n);
  cp_async_commit();
  SMT();   // 1: copies issued
  cp_async_wait<1>();
  __syncthreads();
  SMT();   // 2: first level resident
  double* delta = s_delta;
  // ---- down leg
  for (int k = 0; k < last; ++k) {
    const SmLevel& L = lev[k];
    const SmLevel& N = lev[k + 1];
    for (int s = 0; s < C.pre; ++s) {
      const bool fin = s + 1 == C.pre;
      if (s == 0) sm_half_sweep(0, &L, 0, nullptr, 0, nullptr, false);
      else sm_half_sweep(1, &L, 0, nullptr, 0, nullptr, false);
      sm_half_sweep(1, &L, 1, nullptr, 0, fin ? delta : nullptr, s == 0);
    }
    SMT();   // pre-smoothing of level k done
    if (k == 0) {   // the other levels have arrived behind the first pre-smoothing
      cp_async_wait<0>();
      __syncthreads();
    }
    {   // restricted residual: r_even = -(A - T) delta_odd
      const int n1 = L.nx + 1, m1 = N.nx + 1, nx = L.nx, ny = L.ny;
      const int ncoarse = m1 * (N.ny + 1);
      for (int I = threadIdx.x; I < ncoarse; I += blockDim.x) {
        const int X = I % m1, Y = I / m1;
        const int iy = 2 * Y;
        const double* dlo = delta + (Y - 1) * n1;   // row 2Y - 1
        const double* dup = delta + Y * n1;         // row 2Y + 1
        double acc = 0.0;
#pragma unroll
        for (int d = -1; d <= 1; ++d) {
          const int i = 2 * X + d;
          if (i < 0 || i > nx) continue;
          const int v = iy * n1 + i;
          double s = 0.0;
          if (iy > 0) {
            if (i > 0) s -= L.a0[v] * dlo[i - 1];
            s -= L.a1[v] * dlo[i];
          }
          if (iy < ny) {
            s -= L.a5[v] * dup[i];
            if (i < nx) s -= L.a6[v] * dup[i + 1];
          }
          acc += (d == 0) ? s : 0.5 * s;
        }
        N.b[I] = acc;
      }
    }
    __syncthreads();
    SMT();   // restricted residual of level k done
  }
  // ---- coarsest level
  if (last == 0) {
    cp_async_wait<0>();
    __syncthreads();
  }
  {
    const SmLevel& L = lev[last];
    if (dense) {
      const double* Ai = s_ainv != nullptr ? s_ainv : C.Ainv;
      const int n = L.n, w = threadIdx.x >> 5, nw = blockDim.x >> 5, lane = threadIdx.x & 31;
      for (int r = w; r < n; r += nw) {   // one warp per row: coalesced, one reduction
        double s = 0.0;
        for (int c = lane; c < n; c += 32) s += Ai[(size_t)r * n + c] * L.b[c];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (lane == 0) L.x[r] = s;
      }
      __syncthreads();
    } else {
      for (int s = 0; s < 30; ++s) {
        if (s == 0) sm_half_sweep(0, &L, 0, nullptr, 0, nullptr, false);
        else sm_half_sweep(1, &L, 0, nullptr, 0, nullptr, false);
        sm_half_sweep(1, &L, 1, nullptr, 0, nullptr, false);
      }
    }
  }
  SMT();   // coarsest level done
  // ---- up leg: the correction enters through the first colour-0 half sweep
  for (int k = last - 1; k >= 0; --k) {
    const SmLevel& L = lev[k];
    const SmLevel& N = lev[k + 1];
    for (int s = 0; s < C.post; ++s) {
      if (s == 0) sm_half_sweep(3, &L, 0, N.x, N.nx + 1, nullptr, false);
      else sm_half_sweep(1, &L, 0, nullptr, 0, nullptr, false);
      sm_half_sweep(1, &L, 1, nullptr, 0, nullptr, false);
    }
    SMT();   // post-smoothing of level k done
  }
  {
    const SmLevel& L = lev[0];
    double* xo = C.lev[0].x;
    for (int i = threadIdx.x; i < L.n; i += blockDim.x) xo[i] = L.x[i];
  }
  SMT();   // stored
}

// doubles of shared memory without the optional copy of the dense inverse (0: a line is too long)
size_t smem_core_doubles(const CoarseCycle& C, bool dense_last) {
  size_t d = 0;
  for (int k = 0; k < C.nlev; ++k) {
    if (C.lev[k].nx + 1 > 32 * kSmE) return 0;
    const size_t n = (size_t)(C.lev[k].nx + 1) * (C.lev[k].ny + 1);
    d += 2 * n;
    if (k < C.nlev - 1 || !dense_last) d += 7 * n;
  }
  d += (size_t)(C.lev[0].ny / 2) * (C.lev[0].nx + 1);
  return d;
}

}  // namespace

// shared memory the resident cycle needs for the levels of C (0: it does not apply)
size_t mg_coarse_smem_bytes(const CoarseCycle& C, bool dense_last) {
  const size_t d = smem_core_doubles(C, dense_last);
  if (d == 0 || (C.nlev > 1 && (C.lev[0].ny & 1))) return 0;
  const size_t bytes = d * sizeof(double);
  return bytes <= kSmMaxBytes ? bytes : 0;
}

int mg_coarse_smem_init_attributes() {
  static std::atomic<bool> done{false};
  if (done.load()) return CM_OK;
  CM_CUDA(cudaFuncSetAttribute(coarse_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmMaxBytes));
  done.store(true);
  return CM_OK;
}

int mg_coarse_smem_cycle(const CoarseCycle& C0, cudaStream_t st) {
  CoarseCycle C = C0;
  const bool dense = C.Ainv != nullptr;
  size_t smem = mg_coarse_smem_bytes(C, dense);
  if (smem == 0) {
    set_error("mg_coarse_smem_cycle: the levels do not fit in shared memory");
    return CM_E_UNSUPPORTED;
  }
  if (C.pre < 1 || C.post < 1) {
    set_error("mg_coarse_smem_cycle: needs at least one pre- and one post-smoothing sweep");
    return CM_E_UNSUPPORTED;
  }
  int rc = mg_coarse_smem_init_attributes();
  if (rc) return rc;
  C.ainv_smem = 0;
  if (dense) {
    const size_t nl = (size_t)(C.lev[C.nlev - 1].nx + 1) * (C.lev[C.nlev - 1].ny + 1);
    if (smem + nl * nl * sizeof(double) <= kSmMaxBytes) {
      C.ainv_smem = 1;
      smem += nl * nl * sizeof(double);
    }
  }
  double bytes = 0.0;
  for (int k = 0; k < C.nlev; ++k)
    bytes += (double)(C.lev[k].nx + 1) * (C.lev[k].ny + 1) * 8.0 * (k == 0 ? 9 : 7);   // 4 + 3 operator arrays, b in / x out
  CM_BYTES(bytes);
  if (const char* e = getenv("CM_COARSE_DBG")) C.dbg = atoi(e);
  coarse_smem_kernel<<<1, kSmT, smem, st>>>(C);
  CM_LAUNCH_CHECK();
  return CM_OK;
}

int mg_smem_debug_times(unsigned long long* host_out, int n) {
  int cnt = 0;
  if (cudaMemcpyFromSymbol(&cnt, g_smem_dbg_n, sizeof(int)) != cudaSuccess) return CM_E_CUDA;
  if (cudaMemcpyFromSymbol(host_out, g_smem_dbg, sizeof(unsigned long long) * (n < 64 ? n : 64)) != cudaSuccess)
    return CM_E_CUDA;
  return cnt;
}

}  // namespace cm
