// Second-generation kernels of the structured multigrid preconditioner (cm_pqsolver.cu), built to shorten
// the dependent-launch chain of one preconditioner application and to move every array once:
//   * zebra_ring_kernel   one launch per zebra half sweep on long lines: the couplings to the neighbouring
//                         lines, the tridiagonal solve and (multi-GPU) the push of the strip-boundary line
//                         into the neighbour's memory.  All ten input streams of a line (b, 4 off-line
//                         diagonals, 2 neighbour lines of x, 3 factor arrays) arrive through a ring of
//                         shared-memory stages filled by cp.async.bulk (TMA) and tracked by mbarriers, so HBM
//                         keeps streaming while the CTA runs the two block-wide scans of the current line.
//   * zebra_small_kernel  same contract for short lines (one CTA per line).
//   * resid_restrict_kernel  after a sweep that ended with colour 1 the residual vanishes on the odd rows,
//                         so the restricted residual needs the even fine rows only: one pass, no fine
//                         residual vector.
//   * prolong_odd_kernel  the correction is added on the odd rows only: the even rows are overwritten by the
//                         colour-0 half sweep that follows without being read.
//   * coarse_cycle_kernel the whole V-cycle below a given level in ONE launch (a thread-block cluster,
//                         barrier.cluster between the phases instead of kernel boundaries).
// Replaces the smoother / transfer part of the sparse LU the reference requests per Newton step
// (advection_diffusion_convergence.py:212, ...MixedPrimal_convergence_SUPG.py:137).
#include <cooperative_groups.h>
#include <stdlib.h>

#include "cm_common.cuh"
#include "cm_decl.cuh"
#include "cm_halo.cuh"
#include "cm_mg.cuh"

namespace cg = cooperative_groups;

namespace cm {

// ------------------------------------------------------------------------------------------------
// scans over affine maps v_out = a*v_in + b (first-order linear recurrences)
// ------------------------------------------------------------------------------------------------
struct Aff {
  double a, b;
};
__device__ __forceinline__ Aff aff_compose(const Aff later, const Aff earlier) {
  return Aff{later.a * earlier.a, later.a * earlier.b + later.b};
}

// exclusive scan in thread order (rev: in reverse thread order) over the nw warps of the CTA; returns the
// value entering this thread's segment when the value entering the first segment is 0.  sh: 64 doubles.
__device__ __forceinline__ double aff_block_scan(Aff m, double* sh, const bool rev) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int pos = rev ? 31 - lane : lane;
  const int wpos = rev ? nw - 1 - wid : wid;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double pa = rev ? __shfl_down_sync(0xffffffffu, m.a, o) : __shfl_up_sync(0xffffffffu, m.a, o);
    const double pb = rev ? __shfl_down_sync(0xffffffffu, m.b, o) : __shfl_up_sync(0xffffffffu, m.b, o);
    if (pos >= o) m = aff_compose(m, Aff{pa, pb});
  }
  if (pos == 31) { sh[2 * wpos] = m.a; sh[2 * wpos + 1] = m.b; }
  __syncthreads();
  if (wid == 0) {
    Aff t = (lane < nw) ? Aff{sh[2 * lane], sh[2 * lane + 1]} : Aff{1.0, 0.0};
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double pa = __shfl_up_sync(0xffffffffu, t.a, o);
      const double pb = __shfl_up_sync(0xffffffffu, t.b, o);
      if (lane >= o) t = aff_compose(t, Aff{pa, pb});
    }
    if (lane < nw) { sh[2 * lane] = t.a; sh[2 * lane + 1] = t.b; }
  }
  __syncthreads();
  double ea = rev ? __shfl_down_sync(0xffffffffu, m.a, 1) : __shfl_up_sync(0xffffffffu, m.a, 1);
  double eb = rev ? __shfl_down_sync(0xffffffffu, m.b, 1) : __shfl_up_sync(0xffffffffu, m.b, 1);
  if (pos == 0) { ea = 1.0; eb = 0.0; }
  double r = eb;
  if (wpos > 0) r = ea * sh[2 * (wpos - 1) + 1] + eb;
  __syncthreads();
  return r;
}

// same inside one warp (short lines of the fused coarse cycle)
__device__ __forceinline__ double aff_warp_scan(Aff m, const bool rev) {
  const int lane = threadIdx.x & 31;
  const int pos = rev ? 31 - lane : lane;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double pa = rev ? __shfl_down_sync(0xffffffffu, m.a, o) : __shfl_up_sync(0xffffffffu, m.a, o);
    const double pb = rev ? __shfl_down_sync(0xffffffffu, m.b, o) : __shfl_up_sync(0xffffffffu, m.b, o);
    if (pos >= o) m = aff_compose(m, Aff{pa, pb});
  }
  double eb = rev ? __shfl_down_sync(0xffffffffu, m.b, 1) : __shfl_up_sync(0xffffffffu, m.b, 1);
  if (pos == 0) eb = 0.0;
  return eb;
}

// ------------------------------------------------------------------------------------------------
// TMA / mbarrier helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned s_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  unsigned done = 0;
  long long spins = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(s_u32(bar)), "r"(parity) : "memory");
    if (!done && ++spins > (1ll << 26)) __trap();   // a lost copy must not hang the GPU
  }
}
__device__ __forceinline__ void tma_g2s(void* smem_dst, const void* gsrc, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(s_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void group_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------------------------------------
// zebra half sweep on long lines.  Persistent CTAs; thread t of the consumer warps owns the kE consecutive
// unknowns [kE t, kE t + kE) of the line in registers.  The input streams of a line arrive ONE WHOLE LINE PER
// BULK COPY (few large cp.async.bulk operations: the copy engine has a per-operation cost of ~0.2 us, which
// rules out small tiles) through a ring of S line-sized shared-memory slots in a fixed order
//     x(lower line), A0, A1, x(upper line), A5, A6, b, lf, iu, cu          (b, lf, iu, cu from a zero guess)
// and are folded into the right-hand side as they arrive.  A dedicated producer warp issues the copies:
// full[s] (transaction barrier) tells the consumers that slot s holds stream n, empty[s] (one arrival per
// consumer warp) tells the producer that it has been read.  Together with every copy the producer pulls the
// same stream of the CTA's NEXT line into L2 (cp.async.bulk.prefetch.L2): HBM streams one line ahead of the
// ring, the ring itself refills from L2, and the two block-wide scans of the current line overlap both.
// Element i of a line sits at position i + 4 + s of its slot, s = parity of the line's first element address / 8
// (bulk copies need 16-byte aligned ends; the elements a copy cannot reach at the two ends of an array are
// fetched by plain loads).
// ------------------------------------------------------------------------------------------------
constexpr int kE = 5;          // odd: conflict-free 64-bit shared-memory accesses at stride kE
constexpr int kRingMaxT = 896; // consumer threads + the producer warp

// optional phase stamps (%globaltimer) of the first 64 CTAs x 16 lines, for tools/time_ring.py
__device__ unsigned long long g_ring_dbg[64 * 16 * 8];
#define RT(k)                                                                                   \
  do {                                                                                          \
    if (J.dbg && tid == 0 && blockIdx.x < 64 && it < 16)                                        \
      g_ring_dbg[((int)blockIdx.x * 16 + it) * 8 + (k)] = global_ns();                          \
  } while (0)

// stream a of a half sweep, in consumption order
template <bool COUPLED>
__device__ __forceinline__ const double* ring_base(const LineJob& J, const size_t nv, const int a) {
  if (COUPLED) {
    switch (a) {
      case 0: return J.x;            // lower line
      case 1: return J.A;            // diagonal 0: (-1,-1)
      case 2: return J.A + nv;       // diagonal 1: ( 0,-1)
      case 3: return J.x;            // upper line
      case 4: return J.A + 5 * nv;   // diagonal 5: ( 0,+1)
      case 5: return J.A + 6 * nv;   // diagonal 6: (+1,+1)
      case 6: return J.b;
      case 7: return J.lf;
      case 8: return J.iu;
      default: return J.cu;
    }
  }
  switch (a) {
    case 0: return J.b;
    case 1: return J.lf;
    case 2: return J.iu;
    default: return J.cu;
  }
}
template <bool COUPLED>
__device__ __forceinline__ int ring_row(const LineJob& J, const int a, const int j) {
  if (COUPLED && a == 0) return j > 0 ? j - 1 : 0;
  if (COUPLED && a == 3) return j < J.ny ? j + 1 : J.ny;
  return j;
}
__device__ __forceinline__ int addr_parity(const double* p) {
  return (int)((reinterpret_cast<unsigned long long>(p) >> 3) & 1ull);
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s_u32(bar)) : "memory");
}

// block-wide exclusive scan over the consumer warps only (named barrier 1; the producer warp is not part of it)
__device__ __forceinline__ double aff_consumer_scan(Aff m, double* sh, const bool rev, const int nw) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int pos = rev ? 31 - lane : lane;
  const int wpos = rev ? nw - 1 - wid : wid;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double pa = rev ? __shfl_down_sync(0xffffffffu, m.a, o) : __shfl_up_sync(0xffffffffu, m.a, o);
    const double pb = rev ? __shfl_down_sync(0xffffffffu, m.b, o) : __shfl_up_sync(0xffffffffu, m.b, o);
    if (pos >= o) m = aff_compose(m, Aff{pa, pb});
  }
  if (pos == 31) { sh[2 * wpos] = m.a; sh[2 * wpos + 1] = m.b; }
  group_bar(1, nw * 32);
  if (wid == 0) {
    Aff t = (lane < nw) ? Aff{sh[2 * lane], sh[2 * lane + 1]} : Aff{1.0, 0.0};
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double pa = __shfl_up_sync(0xffffffffu, t.a, o);
      const double pb = __shfl_up_sync(0xffffffffu, t.b, o);
      if (lane >= o) t = aff_compose(t, Aff{pa, pb});
    }
    if (lane < nw) { sh[2 * lane] = t.a; sh[2 * lane + 1] = t.b; }
  }
  group_bar(1, nw * 32);
  double ea = rev ? __shfl_down_sync(0xffffffffu, m.a, 1) : __shfl_up_sync(0xffffffffu, m.a, 1);
  double eb = rev ? __shfl_down_sync(0xffffffffu, m.b, 1) : __shfl_up_sync(0xffffffffu, m.b, 1);
  if (pos == 0) { ea = 1.0; eb = 0.0; }
  double r = eb;
  if (wpos > 0) r = ea * sh[2 * (wpos - 1) + 1] + eb;
  group_bar(1, nw * 32);
  return r;
}

// F32: the seven operator streams of a line (4 off-line diagonals, 3 line factors) come from the fp32 pack
// J.F (rows padded to 16 bytes: no alignment cases) instead of the fp64 arrays, two fp32 lines per ring slot:
//     x(lower), [A0 A1], x(upper), [A5 A6], b, [lf iu], cu                 (b, [lf iu], cu from a zero guess)
// 52 instead of 80 bytes per vertex of the colour (28 instead of 40 from a zero guess) and 7 instead of 10
// streams, so 5 slots hold most of a line.  x, b and every product / sum stay fp64.
template <bool COUPLED, bool F32>
__device__ __forceinline__ bool ring_is_f32(const int a) {
  if (!F32) return false;
  return COUPLED ? (a == 1 || a == 3 || a >= 5) : (a >= 1);
}
// first array of the fp32 pack [A0 A1 A5 A6 lf iu cu] that stream a carries, and how many consecutive ones
template <bool COUPLED>
__device__ __forceinline__ void ring_f32_arrays(const int a, int& first, int& count) {
  if (COUPLED) {
    first = a == 1 ? 0 : (a == 3 ? 2 : (a == 5 ? 4 : 6));
    count = a == 6 ? 1 : 2;
  } else {
    first = a == 1 ? 4 : 6;
    count = a == 2 ? 1 : 2;
  }
}
// fp64 stream a of the F32 ordering -> stream number of the fp64 ordering (ring_base / ring_row)
template <bool COUPLED>
__device__ __forceinline__ int ring_f32_dstream(const int a) {
  if (COUPLED) return a == 0 ? 0 : (a == 2 ? 3 : 6);
  return 0;
}

template <bool COUPLED, bool F32>
__global__ void __launch_bounds__(kRingMaxT, 1)
zebra_ring_kernel(const LineJob J, const Halo H, const int S) {
  constexpr int NARR = F32 ? (COUPLED ? 7 : 3) : (COUPLED ? 10 : 4);
  extern __shared__ __align__(16) double sm[];
  __shared__ double sh[64];
  const int n1 = J.nx + 1;
  const size_t nv = (size_t)n1 * (J.ny + 1);
  const int SW = (n1 + 9) & ~1;          // slot width (doubles)
  double* s_x = sm + (size_t)S * SW;
  unsigned long long* full = reinterpret_cast<unsigned long long*>(s_x + ((n1 + 2) & ~1));
  unsigned long long* empty = full + S;
  const int tid = threadIdx.x;
  const int ncw = (blockDim.x >> 5) - 1;   // consumer warps; the last warp is the producer
  const int TC = ncw * 32;
  if ((int)blockIdx.x >= J.nlines) return;
  const int niter = (J.nlines - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
  const int ntot = niter * NARR;          // streams this CTA consumes

  if (tid == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, (unsigned)ncw); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (tid >= TC) {
    // ---------------- producer warp: lane s owns ring slot s and issues the streams s, s + S, s + 2S, ... in order
    // (one lane per slot: every lane waits on its own pair of barriers only, so no phase can be skipped; the
    // lanes issue in parallel -- a bulk copy stalls the issuing thread for ~0.25 us)
    const int ps = tid - TC;
    if (COUPLED && blockIdx.x == 0 && ps == 0) {   // line 0 of the job reads the ghost row: wait for the neighbour's push
      if (J.wait == 1) halo_wait_below(H);
      if (J.wait == 2) halo_wait_above(H);
      asm volatile("fence.proxy.async;" ::: "memory");
    }
    __syncwarp();
    if (ps >= S) return;
    for (int n = ps; n < ntot; n += S) {
      const int sidx = ps;
      if (n >= S) mbar_wait(empty + sidx, (unsigned)(((n / S) - 1) & 1));   // the consumers have read stream n - S
      const int it2 = n / NARR, a = n - it2 * NARR;
      const int j = J.jstart + ((int)blockIdx.x + it2 * (int)gridDim.x) * J.jstep;
      if (ring_is_f32<COUPLED, F32>(a)) {
        int first, count;
        ring_f32_arrays<COUPLED>(a, first, count);
        const size_t fstride = (size_t)J.n1p * (J.ny + 1);
        const float* src = J.F + (size_t)first * fstride + (size_t)j * J.n1p;
        const unsigned lb = (unsigned)J.n1p * 4u;
        char* dst = reinterpret_cast<char*>(sm + (size_t)sidx * SW);
        mbar_expect_tx(full + sidx, lb * (unsigned)count);
        tma_g2s(dst, src, lb, full + sidx);
        if (count == 2) tma_g2s(dst + lb, src + fstride, lb, full + sidx);
        continue;
      }
      const int ad = F32 ? ring_f32_dstream<COUPLED>(a) : a;
      const double* base = ring_base<COUPLED>(J, nv, ad);
      const int p0 = addr_parity(base);
      // indices relative to the array: the line needs [g0 - 1, g0 + n1 + 1) clipped to the array
      const long long g0 = (long long)ring_row<COUPLED>(J, ad, j) * n1;
      const int sft = (int)((g0 + p0) & 1);
      long long nl = g0 - 1, nh = g0 + n1 + 1;
      if (nl < 0) nl = 0;
      if (nh > (long long)nv) nh = (long long)nv;
      long long t0 = nl - ((nl + p0) & 1), t1 = nh + ((nh + p0) & 1);   // 16-byte aligned ends
      double* slot = sm + (size_t)sidx * SW + (4 + sft) - g0;           // slot[E] = element E of the array
      // (no proxy fence: the consumers only READ the slot through the generic proxy, and their reads are ordered
      // before this point by the empty barrier)
      if (t0 < 0) { t0 += 2; slot[nl] = base[nl]; }
      if (t1 > (long long)nv) { t1 -= 2; slot[nh - 1] = base[nh - 1]; }
      const unsigned bytes = (unsigned)(t1 - t0) * 8u;
      mbar_expect_tx(full + sidx, bytes);
      tma_g2s(slot + t0, base + t0, bytes, full + sidx);
    }
    return;
  }

  // ---------------- consumer warps
  const int lane = tid & 31;
  const int i0 = tid * kE;
  const int ne = max(0, min(kE, n1 - i0));
  for (int it = 0; it < niter; ++it) {
    const int l = (int)blockIdx.x + it * (int)gridDim.x;
    const int j = J.jstart + l * J.jstep;
    const size_t lbase = (size_t)j * n1;
    int n = it * NARR;
    RT(0);
    if (!F32 && J.prefetch && it + 1 < niter) {
      // every stream of this CTA's NEXT line -> L2, one 128-byte line per prefetch instruction, spread over the
      // consumer threads (bulk prefetches through the copy engine cost as much as the copies themselves): HBM
      // streams a line ahead of the ring, the bulk copies then hit L2
      const int jn = J.jstart + (l + (int)gridDim.x) * J.jstep;
#pragma unroll
      for (int a = 0; a < NARR; ++a) {
        const char* p = reinterpret_cast<const char*>(ring_base<COUPLED>(J, nv, a) + (size_t)ring_row<COUPLED>(J, a, jn) * n1);
        for (int c = tid * 128; c < n1 * 8; c += TC * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(p + c));
      }
    }
    // stream `a` of this line: wait, return the pointer to the own first element
    auto acquire = [&](const int a) -> const double* {   // a: stream number of the fp64 ordering
      const int sidx = n % S;
      mbar_wait(full + sidx, (unsigned)((n / S) & 1));
      const int sft = (addr_parity(ring_base<COUPLED>(J, nv, a)) + ((ring_row<COUPLED>(J, a, j) & n1) & 1)) & 1;
      return sm + (size_t)sidx * SW + i0 + 4 + sft;
    };
    auto acquire_f = [&]() -> const float* {   // fp32 stream: own first element of the (first) line in the slot
      const int sidx = n % S;
      mbar_wait(full + sidx, (unsigned)((n / S) & 1));
      return reinterpret_cast<const float*>(sm + (size_t)sidx * SW) + i0;
    };
    auto release = [&]() {
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + (n % S));
      ++n;
    };
    double r[kE], lv[kE], iv[kE], cv[kE];
    if (COUPLED) {
      const bool has_lo = j > 0, has_up = j < J.ny;   // the clamped stand-in rows must not leak NaNs through 0 * x
      double xn[kE + 1];
      {
        const double* p = acquire(0);   // x on the line below: elements i-1 .. i+kE-1
#pragma unroll
        for (int e = 0; e <= kE; ++e) xn[e] = (has_lo && ne > 0 && e <= ne && i0 + e - 1 >= 0) ? p[e - 1] : 0.0;
        release();
      }
      RT(1);
      if (F32) {
        const float* p = acquire_f();
        const float* p2 = p + J.n1p;
#pragma unroll
        for (int e = 0; e < kE; ++e) r[e] = (e < ne) ? -((double)p[e] * xn[e] + (double)p2[e] * xn[e + 1]) : 0.0;
        release();
      } else {
        {
          const double* p = acquire(1);
#pragma unroll
          for (int e = 0; e < kE; ++e) r[e] = (e < ne) ? -p[e] * xn[e] : 0.0;
          release();
        }
        {
          const double* p = acquire(2);
#pragma unroll
          for (int e = 0; e < kE; ++e)
            if (e < ne) r[e] -= p[e] * xn[e + 1];
          release();
        }
      }
      {
        const double* p = acquire(3);   // x on the line above: elements i .. i+kE
#pragma unroll
        for (int e = 0; e <= kE; ++e) xn[e] = (has_up && ne > 0 && e <= ne && i0 + e <= J.nx) ? p[e] : 0.0;
        release();
      }
      if (F32) {
        const float* p = acquire_f();
        const float* p2 = p + J.n1p;
#pragma unroll
        for (int e = 0; e < kE; ++e)
          if (e < ne) r[e] -= (double)p[e] * xn[e] + (double)p2[e] * xn[e + 1];
        release();
      } else {
        {
          const double* p = acquire(4);
#pragma unroll
          for (int e = 0; e < kE; ++e)
            if (e < ne) r[e] -= p[e] * xn[e];
          release();
        }
        {
          const double* p = acquire(5);
#pragma unroll
          for (int e = 0; e < kE; ++e)
            if (e < ne) r[e] -= p[e] * xn[e + 1];
          release();
        }
      }
      {
        const double* p = acquire(6);
#pragma unroll
        for (int e = 0; e < kE; ++e)
          if (e < ne) r[e] += p[e];
        release();
      }
    } else {
      const double* p = acquire(0);
#pragma unroll
      for (int e = 0; e < kE; ++e)
        if (e < ne) r[e] = p[e];
      release();
    }
    RT(2);
    if (F32) {
      {
        const float* p = acquire_f();
        const float* p2 = p + J.n1p;
#pragma unroll
        for (int e = 0; e < kE; ++e)
          if (e < ne) { lv[e] = (double)p[e]; iv[e] = (double)p2[e]; }
        release();
      }
      {
        const float* p = acquire_f();
#pragma unroll
        for (int e = 0; e < kE; ++e)
          if (e < ne) cv[e] = (double)p[e];
        release();
      }
    } else {
      {
        const double* p = acquire(NARR - 3);
#pragma unroll
        for (int e = 0; e < kE; ++e)
          if (e < ne) lv[e] = p[e];
        release();
      }
      {
        const double* p = acquire(NARR - 2);
#pragma unroll
        for (int e = 0; e < kE; ++e)
          if (e < ne) iv[e] = p[e];
        release();
      }
      {
        const double* p = acquire(NARR - 1);
#pragma unroll
        for (int e = 0; e < kE; ++e)
          if (e < ne) cv[e] = p[e];
        release();
      }
    }
    RT(3);
    // forward: z_i = r_i - l_i z_{i-1};  w_i = z_i / u_i
    Aff m{1.0, 0.0};
#pragma unroll
    for (int e = 0; e < kE; ++e)
      if (e < ne) { m.b = r[e] - lv[e] * m.b; m.a = -lv[e] * m.a; }
    double z = aff_consumer_scan(m, sh, false, ncw);
#pragma unroll
    for (int e = 0; e < kE; ++e)
      if (e < ne) { z = r[e] - lv[e] * z; r[e] = z * iv[e]; }
    RT(4);
    // backward: x_i = w_i - (c_i/u_i) x_{i+1}
    m = Aff{1.0, 0.0};
#pragma unroll
    for (int e = kE - 1; e >= 0; --e)
      if (e < ne) { m.b = r[e] - cv[e] * m.b; m.a = -cv[e] * m.a; }
    double xb = aff_consumer_scan(m, sh, true, ncw);
#pragma unroll
    for (int e = kE - 1; e >= 0; --e)
      if (e < ne) { xb = r[e] - cv[e] * xb; s_x[i0 + e] = xb; }
    RT(5);
    group_bar(1, TC);
    RT(6);
    double* xo = J.x + lbase;
    for (int i = tid; i < n1; i += TC) xo[i] = s_x[i];
    RT(7);
    if (l == 0 && J.push && H.active) {   // strip-boundary line: also into the neighbour's copy of x
      double* pbase = J.push == 1 ? H.dn_base : H.up_base;
      if (pbase != nullptr) {
        double* xp = halo_peer_ptr(H, pbase, J.x) + lbase;
        for (int i = tid; i < n1; i += TC) xp[i] = s_x[i];
      }
      __threadfence_system();
      group_bar(1, TC);
      if (tid == 0) {   // publish (cm_halo.cuh: halo_publish, here with the consumers' barrier)
        unsigned long long* c = halo_ctl(H);
        const bool down = J.push == 1;
        const unsigned long long seq = ld_volatile_u64(c + (down ? kChCntDn : kChCntUp)) + 1ull;
        c[down ? kChCntDn : kChCntUp] = seq;
        if (pbase != nullptr)
          st_release_sys(reinterpret_cast<unsigned long long*>(pbase) + (size_t)H.chan * kChStride +
                             (down ? kChFlagAbove : kChFlagBelow), seq);
      }
    }
    group_bar(1, TC);   // s_x is rewritten by the next line
  }
}

// ------------------------------------------------------------------------------------------------
// zebra half sweep on short lines: one CTA per line, direct loads, scans through shared memory
// ------------------------------------------------------------------------------------------------
template <bool XZERO>
__global__ void __launch_bounds__(512, 2)
zebra_small_kernel(const LineJob J, const Halo H, const int E) {
  extern __shared__ double sm[];
  const int n1 = J.nx + 1, nx = J.nx, ny = J.ny;
  double* s_z = sm;
  double* s_c = sm + n1;
  __shared__ double sh[64];
  const int l = blockIdx.x;
  const int j = J.jstart + l * J.jstep;
  const size_t nv = (size_t)n1 * (ny + 1);
  const size_t base = (size_t)j * n1;
  const int T = blockDim.x, tid = threadIdx.x;
  if (!XZERO && l == 0 && H.active) {
    if (tid == 0) {
      if (J.wait == 1) halo_wait_below(H);
      if (J.wait == 2) halo_wait_above(H);
    }
    __syncthreads();
  }
  {
    const size_t lo = (size_t)(j > 0 ? j - 1 : 0) * n1;
    const size_t up = (size_t)(j < ny ? j + 1 : ny) * n1;
    const double* __restrict__ A0 = J.A + base;
    const double* __restrict__ A1 = J.A + nv + base;
    const double* __restrict__ A5 = J.A + 5 * nv + base;
    const double* __restrict__ A6 = J.A + 6 * nv + base;
    const double* xr = J.x;   // not __restrict__: the ghost rows were written by another GPU
    for (int i = tid; i < n1; i += T) {
      double r = J.b[base + i];
      const double lfv = J.lf[base + i];
      if (!XZERO) {
        const int im = i > 0 ? i - 1 : 0, ip = i < nx ? i + 1 : nx;
        if (j > 0) r -= A0[i] * xr[lo + im] + A1[i] * xr[lo + i];
        if (j < ny) r -= A5[i] * xr[up + i] + A6[i] * xr[up + ip];
      }
      s_z[i] = r;
      s_c[i] = lfv;
    }
  }
  __syncthreads();
  {
    const int i0 = tid * E, i1 = min(i0 + E, n1);
    Aff m{1.0, 0.0};
    for (int i = i0; i < i1; ++i) {
      const double lv = s_c[i];
      m.b = s_z[i] - lv * m.b;
      m.a = -lv * m.a;
    }
    double z = aff_block_scan(m, sh, false);
    for (int i = i0; i < i1; ++i) {
      z = s_z[i] - s_c[i] * z;
      s_z[i] = z;
    }
  }
  __syncthreads();
  for (int i = tid; i < n1; i += T) {
    s_z[i] *= J.iu[base + i];
    s_c[i] = J.cu[base + i];
  }
  __syncthreads();
  {
    const int i0 = tid * E, i1 = min(i0 + E, n1);
    Aff m{1.0, 0.0};
    for (int i = i1 - 1; i >= i0; --i) {
      const double c = s_c[i];
      m.b = s_z[i] - c * m.b;
      m.a = -c * m.a;
    }
    double xn = aff_block_scan(m, sh, true);
    for (int i = i1 - 1; i >= i0; --i) {
      xn = s_z[i] - s_c[i] * xn;
      s_z[i] = xn;
    }
  }
  __syncthreads();
  for (int i = tid; i < n1; i += T) J.x[base + i] = s_z[i];
  if (l == 0 && J.push && H.active) {
    double* pbase = J.push == 1 ? H.dn_base : H.up_base;
    if (pbase != nullptr) {
      double* xp = halo_peer_ptr(H, pbase, J.x) + base;
      for (int i = tid; i < n1; i += T) xp[i] = s_z[i];
    }
    halo_publish(H, J.push == 1);
  }
}

// ------------------------------------------------------------------------------------------------
// restricted residual: rc[Y][X] = r[2X] + (r[2X-1] + r[2X+1]) / 2 with r = b - A x on fine row 2Y.
// Valid right after a zebra sweep that ended with colour 1 (the residual of the odd rows is zero, and
// only W/E neighbours of a coarse point lie on even rows).  One CTA per tile of a fine even row.
// wait_below: fine row rb reads the ghost row rb-1 (pushed up by the lower neighbour's last colour-1 sweep).
// ------------------------------------------------------------------------------------------------
constexpr int kRRT = 256;
__global__ void __launch_bounds__(kRRT)
resid_restrict_kernel(const int nxf, const int nyf, const double* __restrict__ A, const double* __restrict__ b,
                      const double* x, double* __restrict__ rc, const int Y0, const int Y1, const Halo H) {
  __shared__ double s_r[kRRT + 2];
  const int n1 = nxf + 1, nxc = nxf >> 1, m1 = nxc + 1;
  const size_t nv = (size_t)n1 * (nyf + 1);
  const int tiles = (m1 + kRRT / 2 - 1) / (kRRT / 2);
  const int Y = Y0 + blockIdx.x / tiles, t = blockIdx.x - (blockIdx.x / tiles) * tiles;
  if (Y >= Y1) return;
  const int iy = 2 * Y;
  if (H.active && Y == Y0) {
    if (threadIdx.x == 0) halo_wait_below(H);
    __syncthreads();
  }
  const int X0 = t * (kRRT / 2);           // first coarse column of the tile
  // fine columns 2*X0 - 1 ... 2*X0 + kRRT - 1  ->  s_r[0 .. kRRT]
  for (int k = threadIdx.x; k <= kRRT; k += kRRT) {
    const int ix = 2 * X0 - 1 + k;
    double s = 0.0;
    if (ix >= 0 && ix <= nxf) {
      const size_t v = (size_t)iy * n1 + ix;
      s = b[v];
#pragma unroll
      for (int q = 0; q < 7; ++q) {
        const int dx = (q == 0 || q == 2) ? -1 : ((q == 4 || q == 6) ? 1 : 0);
        const int dy = (q < 2) ? -1 : (q > 4 ? 1 : 0);
        const int jx = ix + dx, jy = iy + dy;
        if (jx >= 0 && jx <= nxf && jy >= 0 && jy <= nyf) s -= A[(size_t)q * nv + v] * x[(size_t)jy * n1 + jx];
      }
    }
    s_r[k] = s;
  }
  __syncthreads();
  const int X = X0 + threadIdx.x;
  if (threadIdx.x < kRRT / 2 && X <= nxc)
    rc[(size_t)Y * m1 + X] = s_r[2 * threadIdx.x + 1] + 0.5 * (s_r[2 * threadIdx.x] + s_r[2 * threadIdx.x + 2]);
}

// ------------------------------------------------------------------------------------------------
// x += P e on the odd fine rows jb, jb+2, ... < je (jb odd).  CTA 0 handles the LAST odd row (the strip's
// row re-1) alone: it waits for the coarse ghost row above (pushed down by the upper neighbour's last
// colour-0 sweep on the coarse level; wait_coarse), stores the row into the upper neighbour's copy of x and
// publishes an up-push.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
prolong_odd_kernel(const int nxc, const int nyc, const double* ec, double* __restrict__ xf, const int jb,
                   const int je, const Halo H, const int wait_coarse) {
  const int nxf = 2 * nxc, n1 = nxf + 1, m1 = nxc + 1;
  const int nrows = (je - jb + 1) / 2;
  auto corr = [&](int iy, int ix) -> double {
    const int Xc = ix >> 1, Yc = iy >> 1;   // iy odd: coarse rows Yc and Yc + 1
    const size_t c0 = (size_t)Yc * m1 + Xc;
    return 0.5 * (ec[c0] + ((ix & 1) ? ec[c0 + m1 + 1] : ec[c0 + m1]));
  };
  if (blockIdx.x == 0) {
    const int iy = jb + 2 * (nrows - 1);
    if (H.active && wait_coarse) {
      if (threadIdx.x == 0) halo_wait_above(H);
      __syncthreads();
    }
    double* xp = (H.active && H.up_base != nullptr) ? halo_peer_ptr(H, H.up_base, xf) : nullptr;
    for (int ix = threadIdx.x; ix <= nxf; ix += blockDim.x) {
      const size_t v = (size_t)iy * n1 + ix;
      const double val = xf[v] + corr(iy, ix);
      xf[v] = val;
      if (xp != nullptr) xp[v] = val;
    }
    halo_publish(H, false);
    return;
  }
  const long long total = (long long)(nrows - 1) * n1;
  for (long long t = (long long)(blockIdx.x - 1) * blockDim.x + threadIdx.x; t < total;
       t += (long long)(gridDim.x - 1) * blockDim.x) {
    const int rr = (int)(t / n1), ix = (int)(t - (long long)rr * n1);
    const int iy = jb + 2 * rr;
    xf[(size_t)iy * n1 + ix] += corr(iy, ix);
  }
}

// ------------------------------------------------------------------------------------------------
// The V-cycle below a given level in one launch.  NCTA thread blocks (one cluster) share the work of every
// phase; phases are separated by a cluster barrier.  Lines are short (n1 <= 32*kCE): one warp per line.
// ------------------------------------------------------------------------------------------------
constexpr int kCE = 9;   // unknowns per lane: lines of up to 288 unknowns
constexpr int kCoarseT = 512;

template <int NCTA>
__device__ __forceinline__ void phase_sync() {
  if (NCTA == 1) {
    __syncthreads();
  } else {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
}
template <int NCTA>
__device__ __forceinline__ int cta_rank() {
  if (NCTA == 1) return 0;
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return (int)r;
}

__device__ __forceinline__ void warp_line_solve(const CoarseLevel& L, const int j, const bool xzero) {
  const int n1 = L.nx + 1, lane = threadIdx.x & 31;
  const size_t nv = (size_t)n1 * (L.ny + 1);
  const size_t base = (size_t)j * n1;
  const int E = (n1 + 31) >> 5;
  const int i0 = lane * E;
  double r[kCE], lv[kCE], cv[kCE];
  const double* x = L.x;
  const size_t lo = (size_t)(j > 0 ? j - 1 : 0) * n1, up = (size_t)(j < L.ny ? j + 1 : L.ny) * n1;
#pragma unroll
  for (int e = 0; e < kCE; ++e) {
    const int i = i0 + e;
    if (e < E && i < n1) {
      double s = L.b[base + i];
      if (!xzero) {
        const int im = i > 0 ? i - 1 : 0, ip = i < L.nx ? i + 1 : L.nx;
        if (j > 0) s -= __ldg(L.A + base + i) * x[lo + im] + __ldg(L.A + nv + base + i) * x[lo + i];
        if (j < L.ny) s -= __ldg(L.A + 5 * nv + base + i) * x[up + i] + __ldg(L.A + 6 * nv + base + i) * x[up + ip];
      }
      r[e] = s;
      lv[e] = __ldg(L.lf + base + i);
      cv[e] = __ldg(L.cu + base + i);
    }
  }
  Aff m{1.0, 0.0};
#pragma unroll
  for (int e = 0; e < kCE; ++e)
    if (e < E && i0 + e < n1) { m.b = r[e] - lv[e] * m.b; m.a = -lv[e] * m.a; }
  double z = aff_warp_scan(m, false);
#pragma unroll
  for (int e = 0; e < kCE; ++e)
    if (e < E && i0 + e < n1) { z = r[e] - lv[e] * z; r[e] = z * __ldg(L.iu + base + i0 + e); }
  m = Aff{1.0, 0.0};
#pragma unroll
  for (int e = kCE - 1; e >= 0; --e)
    if (e < E && i0 + e < n1) { m.b = r[e] - cv[e] * m.b; m.a = -cv[e] * m.a; }
  double xn = aff_warp_scan(m, true);
#pragma unroll
  for (int e = kCE - 1; e >= 0; --e)
    if (e < E && i0 + e < n1) { xn = r[e] - cv[e] * xn; L.x[base + i0 + e] = xn; }
}

__device__ int g_cc_n;   // phase counter of the coarse-cycle stamps (diagnostics)
#define CCT()                                                                    \
  do {                                                                           \
    if (C.dbg && cta_rank<NCTA>() == 0 && threadIdx.x == 0 && g_cc_n < 64 * 16 * 8) \
      g_ring_dbg[g_cc_n++] = global_ns();                                        \
  } while (0)

template <int NCTA>
__device__ __forceinline__ void coarse_half_sweep(const CoarseLevel& L, const int colour, const bool xzero) {
  const int nw = (blockDim.x >> 5) * NCTA;
  const int w = cta_rank<NCTA>() * (blockDim.x >> 5) + (threadIdx.x >> 5);
  for (int j = colour + 2 * w; j <= L.ny; j += 2 * nw) warp_line_solve(L, j, xzero);
  phase_sync<NCTA>();
}

template <int NCTA>
__global__ void __launch_bounds__(kCoarseT, 1)
coarse_cycle_kernel(const CoarseCycle C) {
  const int gt = cta_rank<NCTA>() * blockDim.x + threadIdx.x, gn = NCTA * blockDim.x;
  const int last = C.nlev - 1;
  if (C.dbg && cta_rank<NCTA>() == 0 && threadIdx.x == 0) g_cc_n = 0;
  CCT();
  // downwards: smooth, restricted residual
  for (int k = 0; k < last; ++k) {
    const CoarseLevel& L = C.lev[k];
    const CoarseLevel& N = C.lev[k + 1];
    for (int s = 0; s < C.pre; ++s) {
      coarse_half_sweep<NCTA>(L, 0, s == 0);
      CCT();
      coarse_half_sweep<NCTA>(L, 1, false);
      CCT();
    }
    {
      const int n1 = L.nx + 1, m1 = N.nx + 1;
      const size_t nv = (size_t)n1 * (L.ny + 1);
      const int ncoarse = m1 * (N.ny + 1);
      const double* x = L.x;
      for (int I = gt; I < ncoarse; I += gn) {
        const int X = I % m1, Y = I / m1;
        const int iy = 2 * Y;
        double acc = 0.0;
#pragma unroll
        for (int d = -1; d <= 1; ++d) {
          const int ix = 2 * X + d;
          if (ix < 0 || ix > L.nx) continue;
          const size_t v = (size_t)iy * n1 + ix;
          double s = L.b[v];
#pragma unroll
          for (int q = 0; q < 7; ++q) {
            const int dx = (q == 0 || q == 2) ? -1 : ((q == 4 || q == 6) ? 1 : 0);
            const int dy = (q < 2) ? -1 : (q > 4 ? 1 : 0);
            const int jx = ix + dx, jy = iy + dy;
            if (jx >= 0 && jx <= L.nx && jy >= 0 && jy <= L.ny) s -= __ldg(L.A + (size_t)q * nv + v) * x[(size_t)jy * n1 + jx];
          }
          acc += (d == 0) ? s : 0.5 * s;
        }
        N.b[I] = acc;
      }
    }
    phase_sync<NCTA>();
    CCT();
  }
  // coarsest level
  {
    const CoarseLevel& L = C.lev[last];
    const int n = (L.nx + 1) * (L.ny + 1);
    if (C.Ainv != nullptr) {
      for (int r = gt; r < n; r += gn) {
        double s = 0.0;
        for (int c = 0; c < n; ++c) s += __ldg(C.Ainv + (size_t)r * n + c) * L.b[c];
        L.x[r] = s;
      }
      phase_sync<NCTA>();
    } else {
      for (int s = 0; s < 30; ++s) {
        coarse_half_sweep<NCTA>(L, 0, s == 0);
        coarse_half_sweep<NCTA>(L, 1, false);
      }
    }
    CCT();
  }
  // upwards: correction on the odd rows, smooth
  for (int k = last - 1; k >= 0; --k) {
    const CoarseLevel& L = C.lev[k];
    const CoarseLevel& N = C.lev[k + 1];
    const int n1 = L.nx + 1, m1 = N.nx + 1;
    const int nodd = L.ny / 2;   // odd rows 1, 3, ..., < ny (ny even on every level that has a coarser one)
    const double* ec = N.x;
    for (int t = gt; t < nodd * n1; t += gn) {
      const int rr = t / n1, ix = t - rr * n1;
      const int iy = 2 * rr + 1;
      const size_t c0 = (size_t)(iy >> 1) * m1 + (ix >> 1);
      L.x[(size_t)iy * n1 + ix] += 0.5 * (ec[c0] + ((ix & 1) ? ec[c0 + m1 + 1] : ec[c0 + m1]));
    }
    phase_sync<NCTA>();
    CCT();
    for (int s = 0; s < C.post; ++s) {
      coarse_half_sweep<NCTA>(L, 0, false);
      CCT();
      coarse_half_sweep<NCTA>(L, 1, false);
      CCT();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
// opt-in shared-memory sizes, once per process (idempotent; called at context creation so that no attribute
// call lands inside a stream capture)
int mg_init_attributes() {
  static std::atomic<bool> done{false};
  if (done.load()) return CM_OK;
  CM_CUDA(cudaFuncSetAttribute(zebra_ring_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  CM_CUDA(cudaFuncSetAttribute(zebra_ring_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  CM_CUDA(cudaFuncSetAttribute(zebra_ring_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  CM_CUDA(cudaFuncSetAttribute(zebra_ring_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
  CM_CUDA(cudaFuncSetAttribute(zebra_small_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CM_CUDA(cudaFuncSetAttribute(zebra_small_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  done.store(true);
  return CM_OK;
}

static bool ring_fits(int n1) { return n1 >= 1024 && ((n1 + kE - 1) / kE + 31) / 32 * 32 + 32 <= kRingMaxT; }
bool mg_ring_fits(int nx) { return ring_fits(nx + 1); }

// ------------------------------------------------------------------------------------------------
// fp32 pack of the operator streams of the long-line kernel: [A0 A1 A5 A6 lf iu cu][row][n1p], n1p = n1 rounded
// up to 4 floats (every row starts on a 16-byte boundary, the pad is zero).  Written once per Jacobian.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pack_f32_kernel(const int n1, const int n1p, const size_t nv, const int r0, const int r1, const double* __restrict__ A,
                const double* __restrict__ lf, const double* __restrict__ iu, const double* __restrict__ cu,
                float* __restrict__ F) {
  const size_t fstride = (size_t)n1p * (nv / n1);
  const long long total = (long long)(r1 - r0) * n1p;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
    const int row = r0 + (int)(t / n1p), i = (int)(t - (long long)(row - r0) * n1p);
    const size_t o = (size_t)row * n1p + i;
    const bool on = i < n1;
    const size_t v = (size_t)row * n1 + (on ? i : 0);
    F[o] = on ? (float)A[v] : 0.0f;
    F[fstride + o] = on ? (float)A[nv + v] : 0.0f;
    F[2 * fstride + o] = on ? (float)A[5 * nv + v] : 0.0f;
    F[3 * fstride + o] = on ? (float)A[6 * nv + v] : 0.0f;
    F[4 * fstride + o] = on ? (float)lf[v] : 0.0f;
    F[5 * fstride + o] = on ? (float)iu[v] : 0.0f;
    F[6 * fstride + o] = on ? (float)cu[v] : 0.0f;
  }
}

size_t mg_pack_f32_doubles(int nx, int ny) {   // doubles of workspace the pack of one level takes
  const size_t n1p = (size_t)((nx + 1 + 3) & ~3);
  return (7 * n1p * (size_t)(ny + 1) + 1) / 2 + 2;
}

// rows [r0, r1) (r1 < 0: all)
int mg_pack_f32(int nx, int ny, const double* A, const double* lf, const double* iu, const double* cu, float* F,
                int r0, int r1, cudaStream_t st) {
  if (r1 < 0) { r0 = 0; r1 = ny + 1; }
  if (r1 <= r0) return CM_OK;
  const int n1 = nx + 1, n1p = (n1 + 3) & ~3;
  const size_t nv = (size_t)n1 * (ny + 1);
  long long g = ((long long)(r1 - r0) * n1p + 255) / 256;
  if (g > 148 * 16) g = 148 * 16;
  pack_f32_kernel<<<(int)g, 256, 0, st>>>(n1, n1p, nv, r0, r1, A, lf, iu, cu, F);
  CM_BYTES((double)(r1 - r0) * n1 * (7 * 8.0 + 7 * 4.0));
  CM_LAUNCH_CHECK();
  return CM_OK;
}

// shared memory of one CTA: S line-sized slots + the output line + 2 S barriers
static size_t ring_smem(int n1, int S) {
  const int SW = (n1 + 9) & ~1;
  return ((size_t)S * SW + ((n1 + 2) & ~1) + 2 * S + 2) * sizeof(double);
}

// Ring geometry: from a zero guess the four streams of a line fit (the next line loads during the scans);
// with couplings 5 of the 10 streams are in flight in shared memory and the rest waits in L2.
static void ring_config(int n1, int narr, int& T, int& S, size_t& smem, int& occ) {
  const int nw = ((n1 + kE - 1) / kE + 31) / 32;
  T = (nw + 1) * 32;   // + the producer warp
  S = narr <= 4 ? narr : 5;   // measured (tools/time_ring.py): 5 slots beat a whole line in flight on 2049-point lines
  if (const char* e = getenv("CM_RING_S")) { const int v = atoi(e); if (v >= 2 && v <= 16) S = v; }
  while (S > 2 && ring_smem(n1, S) > 220 * 1024) --S;
  smem = ring_smem(n1, S);
  occ = (int)((228 * 1024) / (smem + 1024));
  if (occ < 1) occ = 1;
  if (occ > 4) occ = 4;
  if (occ * T > 2048) occ = 2048 / T;
}

// lines of `colour` among the rows [r0, r1) of a grid with ny+1 rows; r1 < 0: whole grid.  Multi-GPU strips
// (H.active): r0 is even; colour 0 starts at r0 (waits below, pushes down), colour 1 at the last odd row
// (waits above, pushes up).
int mg_zebra(int nx, int ny, int colour, const double* A, const double* lf, const double* iu, const double* cu,
             const double* b, double* x, int x_zero, const Halo& H, int r0, int r1, cudaStream_t st, bool prof,
             const float* F) {
  if (r1 < 0) { r0 = 0; r1 = ny + 1; }
  const int n1 = nx + 1;
  LineJob J{};
  J.nx = nx; J.ny = ny; J.A = A; J.lf = lf; J.iu = iu; J.cu = cu; J.b = b; J.x = x;
  if (colour == 0) {
    J.jstart = r0 + (r0 & 1);
    J.jstep = 2;
    J.nlines = (r1 - J.jstart + 1) / 2;
    J.wait = 1; J.push = 1;
  } else {
    const int jlast = ((r1 - 1) & 1) ? r1 - 1 : r1 - 2;   // last odd row < r1
    const int jfirst = r0 + ((r0 & 1) ? 0 : 1);
    J.jstart = jlast;
    J.jstep = -2;
    J.nlines = jlast >= jfirst ? (jlast - jfirst) / 2 + 1 : 0;
    J.wait = 2; J.push = 2;
  }
  if (!H.active) { J.wait = 0; J.push = 0; }
  if (J.nlines <= 0) {
    if (H.active) { set_error("mg_zebra: a strip without lines of colour %d", colour); return CM_E_INVALID; }
    return CM_OK;
  }
  int rc_attr;
  bool use_ring = ring_fits(n1);
  const bool f32 = use_ring && F != nullptr;
  const double bytes = (double)J.nlines * n1 * (f32 ? (x_zero ? 28.0 : 52.0) : (x_zero ? 40.0 : 80.0));
  CM_BYTES(bytes);
  if (prof) prof_begin(kProfZebraFine, bytes, st);
  if (const char* e = getenv("CM_ZEBRA_SMALL_BELOW")) {   // tuning: one CTA per line when the launch has few lines
    if (J.nlines <= atoi(e)) use_ring = false;
  }
  if (use_ring) {
    int T, S, occ;
    size_t smem;
    ring_config(n1, f32 ? (x_zero ? 3 : 7) : (x_zero ? 4 : 10), T, S, smem, occ);
    J.F = f32 ? F : nullptr;
    J.n1p = (n1 + 3) & ~3;
    if (const char* e = getenv("CM_RING_DBG")) J.dbg = atoi(e);
    J.prefetch = 0;   // measured: pulling the next line into L2 ahead of the ring costs more than it hides (profiles/README.md)
    if (const char* e = getenv("CM_RING_PREFETCH")) J.prefetch = atoi(e);
    if (smem > 220 * 1024) {
      set_error("zebra ring: %zu bytes of shared memory requested (CM_RING_S too large)", smem);
      return CM_E_INVALID;
    }
    if ((rc_attr = mg_init_attributes())) return rc_attr;
    const int grid = J.nlines < 148 * occ ? J.nlines : 148 * occ;
    if (f32) {
      if (x_zero) zebra_ring_kernel<false, true><<<grid, T, smem, st>>>(J, H, S);
      else zebra_ring_kernel<true, true><<<grid, T, smem, st>>>(J, H, S);
    } else if (x_zero) {
      zebra_ring_kernel<false, false><<<grid, T, smem, st>>>(J, H, S);
    } else {
      zebra_ring_kernel<true, false><<<grid, T, smem, st>>>(J, H, S);
    }
  } else {
    const int T = (n1 > 2048) ? 512 : ((n1 > 1024) ? 256 : ((n1 > 128) ? 128 : 32));
    int E = (n1 + T - 1) / T;
    if ((E & 1) == 0) ++E;
    const size_t smem = 2 * (size_t)n1 * sizeof(double);
    if (smem > 200 * 1024) {
      set_error("zebra line relaxation: line of %d unknowns exceeds shared memory", n1);
      return CM_E_UNSUPPORTED;
    }
    if ((rc_attr = mg_init_attributes())) return rc_attr;
    if (x_zero) zebra_small_kernel<true><<<J.nlines, T, smem, st>>>(J, H, E);
    else zebra_small_kernel<false><<<J.nlines, T, smem, st>>>(J, H, E);
  }
  if (prof) prof_end(kProfZebraFine, st);
  CM_LAUNCH_CHECK();
  return CM_OK;
}

// coarse rows [Y0, Y1) (Y1 < 0: all)
int mg_resid_restrict(int nxf, int nyf, const double* A, const double* b, const double* x, double* rc,
                      const Halo& H, int Y0, int Y1, cudaStream_t st) {
  const int nyc = nyf >> 1, m1 = (nxf >> 1) + 1;
  if (Y1 < 0) { Y0 = 0; Y1 = nyc + 1; }
  if (Y1 <= Y0) return CM_OK;
  const int tiles = (m1 + kRRT / 2 - 1) / (kRRT / 2);
  resid_restrict_kernel<<<(Y1 - Y0) * tiles, kRRT, 0, st>>>(nxf, nyf, A, b, x, rc, Y0, Y1, H);
  CM_BYTES((double)(Y1 - Y0) * (nxf + 1) * (8 * 8.0 + 16.0) + (double)(Y1 - Y0) * m1 * 8.0);
  CM_LAUNCH_CHECK();
  return CM_OK;
}

// fine rows [r0, r1) (r1 < 0: all): correction on the odd rows only
int mg_prolong_odd(int nxc, int nyc, const double* ec, double* xf, const Halo& H, int wait_coarse, int r0,
                   int r1, cudaStream_t st) {
  const int nyf = 2 * nyc, n1 = 2 * nxc + 1;
  if (r1 < 0) { r0 = 0; r1 = nyf + 1; }
  const int jb = r0 + ((r0 & 1) ? 0 : 1);
  const int je = r1;
  const int nrows = je > jb ? (je - jb + 1) / 2 : 0;
  if (nrows <= 0) {
    if (H.active) { set_error("mg_prolong_odd: a strip without odd rows"); return CM_E_INVALID; }
    return CM_OK;
  }
  long long g = ((long long)(nrows - 1) * n1 + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  prolong_odd_kernel<<<(int)g + 1, 256, 0, st>>>(nxc, nyc, ec, xf, jb, je, H, wait_coarse);
  CM_BYTES((double)nrows * n1 * (16.0 + 4.0));
  CM_LAUNCH_CHECK();
  return CM_OK;
}

int mg_coarse_max_n1() { return 32 * kCE; }

int mg_ring_debug_times(unsigned long long* host_out, int n) {
  return cudaMemcpyFromSymbol(host_out, g_ring_dbg, sizeof(unsigned long long) * n) == cudaSuccess ? CM_OK : CM_E_CUDA;
}

int mg_coarse_cycle(const CoarseCycle& C0, int ncta, cudaStream_t st) {
  CoarseCycle C = C0;
  if (const char* e = getenv("CM_COARSE_DBG")) C.dbg = atoi(e);
  for (int k = 0; k < C.nlev; ++k)
    if (C.lev[k].nx + 1 > 32 * kCE) {
      set_error("mg_coarse_cycle: lines of %d unknowns are too long for the fused cycle", C.lev[k].nx + 1);
      return CM_E_UNSUPPORTED;
    }
  double bytes = 0.0;
  for (int k = 0; k < C.nlev; ++k)
    bytes += (double)(C.lev[k].nx + 1) * (C.lev[k].ny + 1) * 8.0 * 13;   // every array of the level once
  CM_BYTES(bytes);
  if (ncta <= 1) {
    coarse_cycle_kernel<1><<<1, kCoarseT, 0, st>>>(C);
  } else {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(8, 1, 1);
    cfg.blockDim = dim3(kCoarseT, 1, 1);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 8;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CM_CUDA(cudaLaunchKernelEx(&cfg, coarse_cycle_kernel<8>, C));
  }
  CM_LAUNCH_CHECK();
  return CM_OK;
}

}  // namespace cm
