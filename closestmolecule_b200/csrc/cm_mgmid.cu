// Zebra half sweeps for the CACHE-RESIDENT middle levels of the multigrid hierarchy (lines of 65 .. 1280
// unknowns: from 1025 x 513 down to where the shared-memory resident coarse cycle takes over).  Every array of
// these levels sits in the 126 MB L2, so a phase costs latency, not bytes: the kernels here are built around
// the dependent chain  "one L2 round trip -> two scans -> store"  and around having FEWER phases per level:
//   mode 0  colour-c half sweep from a zero guess
//   mode 1  colour-c half sweep with the couplings to the neighbouring lines
//   mode 2  colour-1 half sweep FUSED with the restricted residual: the CTA of even row 2Y solves the odd
//           lines 2Y-1 and 2Y+1 (two thread groups side by side; every odd line is solved by both of its
//           even neighbours, bit-identically -- the duplicate costs L2 traffic, not latency), stores line
//           2Y+1, evaluates r = b - A x on row 2Y from the two fresh lines in shared memory and writes the
//           full-weighting restriction into the coarse right-hand side: the separate residual / restriction
//           launch of the down leg disappears
//   mode 3  colour-0 half sweep FUSED with the prolongation: the coarse correction is only ever read by the
//           first colour-0 sweep of the up leg (through the odd rows next to the line) -- the odd rows are
//           overwritten by the colour-1 sweep that follows without being read -- so x_odd + P e is formed on
//           the fly in the right-hand side and never stored: the prolongation launch disappears
// With V(1,2): 6 dependent launches per level instead of 8.  All input streams of a line are requested in one
// batch of independent coalesced loads (a single exposed L2 latency), moved into the blocked layout of the scans
// through shared memory (stride 5: conflict free), and the line is solved by two first-order recurrences
// evaluated as scans over affine maps, exactly like the long-line kernel (cm_mg.cu).
// Replaces the smoother / transfer part of the sparse LU the reference requests per Newton step
// (advection_diffusion_convergence.py:212, ...MixedPrimal_convergence_SUPG.py:137).
#include "cm_common.cuh"
#include "cm_decl.cuh"
#include "cm_mg.cuh"

namespace cm {

namespace {

constexpr int kME = 5;        // unknowns per thread in the blocked layout of the scans (odd: conflict free)
constexpr int kMidMaxTL = 256;   // threads per line
constexpr int kMidMaxG = 4;      // lines per CTA

struct Aff2 {
  double a, b;
};
__device__ __forceinline__ Aff2 aff2_compose(const Aff2 later, const Aff2 earlier) {
  return Aff2{later.a * earlier.a, later.a * earlier.b + later.b};
}
__device__ __forceinline__ void mid_group_bar(const int id, const int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// exclusive scan over the affine maps of the TL = 32 * wpl threads of one line group, in thread order (rev: in
// reverse thread order); returns the value entering this thread's segment when 0 enters the first one.
// t: thread index inside the group; sh: 2 * 8 doubles of the group; bar: named barrier of the group
__device__ __forceinline__ double mid_group_scan(Aff2 m, double* sh, const bool rev, const int t, const int wpl,
                                                 const int bar) {
  const int lane = t & 31, wid = t >> 5;
  const int pos = rev ? 31 - lane : lane;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double pa = rev ? __shfl_down_sync(0xffffffffu, m.a, o) : __shfl_up_sync(0xffffffffu, m.a, o);
    const double pb = rev ? __shfl_down_sync(0xffffffffu, m.b, o) : __shfl_up_sync(0xffffffffu, m.b, o);
    if (pos >= o) m = aff2_compose(m, Aff2{pa, pb});
  }
  double ea = rev ? __shfl_down_sync(0xffffffffu, m.a, 1) : __shfl_up_sync(0xffffffffu, m.a, 1);
  double eb = rev ? __shfl_down_sync(0xffffffffu, m.b, 1) : __shfl_up_sync(0xffffffffu, m.b, 1);
  if (pos == 0) { ea = 1.0; eb = 0.0; }
  if (wpl == 1) return eb;
  const int wpos = rev ? wpl - 1 - wid : wid;
  if (pos == 31) { sh[2 * wpos] = m.a; sh[2 * wpos + 1] = m.b; }
  mid_group_bar(bar, wpl * 32);
  // every warp folds the (at most 7) warp totals before its own: no second barrier-separated pass
  double in = 0.0;
  for (int k = 0; k < wpos; ++k) in = sh[2 * k] * in + sh[2 * k + 1];
  mid_group_bar(bar, wpl * 32);   // sh is reused by the next scan
  return ea * in + eb;
}

// solve the line whose right-hand side and factors sit in shared memory (blocked reads at stride kME); the
// solution replaces the right-hand side.  Each thread touches its own kME entries only.
__device__ __forceinline__ void mid_line_solve(double* s_r, const double* s_l, const double* s_i, const double* s_c,
                                               const int n1, const int t, const int wpl, double* sh, const int bar) {
  const int i0 = t * kME;
  const int ne = max(0, min(kME, n1 - i0));
  double r[kME], lv[kME], cv[kME], iv[kME];
#pragma unroll
  for (int e = 0; e < kME; ++e) {
    const bool on = e < ne;
    r[e] = on ? s_r[i0 + e] : 0.0;
    lv[e] = on ? s_l[i0 + e] : 0.0;
    iv[e] = on ? s_i[i0 + e] : 0.0;
    cv[e] = on ? s_c[i0 + e] : 0.0;
  }
  // forward: z_i = r_i - l_i z_{i-1};  w_i = z_i / u_i
  Aff2 m{1.0, 0.0};
#pragma unroll
  for (int e = 0; e < kME; ++e)
    if (e < ne) { m.b = r[e] - lv[e] * m.b; m.a = -lv[e] * m.a; }
  double z = mid_group_scan(m, sh, false, t, wpl, bar);
#pragma unroll
  for (int e = 0; e < kME; ++e)
    if (e < ne) { z = r[e] - lv[e] * z; r[e] = z * iv[e]; }
  // backward: x_i = w_i - (c_i / u_i) x_{i+1}
  m = Aff2{1.0, 0.0};
#pragma unroll
  for (int e = kME - 1; e >= 0; --e)
    if (e < ne) { m.b = r[e] - cv[e] * m.b; m.a = -cv[e] * m.a; }
  double xb = mid_group_scan(m, sh, true, t, wpl, bar);
#pragma unroll
  for (int e = kME - 1; e >= 0; --e)
    if (e < ne) { xb = r[e] - cv[e] * xb; s_r[i0 + e] = xb; }
}

// x_odd + P e at fine column i of the odd fine row between the coarse rows Yc and Yc + 1 (7-point stencil of the
// right-diagonal meshes: the diagonal neighbour of an odd column is (X + 1, Yc + 1))
__device__ __forceinline__ double mid_corr(const double* __restrict__ ec, const int m1, const int Yc, const int i) {
  const size_t c0 = (size_t)Yc * m1 + (i >> 1);
  return 0.5 * (ec[c0] + ec[c0 + m1 + (i & 1)]);
}

template <int MODE>
__global__ void __launch_bounds__(kMidMaxTL, MODE == 2 ? 1 : 2)
zebra_mid_kernel(const MidJob J) {
  extern __shared__ __align__(16) double sm[];
  __shared__ double sh_all[kMidMaxG][16];
  const int n1 = J.nx + 1, nx = J.nx, ny = J.ny;
  const int TL = J.tl, wpl = TL >> 5;
  const int g = threadIdx.x / TL, t = threadIdx.x - g * TL;
  const int SW = (n1 + 3) & ~1;
  double* s_r = sm + (size_t)g * 4 * SW;
  double* s_l = s_r + SW;
  double* s_i = s_l + SW;
  double* s_c = s_i + SW;
  const size_t nv = (size_t)n1 * (ny + 1);
  int j;
  bool active;
  if (MODE == 2) {
    j = 2 * (int)blockIdx.x - 1 + 2 * g;   // group 0: the odd line below even row 2Y, group 1: the one above
    active = j >= 1 && j <= ny - 1;
  } else {
    const int l = (int)blockIdx.x * J.groups + g;
    j = J.colour + 2 * l;
    active = l < J.nlines;
  }
  constexpr int KS = kME;   // coalesced slots per thread: TL * kME >= n1
  // ---- even row of the fused residual: request its streams first (they do not depend on the solves)
  double eb[3], eA[7][3], exm[3][3];
  const int T2 = 2 * TL;
  const int iy = 2 * (int)blockIdx.x;
  if (MODE == 2) {
    const size_t rbase = (size_t)iy * n1;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int i = (int)threadIdx.x + k * T2;
      const bool on = i < n1;
      eb[k] = on ? J.b[rbase + i] : 0.0;
#pragma unroll
      for (int q = 0; q < 7; ++q) eA[q][k] = on ? J.A[(size_t)q * nv + rbase + i] : 0.0;
      exm[0][k] = (on && i > 0) ? J.x[rbase + i - 1] : 0.0;
      exm[1][k] = on ? J.x[rbase + i] : 0.0;
      exm[2][k] = (on && i < nx) ? J.x[rbase + i + 1] : 0.0;
    }
  }
  if (active) {
    const size_t base = (size_t)j * n1;
    const size_t lo = (size_t)(j > 0 ? j - 1 : 0) * n1;
    const size_t up = (size_t)(j < ny ? j + 1 : ny) * n1;
    const double* __restrict__ A0 = J.A + base;
    const double* __restrict__ A1 = J.A + nv + base;
    const double* __restrict__ A5 = J.A + 5 * nv + base;
    const double* __restrict__ A6 = J.A + 6 * nv + base;
    const double* xr = J.x;
    const int m1 = (nx >> 1) + 1;
    double rb[KS], rl[KS], ri[KS], rc[KS];
    double a0[KS], a1[KS], a5[KS], a6[KS], xl0[KS], xl1[KS], xu0[KS], xu1[KS];
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      const int i = t + k * TL;
      const bool on = i < n1;
      rb[k] = on ? J.b[base + i] : 0.0;
      rl[k] = on ? J.lf[base + i] : 0.0;
      ri[k] = on ? J.iu[base + i] : 0.0;
      rc[k] = on ? J.cu[base + i] : 0.0;
      if (MODE != 0) {
        const int im = i > 0 ? i - 1 : 0, ip = i < nx ? i + 1 : nx;
        const bool dn = on && j > 0, upn = on && j < ny;
        a0[k] = dn ? A0[i] : 0.0;
        a1[k] = dn ? A1[i] : 0.0;
        a5[k] = upn ? A5[i] : 0.0;
        a6[k] = upn ? A6[i] : 0.0;
        xl0[k] = dn ? xr[lo + im] : 0.0;
        xl1[k] = dn ? xr[lo + i] : 0.0;
        xu0[k] = upn ? xr[up + i] : 0.0;
        xu1[k] = upn ? xr[up + ip] : 0.0;
        if (MODE == 3) {   // the odd neighbour rows carry the coarse-grid correction
          if (dn) { xl0[k] += mid_corr(J.ec, m1, (j >> 1) - 1, im); xl1[k] += mid_corr(J.ec, m1, (j >> 1) - 1, i); }
          if (upn) { xu0[k] += mid_corr(J.ec, m1, j >> 1, i); xu1[k] += mid_corr(J.ec, m1, j >> 1, ip); }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < KS; ++k) {
      const int i = t + k * TL;
      if (i < n1) {
        double r = rb[k];
        if (MODE != 0) {
          r -= a0[k] * xl0[k] + a1[k] * xl1[k];
          r -= a5[k] * xu0[k] + a6[k] * xu1[k];
        }
        s_r[i] = r;
        s_l[i] = rl[k];
        s_i[i] = ri[k];
        s_c[i] = rc[k];
      }
    }
    if (wpl == 1) __syncwarp(); else mid_group_bar(1 + g, TL);
    mid_line_solve(s_r, s_l, s_i, s_c, n1, t, wpl, sh_all[g], 1 + g);
    if (MODE != 2) {
      if (wpl == 1) __syncwarp(); else mid_group_bar(1 + g, TL);
      double* xo = J.x + base;
      for (int i = t; i < n1; i += TL) xo[i] = s_r[i];
    }
  }
  if (MODE == 2) {
    __syncthreads();
    // line 2Y+1 (group 1) is stored by this CTA; line 2Y-1 was stored by the CTA below
    const double* s_lo = sm;                       // solution of line 2Y-1 (valid when iy > 0)
    const double* s_up = sm + (size_t)4 * SW;      // solution of line 2Y+1 (valid when iy < ny)
    double* s_res = sm + (size_t)8 * SW;           // residual of row 2Y, entry i + 1; zero at both ends
    if (iy < ny) {
      double* xo = J.x + (size_t)(iy + 1) * n1;
      for (int i = (int)threadIdx.x; i < n1; i += T2) xo[i] = s_up[i];
    }
    const bool has_lo = iy > 0, has_up = iy < ny;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int i = (int)threadIdx.x + k * T2;
      if (i < n1) {
        double s = eb[k];
        // same order of the seven terms as resid_restrict_kernel (cm_mg.cu)
        if (has_lo) {
          if (i > 0) s -= eA[0][k] * s_lo[i - 1];
          s -= eA[1][k] * s_lo[i];
        }
        if (i > 0) s -= eA[2][k] * exm[0][k];
        s -= eA[3][k] * exm[1][k];
        if (i < nx) s -= eA[4][k] * exm[2][k];
        if (has_up) {
          s -= eA[5][k] * s_up[i];
          if (i < nx) s -= eA[6][k] * s_up[i + 1];
        }
        s_res[i + 1] = s;
      }
    }
    if (threadIdx.x == 0) { s_res[0] = 0.0; s_res[n1 + 1] = 0.0; }
    __syncthreads();
    const int m1 = (nx >> 1) + 1;
    double* out = J.bc + (size_t)blockIdx.x * m1;
    for (int X = (int)threadIdx.x; X < m1; X += T2) out[X] = s_res[2 * X + 1] + 0.5 * (s_res[2 * X] + s_res[2 * X + 2]);
  }
}

}  // namespace

bool mg_mid_fits(int nx, int ny) {
  return nx + 1 <= kMidMaxTL * kME && nx + 1 >= 9 && ny >= 2;
}

// the fused sweep + restricted residual runs two line groups per CTA: lines of up to 128 * kME unknowns
bool mg_mid_fused_restrict_fits(int nx, int ny) {
  return mg_mid_fits(nx, ny) && nx + 1 <= (kMidMaxTL / 2) * kME && !(nx & 1) && !(ny & 1);
}

static int mid_tl(int n1) {
  int tl = 32;
  while (tl * kME < n1) tl <<= 1;
  return tl;
}

static size_t mid_smem(int n1, int groups, bool fused) {
  const int SW = (n1 + 3) & ~1;
  return ((size_t)groups * 4 * SW + (fused ? (size_t)n1 + 4 : 0)) * sizeof(double);
}

int mg_mid_init_attributes() {
  static std::atomic<bool> done{false};
  if (done.load()) return CM_OK;
  const int big = 100 * 1024;
  CM_CUDA(cudaFuncSetAttribute(zebra_mid_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CM_CUDA(cudaFuncSetAttribute(zebra_mid_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CM_CUDA(cudaFuncSetAttribute(zebra_mid_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  CM_CUDA(cudaFuncSetAttribute(zebra_mid_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  done.store(true);
  return CM_OK;
}

// mode 0 / 1: half sweep of `colour`; mode 2: colour-1 half sweep + restricted residual -> bc (the coarse grid is
// (nx/2) x (ny/2), ny even); mode 3: colour-0 half sweep reading x_odd + P ec.  Whole grid, single GPU.
int mg_zebra_mid(int mode, int nx, int ny, int colour, const double* A, const double* lf, const double* iu,
                 const double* cu, const double* b, double* x, const double* ec, double* bc, cudaStream_t st) {
  const int n1 = nx + 1;
  if (!mg_mid_fits(nx, ny) || mode < 0 || mode > 3) {
    set_error("mg_zebra_mid: lines of %d unknowns / mode %d are not supported", n1, mode);
    return CM_E_UNSUPPORTED;
  }
  if ((mode == 2 || mode == 3) && ((nx & 1) || (ny & 1))) {
    set_error("mg_zebra_mid: the fused transfer modes need even nx, ny");
    return CM_E_INVALID;
  }
  if (mode == 2 && !mg_mid_fused_restrict_fits(nx, ny)) {
    set_error("mg_zebra_mid: lines of %d unknowns are too long for the fused restricted residual", n1);
    return CM_E_UNSUPPORTED;
  }
  int rc = mg_mid_init_attributes();
  if (rc) return rc;
  MidJob J{};
  J.nx = nx; J.ny = ny; J.A = A; J.lf = lf; J.iu = iu; J.cu = cu; J.b = b; J.x = x; J.ec = ec; J.bc = bc;
  J.tl = mid_tl(n1);
  if (mode == 2) {
    J.colour = 1;
    J.groups = 2;
    J.nlines = ny / 2;
    const int grid = ny / 2 + 1;   // one CTA per even row
    zebra_mid_kernel<2><<<grid, 2 * J.tl, mid_smem(n1, 2, true), st>>>(J);
    // algorithmic bytes: every odd line ONCE (the duplicate solve of each odd line is L2 traffic the fusion pays,
    // not compulsory traffic), the even rows once, the coarse right-hand side
    CM_BYTES((double)(ny / 2) * n1 * 80.0 + (double)(ny / 2 + 1) * n1 * 80.0 + (double)grid * ((nx >> 1) + 1) * 8.0);
  } else {
    J.colour = mode == 3 ? 0 : colour;
    J.nlines = (ny + 1 - J.colour + 1) / 2;
    if (J.nlines <= 0) return CM_OK;
    J.groups = 128 / J.tl > 0 ? 128 / J.tl : 1;
    if (J.groups > kMidMaxG) J.groups = kMidMaxG;
    const int grid = (J.nlines + J.groups - 1) / J.groups;
    const size_t smem = mid_smem(n1, J.groups, false);
    const int T = J.groups * J.tl;
    if (mode == 0) zebra_mid_kernel<0><<<grid, T, smem, st>>>(J);
    else if (mode == 1) zebra_mid_kernel<1><<<grid, T, smem, st>>>(J);
    else zebra_mid_kernel<3><<<grid, T, smem, st>>>(J);
    CM_BYTES((double)J.nlines * n1 * (mode == 0 ? 40.0 : (mode == 3 ? 88.0 : 80.0)));
  }
  CM_LAUNCH_CHECK();
  return CM_OK;
}

}  // namespace cm
