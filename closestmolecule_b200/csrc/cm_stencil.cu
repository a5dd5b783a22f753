// Structured-grid kernels of the field-split / geometric-multigrid Newton linear solver.
//
// On RectangleMesh-numbered meshes (create_flat_boundary.py:65) every P1 row is a 7-point stencil
// (W, E, S, N and the two "right diagonal" neighbours), so the four blocks of the row-scaled P-Q
// Jacobian are stored as diagonals S[k][node] (no column index traffic) and the multigrid
// hierarchy of the p-block is built by Galerkin coarsening in stencil form.  Smoothing and the
// q-block solve are zebra x-line relaxations: batched tridiagonal solves along r2, each done by
// one CTA as two parallel scans over affine maps using factors precomputed once per Jacobian.
#include <stdlib.h>

#include "cm_common.cuh"
#include "cm_decl.cuh"
#include "cm_halo.cuh"

namespace cm {

// stencil slot k -> (dx,dy):  0:(-1,-1) 1:(0,-1) 2:(-1,0) 3:(0,0) 4:(+1,0) 5:(0,+1) 6:(+1,+1)
__host__ __device__ __forceinline__ int st_dx(int k) { return (k == 0 || k == 2) ? -1 : ((k == 4 || k == 6) ? 1 : 0); }
__host__ __device__ __forceinline__ int st_dy(int k) { return (k < 2) ? -1 : (k > 4 ? 1 : 0); }
__host__ __device__ __forceinline__ int st_slot(int dx, int dy) {
  if (dy == -1) return dx == -1 ? 0 : (dx == 0 ? 1 : -1);
  if (dy == 0) return dx == -1 ? 2 : (dx == 0 ? 3 : (dx == 1 ? 4 : -1));
  if (dy == 1) return dx == 0 ? 5 : (dx == 1 ? 6 : -1);
  return -1;
}

constexpr int kSB = 256;

// extra q-q couplings of the C0IP interior-facet pattern on right-diagonal meshes: the two vertices
// opposite an interior facet (..._C0IP.py:147): across the cell diagonal, a horizontal and a vertical edge
__host__ __device__ __forceinline__ int xq_dx(int e) { return e == 0 ? -1 : (e == 1 ? 1 : (e == 2 ? 1 : (e == 3 ? -1 : (e == 4 ? 2 : -2)))); }
__host__ __device__ __forceinline__ int xq_dy(int e) { return e == 0 ? 1 : (e == 1 ? -1 : (e == 2 ? 2 : (e == 3 ? -2 : (e == 4 ? 1 : -1)))); }
__host__ __device__ __forceinline__ int xq_slot(int dx, int dy) {
  for (int e = 0; e < 6; ++e)
    if (xq_dx(e) == dx && xq_dy(e) == dy) return e;
  return -1;
}

// ------------------------------------------------------------------------------------------------
// interleaved CSR (DOLFIN layout) -> stencil blocks PP, PQ, QP, QQ [7][nv] (unscaled: the Galerkin
// hierarchy must see the true FE operator) + row equilibration factors sp, sq = 1/max|row|, which
// the outer Krylov operator applies on the fly
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSB)
extract_blocks_kernel(const int nx, const int nv, const int* __restrict__ nrowptr,
                      const int* __restrict__ ncol, const double2* __restrict__ vals2,
                      double* __restrict__ PP, double* __restrict__ PQ, double* __restrict__ QP,
                      double* __restrict__ QQ, double* __restrict__ sp, double* __restrict__ sq,
                      int* __restrict__ err, const int vb, const int ve, double* __restrict__ QQx,
                      int* __restrict__ has_extra) {
  const int v = vb + blockIdx.x * kSB + threadIdx.x;
  if (v >= ve) return;
  const int n1 = nx + 1;
  const int rp = nrowptr[v], deg = nrowptr[v + 1] - rp;
  const double2* prow = vals2 + 2 * (size_t)rp;
  const double2* qrow = prow + deg;
  double mp = 0.0, mq = 0.0;
  for (int k = 0; k < deg; ++k) {
    const double2 a = prow[k], c = qrow[k];
    mp = fmax(mp, fmax(fabs(a.x), fabs(a.y)));
    mq = fmax(mq, fmax(fabs(c.x), fabs(c.y)));
  }
  const double s1 = mp > 0.0 ? 1.0 / mp : 1.0, s2 = mq > 0.0 ? 1.0 / mq : 1.0;
  sp[v] = s1;
  sq[v] = s2;
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    PP[(size_t)k * nv + v] = 0.0; PQ[(size_t)k * nv + v] = 0.0;
    QP[(size_t)k * nv + v] = 0.0; QQ[(size_t)k * nv + v] = 0.0;
  }
  const int ix = v % n1, iy = v / n1;
  if (QQx != nullptr)
    for (int e = 0; e < 6; ++e) QQx[(size_t)e * nv + v] = 0.0;
  for (int k = 0; k < deg; ++k) {
    const int w = ncol[rp + k];
    const int s = st_slot(w % n1 - ix, w / n1 - iy);
    if (s < 0) {
      // interior-facet (C0IP) couple: only the q-q entry can be non-zero there
      const int e = xq_slot(w % n1 - ix, w / n1 - iy);
      if (e < 0 || QQx == nullptr) { atomicExch(err, 1); continue; }
      const double2 a = prow[k], c = qrow[k];
      if (a.x != 0.0 || a.y != 0.0 || c.x != 0.0) { atomicExch(err, 1); continue; }
      QQx[(size_t)e * nv + v] = c.y;
      if (c.y != 0.0) *has_extra = 1;
      continue;
    }
    const double2 a = prow[k], c = qrow[k];
    PP[(size_t)s * nv + v] = a.x; PQ[(size_t)s * nv + v] = a.y;
    QP[(size_t)s * nv + v] = c.x; QQ[(size_t)s * nv + v] = c.y;
  }
}

// 7-point stand-in for the 13-point C0IP q-block, used by the line relaxation only: every interior-facet
// couple is replaced by the linear extrapolation of its end point through stencil neighbours,
//   q(-1,+1) ~ q(-1,0)+q(0,+1)-q(0,0),  q(+1,-1) ~ q(+1,0)+q(0,-1)-q(0,0),  q(+1,+2) ~ q(+1,+1)+q(0,+1)-q(0,0),
//   q(-1,-2) ~ q(-1,-1)+q(0,-1)-q(0,0), q(+2,+1) ~ q(+1,+1)+q(+1,0)-q(0,0), q(-2,-1) ~ q(-1,-1)+q(-1,0)-q(0,0),
// so that the lumped operator acts like the full one on affine functions (which the gradient-jump penalty
// of ..._C0IP.py:147 annihilates) and stays within a factor of two of it on the roughest modes
// (1-D analogue: [1,-4,6,-4,1] -> [-2,4,-2]).  Lagging the six diagonals inside the relaxation instead
// is not a convergent splitting.
__global__ void __launch_bounds__(kSB)
lump_extras_kernel(const int nv, const double* __restrict__ QQ, const double* __restrict__ QQx,
                   double* __restrict__ QQl) {
  const int v = blockIdx.x * kSB + threadIdx.x;
  if (v >= nv) return;
  double a[7], x[6];
#pragma unroll
  for (int k = 0; k < 7; ++k) a[k] = QQ[(size_t)k * nv + v];
#pragma unroll
  for (int e = 0; e < 6; ++e) x[e] = QQx[(size_t)e * nv + v];
  a[2] += x[0] + x[5];
  a[5] += x[0] + x[2];
  a[4] += x[1] + x[4];
  a[1] += x[1] + x[3];
  a[6] += x[2] + x[4];
  a[0] += x[3] + x[5];
  a[3] -= x[0] + x[1] + x[2] + x[3] + x[4] + x[5];
#pragma unroll
  for (int k = 0; k < 7; ++k) QQl[(size_t)k * nv + v] = a[k];
}

// interleaved b -> split, scaled:  out[v] = sp*b[2v], out[nv+v] = sq*b[2v+1]
__global__ void __launch_bounds__(kSB)
split_scale_kernel(const int nv, const double2* __restrict__ b, const double* __restrict__ sp,
                   const double* __restrict__ sq, double* __restrict__ out, const int vb, const int ve) {
  const int v = vb + blockIdx.x * kSB + threadIdx.x;
  if (v >= ve) return;
  const double2 t = b[v];
  out[v] = sp[v] * t.x;
  out[nv + v] = sq[v] * t.y;
}

__global__ void __launch_bounds__(kSB)
interleave_kernel(const int nv, const double* __restrict__ xs, double2* __restrict__ x, const int vb, const int ve) {
  const int v = vb + blockIdx.x * kSB + threadIdx.x;
  if (v >= ve) return;
  x[v] = make_double2(xs[v], xs[nv + v]);
}

// ------------------------------------------------------------------------------------------------
// stencil mat-vecs
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void stencil_offsets(int n1, int* off) {
  off[0] = -n1 - 1; off[1] = -n1; off[2] = -1; off[3] = 0; off[4] = 1; off[5] = n1; off[6] = n1 + 1;
}

__device__ __forceinline__ unsigned stencil_valid(int ix, int iy, int nx, int ny) {
  unsigned m = 0;
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    const int jx = ix + st_dx(k), jy = iy + st_dy(k);
    if (jx >= 0 && jx <= nx && jy >= 0 && jy <= ny) m |= 1u << k;
  }
  return m;
}

// row-scaled operator on split vectors: y_p = sp*(PP x_p + PQ x_q), y_q = sq*(QP x_p + QQ x_q)
__global__ void __launch_bounds__(kSB)
pq_stencil_spmv_kernel(const int nx, const int ny, const double* __restrict__ PP,
                       const double* __restrict__ PQ, const double* __restrict__ QP,
                       const double* __restrict__ QQ, const double* __restrict__ sp,
                       const double* __restrict__ sq, const double* x,
                       double* __restrict__ y, const int vb, const int ve,
                       const double* __restrict__ QQx, const int* __restrict__ has_extra, const Halo H,
                       double* __restrict__ zcopy, const long long zstride, const double* __restrict__ jptr,
                       const int halo_channels, const double* __restrict__ xscale) {
  const int nv = (nx + 1) * (ny + 1);
  const int n1 = nx + 1;
  if (H.active) {
    // strip-partitioned rank: the first / last owned row read the ghost rows the neighbours' last half sweeps
    // pushed (cm_halo.cuh); only the CTAs that touch those rows look at the flags
    const int c0 = vb + blockIdx.x * kSB, c1 = min(c0 + kSB, ve);
    if (threadIdx.x == 0) {
      for (int ch = 0; ch < halo_channels; ++ch) {   // p- and q-part were produced on their own channels
        Halo Hc = H;
        Hc.chan = ch;
        if (c0 < vb + n1) halo_wait_below(Hc);
        if (c1 > ve - n1) halo_wait_above(Hc);
      }
    }
    __syncthreads();
  }
  const int v = vb + blockIdx.x * kSB + threadIdx.x;
  if (v >= ve) return;
  const unsigned valid = stencil_valid(v % n1, v / n1, nx, ny);
  int off[7];
  stencil_offsets(n1, off);
  double yp = 0.0, yq = 0.0;
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    if (valid & (1u << k)) {
      const double xp = x[v + off[k]], xq = x[nv + v + off[k]];
      const size_t e = (size_t)k * nv + v;
      yp += PP[e] * xp + PQ[e] * xq;
      yq += QP[e] * xp + QQ[e] * xq;
    }
  }
  if (QQx != nullptr && *has_extra) {
    const int ix = v % n1, iy = v / n1;
#pragma unroll
    for (int e = 0; e < 6; ++e) {
      const int jx = ix + xq_dx(e), jy = iy + xq_dy(e);
      if (jx >= 0 && jx <= nx && jy >= 0 && jy <= ny)
        yq += QQx[(size_t)e * nv + v] * x[nv + jy * n1 + jx];
    }
  }
  // xscale: x came out of a preconditioner application to the UNNORMALISED basis vector; the operator and the
  // preconditioner are linear, so the normalisation 1 / ||V_j|| is applied here, to y and to the kept copy of x
  const double xs = xscale != nullptr ? xscale[(int)(*jptr)] : 1.0;
  y[v] = sp[v] * yp * xs;
  y[nv + v] = sq[v] * yq * xs;
  if (zcopy != nullptr) {   // flexible GMRES: keep the preconditioned basis vector z_j = x (cm_gmres_dev.cu)
    double* zj = zcopy + (size_t)zstride * (size_t)(int)(*jptr);
    zj[v] = x[v] * xs;
    zj[nv + v] = x[nv + v] * xs;
  }
}

// out = r / s (undo the row equilibration before the unscaled preconditioner)
__global__ void __launch_bounds__(kSB)
unscale_kernel(const int nv, const double* __restrict__ r, const double* __restrict__ sp,
               const double* __restrict__ sq, double* __restrict__ out, const int vb, const int ve) {
  const int v = vb + blockIdx.x * kSB + threadIdx.x;
  if (v >= ve) return;
  out[v] = r[v] / sp[v];
  out[nv + v] = r[nv + v] / sq[v];
}

// same with the source vector chosen on the device: r = vscale[j] * (rbase + stride * j), j = (int)(*jptr) (the
// Krylov basis vector of the current GMRES iteration, cm_gmres_dev.cu); jptr == nullptr: r = rbase
__global__ void __launch_bounds__(kSB)
unscale_dev_kernel(const int nv, const double* __restrict__ rbase, const long long stride,
                   const double* __restrict__ jptr, const double* __restrict__ vscale,
                   const double* __restrict__ sp, const double* __restrict__ sq, double* __restrict__ out,
                   const int vb, const int ve) {
  const int v = vb + blockIdx.x * kSB + threadIdx.x;
  if (v >= ve) return;
  const int j = jptr ? (int)(*jptr) : 0;
  const double* r = rbase + (size_t)stride * (size_t)j;
  const double a = vscale ? vscale[j] : 1.0;   // the basis vectors are stored unnormalised
  out[v] = a * r[v] / sp[v];
  out[nv + v] = a * r[nv + v] / sq[v];
}

// r = b - A x  (scalar 7-point stencil; b may be null -> r = -A x)
__global__ void __launch_bounds__(kSB)
stencil_residual_kernel(const int nx, const int ny, const double* __restrict__ A,
                        const double* __restrict__ b, const double* __restrict__ x,
                        double* __restrict__ r, const int vb, const int ve) {
  const int nv = (nx + 1) * (ny + 1);
  const int v = vb + blockIdx.x * kSB + threadIdx.x;
  if (v >= ve) return;
  const int n1 = nx + 1;
  const unsigned valid = stencil_valid(v % n1, v / n1, nx, ny);
  int off[7];
  stencil_offsets(n1, off);
  double s = b ? b[v] : 0.0;
#pragma unroll
  for (int k = 0; k < 7; ++k)
    if (valid & (1u << k)) s -= A[(size_t)k * nv + v] * x[v + off[k]];
  r[v] = s;
}

// ------------------------------------------------------------------------------------------------
// transfers: P1 interpolation on nested right-diagonal meshes (weights 1, 1/2 on the 6 neighbours)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSB)
restrict_kernel(const int nxc, const int nyc, const double* __restrict__ rf, double* __restrict__ rc, const int vb, const int ve) {
  const int I = vb + blockIdx.x * kSB + threadIdx.x;
  if (I >= ve) return;
  const int X = I % (nxc + 1), Y = I / (nxc + 1);
  const int nxf = 2 * nxc, nyf = 2 * nyc, n1 = nxf + 1;
  const int ix = 2 * X, iy = 2 * Y;
  const int c = iy * n1 + ix;
  const unsigned valid = stencil_valid(ix, iy, nxf, nyf);
  int off[7];
  stencil_offsets(n1, off);
  double s = rf[c];
#pragma unroll
  for (int k = 0; k < 7; ++k)
    if (k != 3 && (valid & (1u << k))) s += 0.5 * rf[c + off[k]];
  rc[I] = s;
}

__global__ void __launch_bounds__(kSB)
prolong_add_kernel(const int nxc, const int nyc, const double* __restrict__ ec, double* __restrict__ xf, const int vb, const int ve) {
  const int nxf = 2 * nxc, nyf = 2 * nyc, n1 = nxf + 1;
  const int i = vb + blockIdx.x * kSB + threadIdx.x;
  if (i >= ve) return;
  const int ix = i % n1, iy = i / n1;
  const int X = ix >> 1, Y = iy >> 1, m1 = nxc + 1;
  const int ox = ix & 1, oy = iy & 1;
  const int c0 = Y * m1 + X;
  double e;
  if (!ox && !oy) e = ec[c0];
  else if (ox && !oy) e = 0.5 * (ec[c0] + ec[c0 + 1]);
  else if (!ox && oy) e = 0.5 * (ec[c0] + ec[c0 + m1]);
  else e = 0.5 * (ec[c0] + ec[c0 + m1 + 1]);  // midpoint of the coarse right diagonal
  xf[i] += e;
}

// Galerkin coarse operator Ac = P^T A P in stencil form (one thread per coarse node)
__global__ void __launch_bounds__(kSB)
rap_kernel(const int nxc, const int nyc, const double* __restrict__ Af, double* __restrict__ Ac, const int vb, const int ve) {
  const int nc = (nxc + 1) * (nyc + 1);
  const int I = vb + blockIdx.x * kSB + threadIdx.x;
  if (I >= ve) return;
  const int X = I % (nxc + 1), Y = I / (nxc + 1);
  const int nxf = 2 * nxc, nyf = 2 * nyc, n1 = nxf + 1;
  const size_t nf = (size_t)n1 * (nyf + 1);
  double acc[7] = {0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int a = 0; a < 7; ++a) {
    const int ax = st_dx(a), ay = st_dy(a);
    const int ix = 2 * X + ax, iy = 2 * Y + ay;
    if (ix < 0 || ix > nxf || iy < 0 || iy > nyf) continue;
    const double wa = (a == 3) ? 1.0 : 0.5;
    const size_t i = (size_t)iy * n1 + ix;
#pragma unroll
    for (int e = 0; e < 7; ++e) {
      const double av = wa * Af[(size_t)e * nf + i];
      const int sx = ax + st_dx(e), sy = ay + st_dy(e);  // fine offset of column j from 2I
#pragma unroll
      for (int D = 0; D < 7; ++D) {
        const int bx = sx - 2 * st_dx(D), by = sy - 2 * st_dy(D);
        if (bx < -1 || bx > 1 || by < -1 || by > 1) continue;
        const int bs = st_slot(bx, by);
        if (bs < 0) continue;
        acc[D] += av * ((bs == 3) ? 1.0 : 0.5);
      }
    }
  }
#pragma unroll
  for (int D = 0; D < 7; ++D) Ac[(size_t)D * nc + I] = acc[D];
}

// ------------------------------------------------------------------------------------------------
// x-line tridiagonal factors (once per Jacobian): T = L U per grid line, no pivoting
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
line_factor_kernel(const int nx, const int ny, const double* __restrict__ A, double* __restrict__ lf,
                   double* __restrict__ iu, double* __restrict__ cu, int* __restrict__ err) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j > ny) return;
  const int n1 = nx + 1;
  const size_t nv = (size_t)n1 * (ny + 1);
  const size_t base = (size_t)j * n1;
  double uprev = 1.0, cprev = 0.0;
  bool bad = false;
  for (int i = 0; i <= nx; ++i) {
    const size_t v = base + i;
    const double a = (i > 0) ? A[2 * nv + v] : 0.0;
    const double d = A[3 * nv + v];
    const double c = (i < nx) ? A[4 * nv + v] : 0.0;
    const double l = (i > 0) ? a / uprev : 0.0;
    const double u = d - l * cprev;
    const double inv = 1.0 / u;
    if (!isfinite(inv) || fabs(u) < 1e-300) bad = true;
    lf[v] = l;
    iu[v] = inv;
    cu[v] = c * inv;
    uprev = u;
    cprev = c;
  }
  if (bad) atomicExch(err, 2);
}

// Parallel version: the pivot recurrence u_i = d_i - a_i c_{i-1} / u_{i-1} is a Moebius map of
// u_{i-1}; compositions are 2x2 matrix products, so one CTA per grid line evaluates it as a
// block-wide scan over (projectively normalised) matrices, then every thread replays its own
// segment with the exact incoming pivot and writes l_i, 1/u_i, c_i/u_i.
struct Mob {
  double a, b, c, d;  // u -> (a u + b) / (c u + d)
};
// a Moebius map is defined up to a common factor: keep the entries near 1 by an EXACT power-of-two scaling
// (exponent arithmetic, no fp64 division: the factorisation kernel used to spend most of its time in 1/s)
__device__ __forceinline__ Mob mob_norm(Mob m) {
  const double s = fmax(fmax(fabs(m.a), fabs(m.b)), fmax(fabs(m.c), fabs(m.d)));
  if (s > 0.0 && isfinite(s)) {
    int f = 2046 - ((__double2hiint(s) >> 20) & 0x7ff);   // biased exponent of 2^-e, e = exponent of s
    f = f < 1 ? 1 : (f > 2046 ? 2046 : f);
    const double r = __hiloint2double(f << 20, 0);
    m.a *= r; m.b *= r; m.c *= r; m.d *= r;
  }
  return m;
}
__device__ __forceinline__ Mob mob_compose(const Mob l, const Mob e) {  // l after e
  Mob r;
  r.a = l.a * e.a + l.b * e.c;
  r.b = l.a * e.b + l.b * e.d;
  r.c = l.c * e.a + l.d * e.c;
  r.d = l.c * e.b + l.d * e.d;
  return mob_norm(r);
}
__device__ __forceinline__ Mob mob_shfl_up(const Mob m, int o) {
  Mob r;
  r.a = __shfl_up_sync(0xffffffffu, m.a, o);
  r.b = __shfl_up_sync(0xffffffffu, m.b, o);
  r.c = __shfl_up_sync(0xffffffffu, m.c, o);
  r.d = __shfl_up_sync(0xffffffffu, m.d, o);
  return r;
}

__global__ void __launch_bounds__(256)
line_factor_scan_kernel(const int nx, const int ny, const int E, const double* __restrict__ A,
                        double* __restrict__ lf, double* __restrict__ iu, double* __restrict__ cu,
                        int* __restrict__ err, const int j0) {
  extern __shared__ double sm[];
  const int n1 = nx + 1;
  double* s_a = sm;
  double* s_d = sm + n1;
  double* s_c = sm + 2 * n1;
  __shared__ double sh[4 * 32];
  const int j = j0 + blockIdx.x;
  const size_t nv = (size_t)n1 * (ny + 1);
  const size_t base = (size_t)j * n1;
  const int T = blockDim.x, tid = threadIdx.x;
  const int lane = tid & 31, wid = tid >> 5, nw = T >> 5;
  for (int i = tid; i < n1; i += T) {
    const size_t v = base + i;
    s_a[i] = (i > 0) ? A[2 * nv + v] : 0.0;
    s_d[i] = A[3 * nv + v];
    s_c[i] = (i < nx) ? A[4 * nv + v] : 0.0;
  }
  __syncthreads();
  const int i0 = tid * E, i1 = min(i0 + E, n1);
  Mob m{1.0, 0.0, 0.0, 1.0};
  for (int i = i0; i < i1; ++i) {
    const double d = s_d[i], ac = (i > 0) ? s_a[i] * s_c[i - 1] : 0.0;
    // M_i = [[d, -ac], [1, 0]] applied after m
    Mob r;
    r.a = d * m.a - ac * m.c;
    r.b = d * m.b - ac * m.d;
    r.c = m.a;
    r.d = m.b;
    m = mob_norm(r);
  }
  // inclusive scan in thread order
  Mob inc = m;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const Mob p = mob_shfl_up(inc, o);
    if (lane >= o) inc = mob_compose(inc, p);
  }
  if (lane == 31) { sh[4 * wid] = inc.a; sh[4 * wid + 1] = inc.b; sh[4 * wid + 2] = inc.c; sh[4 * wid + 3] = inc.d; }
  __syncthreads();
  if (wid == 0) {
    Mob t = (lane < nw) ? Mob{sh[4 * lane], sh[4 * lane + 1], sh[4 * lane + 2], sh[4 * lane + 3]}
                        : Mob{1.0, 0.0, 0.0, 1.0};
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const Mob p = mob_shfl_up(t, o);
      if (lane >= o) t = mob_compose(t, p);
    }
    if (lane < nw) { sh[4 * lane] = t.a; sh[4 * lane + 1] = t.b; sh[4 * lane + 2] = t.c; sh[4 * lane + 3] = t.d; }
  }
  __syncthreads();
  Mob ex = mob_shfl_up(inc, 1);
  if (lane == 0) ex = Mob{1.0, 0.0, 0.0, 1.0};
  if (wid > 0) ex = mob_compose(ex, Mob{sh[4 * (wid - 1)], sh[4 * (wid - 1) + 1], sh[4 * (wid - 1) + 2], sh[4 * (wid - 1) + 3]});
  // pivot entering this segment: the prefix map applied to u_{-1} = 1 (a_0 = 0 makes it irrelevant)
  // 1 / u_{i-1} entering the segment (one division per unknown below: l_i = a_i / u_{i-1} reuses the inverse)
  double iprev = (ex.c + ex.d) / (ex.a + ex.b);
  bool bad = false;
  // the factors replace a / d / c in shared memory and leave by coalesced stores (17 consecutive unknowns per
  // thread written straight to global memory were 136-byte strided 8-byte stores: the kernel was LSU bound)
  double cprev = (i0 > 0 && i0 < n1) ? s_c[i0 - 1] : 0.0;
  __syncthreads();   // every thread has read its left neighbour's c before anyone overwrites it
  for (int i = i0; i < i1; ++i) {
    const double a = s_a[i], c = s_c[i];
    const double l = (i > 0) ? a * iprev : 0.0;
    const double u = s_d[i] - ((i > 0) ? l * cprev : 0.0);
    const double inv = 1.0 / u;
    if (!isfinite(inv) || fabs(u) < 1e-300) bad = true;
    s_a[i] = l;
    s_d[i] = inv;
    s_c[i] = c * inv;
    cprev = c;
    iprev = inv;
  }
  if (bad) atomicExch(err, 2);
  __syncthreads();
  for (int i = tid; i < n1; i += T) {
    const size_t v = base + i;
    lf[v] = s_a[i];
    iu[v] = s_d[i];
    cu[v] = s_c[i];
  }
}

// ------------------------------------------------------------------------------------------------
// zebra x-line relaxation: for every grid line of one colour solve
//   T x_line = b_line - (A - T) x      (couplings to the lines above/below use the current x)
// One CTA per line.  Forward and backward substitution are first-order linear recurrences,
// evaluated as block-wide scans over affine maps (A,B): v_out = A*v_in + B.
// ------------------------------------------------------------------------------------------------
struct Affine {
  double a, b;
};
__device__ __forceinline__ Affine compose(const Affine later, const Affine earlier) {
  return Affine{later.a * earlier.a, later.a * earlier.b + later.b};
}

// exclusive scan over threads in thread order; returns the value entering this thread's segment
// when the value entering thread 0 is 0.  sh must hold 2*32 doubles.
__device__ __forceinline__ double affine_exclusive_scan(Affine m, double* sh) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double pa = __shfl_up_sync(0xffffffffu, m.a, o);
    const double pb = __shfl_up_sync(0xffffffffu, m.b, o);
    if (lane >= o) m = compose(m, Affine{pa, pb});
  }
  if (lane == 31) { sh[2 * wid] = m.a; sh[2 * wid + 1] = m.b; }
  __syncthreads();
  if (wid == 0) {
    Affine t = (lane < nw) ? Affine{sh[2 * lane], sh[2 * lane + 1]} : Affine{1.0, 0.0};
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double pa = __shfl_up_sync(0xffffffffu, t.a, o);
      const double pb = __shfl_up_sync(0xffffffffu, t.b, o);
      if (lane >= o) t = compose(t, Affine{pa, pb});
    }
    if (lane < nw) { sh[2 * lane] = t.a; sh[2 * lane + 1] = t.b; }
  }
  __syncthreads();
  // inclusive prefix of the previous lane, composed with the inclusive prefix of previous warps
  double ea = __shfl_up_sync(0xffffffffu, m.a, 1);
  double eb = __shfl_up_sync(0xffffffffu, m.b, 1);
  if (lane == 0) { ea = 1.0; eb = 0.0; }
  double r = eb;
  if (wid > 0) r = ea * sh[2 * (wid - 1) + 1] + eb;
  __syncthreads();
  return r;
}

constexpr int kZebraPrefetch = 9;
__device__ unsigned long long g_zdbg[2048 * 8];
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define ZT(k) do { if ((dbg & 32) && threadIdx.x == 0) g_zdbg[blockIdx.x * 8 + (k)] = gtime(); } while (0)  // striped elements per thread whose (1/u, c/u) are prefetched

template <bool XZERO>
__global__ void __launch_bounds__(512, 2)
zebra_line_kernel(const int nx, const int ny, const int colour, const int E,
                  const double* __restrict__ A, const double* __restrict__ lf,
                  const double* __restrict__ iu, const double* __restrict__ cu,
                  const double* __restrict__ b, double* __restrict__ x, const int dbg,
                  const int jfirst, const int jend) {
  extern __shared__ double sm[];
  const int n1 = nx + 1;
  double* s_z = sm;
  double* s_c = sm + n1;
  __shared__ double sh[64];
  const int j = jfirst + 2 * blockIdx.x;  // jfirst has the parity of `colour`
  if (j >= jend) return;
  (void)colour;
  const size_t nv = (size_t)n1 * (ny + 1);
  const size_t base = (size_t)j * n1;
  const int T = blockDim.x, tid = threadIdx.x;
  ZT(0);
  // right-hand side: b minus the couplings to the neighbouring lines.  Branch-free: stencil entries
  // that point outside the grid are stored as zeros, so the neighbour index is only clamped -- every
  // load of an iteration is independent and can be in flight at once.
  {
    const size_t lo = (size_t)(j > 0 ? j - 1 : 0) * n1;
    const size_t up = (size_t)(j < ny ? j + 1 : ny) * n1;
    const double* __restrict__ A0 = A + base;
    const double* __restrict__ A1 = A + nv + base;
    const double* __restrict__ A5 = A + 5 * nv + base;
    const double* __restrict__ A6 = A + 6 * nv + base;
#pragma unroll 3
    for (int i = tid; i < n1; i += T) {
      double r = b[base + i];
      const double l = lf[base + i];
      if (!XZERO) {
        const int im = i > 0 ? i - 1 : 0, ip = i < nx ? i + 1 : nx;
        const double a0 = A0[i], a1 = A1[i], a5 = A5[i], a6 = A6[i];
        const double x0 = x[lo + im], x1 = x[lo + i], x5 = x[up + i], x6 = x[up + ip];
        r -= a0 * x0 + a1 * x1 + a5 * x5 + a6 * x6;
      }
      s_z[i] = r;
      s_c[i] = l;
    }
  }
  // prefetch the back-substitution factors of this thread's striped elements: their latency hides
  // behind the forward scan
  double piu[kZebraPrefetch], pcu[kZebraPrefetch];
#pragma unroll
  for (int k = 0; k < kZebraPrefetch; ++k) {
    const int i = tid + k * T;
    if (i < n1) { piu[k] = iu[base + i]; pcu[k] = cu[base + i]; }
  }
  __syncthreads();
  ZT(1);
  // forward: z_i = r_i - l_i z_{i-1}
  {
    const int i0 = tid * E, i1 = min(i0 + E, n1);
    Affine m{1.0, 0.0};
    for (int i = i0; i < i1; ++i) {
      const double l = s_c[i];
      m.b = s_z[i] - l * m.b;
      m.a = -l * m.a;
    }
    double z = affine_exclusive_scan(m, sh);
    for (int i = i0; i < i1; ++i) {
      z = s_z[i] - s_c[i] * z;
      s_z[i] = z;
    }
  }
  __syncthreads();
  ZT(2);
#pragma unroll
  for (int k = 0; k < kZebraPrefetch; ++k) {
    const int i = tid + k * T;
    if (i < n1) { s_z[i] *= piu[k]; s_c[i] = pcu[k]; }
  }
  for (int i = tid + kZebraPrefetch * T; i < n1; i += T) {
    s_z[i] *= iu[base + i];
    s_c[i] = cu[base + i];
  }
  __syncthreads();
  ZT(3);
  // backward: x_i = w_i - cu_i x_{i+1}; thread t takes segment T-1-t so scan order = thread order
  {
    const int seg = T - 1 - tid;
    const int i0 = seg * E, i1 = min(i0 + E, n1);
    Affine m{1.0, 0.0};
    for (int i = i1 - 1; i >= i0; --i) {
      const double c = s_c[i];
      m.b = s_z[i] - c * m.b;
      m.a = -c * m.a;
    }
    double xn = affine_exclusive_scan(m, sh);
    for (int i = i1 - 1; i >= i0; --i) {
      xn = s_z[i] - s_c[i] * xn;
      s_z[i] = xn;
    }
  }
  __syncthreads();
  ZT(4);
  for (int i = tid; i < n1; i += T) x[base + i] = s_z[i];
  ZT(5);
}

int st_zebra_debug_times(unsigned long long* host_out, int n) {
  return cudaMemcpyFromSymbol(host_out, g_zdbg, sizeof(unsigned long long) * n) == cudaSuccess ? 0 : -2;
}

// ------------------------------------------------------------------------------------------------
// Two-kernel variant of the zebra half sweep, kept for the 13-point (C0IP) q-block whose six extra
// diagonals do not fit the fused kernel of cm_mg.cu: (1) offline_rhs_kernel, a streaming pass
// rhs = b - (A - T) x on the lines of one colour; (2) line_solve_tma_kernel on that right-hand side.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kSB)
offline_rhs_kernel(const int nx, const int ny, const int jfirst, const int nlines,
                   const double* __restrict__ A, const double* __restrict__ b,
                   const double* __restrict__ x, double* __restrict__ out,
                   const double* __restrict__ Ax, const int* __restrict__ has_extra) {
  const int n1 = nx + 1;
  const long long t = blockIdx.x * (long long)kSB + threadIdx.x;
  if (t >= (long long)nlines * n1) return;
  const int l = (int)(t / n1), i = (int)(t - (long long)l * n1);
  const int j = jfirst + 2 * l;
  const size_t nv = (size_t)n1 * (ny + 1);
  const size_t v = (size_t)j * n1 + i;
  const size_t lo = (size_t)(j > 0 ? j - 1 : 0) * n1, up = (size_t)(j < ny ? j + 1 : ny) * n1;
  const int im = i > 0 ? i - 1 : 0, ip = i < nx ? i + 1 : nx;
  const double a0 = A[v], a1 = A[nv + v], a5 = A[5 * nv + v], a6 = A[6 * nv + v];
  double r = b[v] - (a0 * x[lo + im] + a1 * x[lo + i] + a5 * x[up + i] + a6 * x[up + ip]);
  if (Ax != nullptr && *has_extra) {
#pragma unroll
    for (int e = 0; e < 6; ++e) {
      const int jx = i + xq_dx(e), jy = j + xq_dy(e);
      if (jx >= 0 && jx <= nx && jy >= 0 && jy <= ny) r -= Ax[(size_t)e * nv + v] * x[(size_t)jy * n1 + jx];
    }
  }
  out[v] = r;
}

// ------------------------------------------------------------------------------------------------
// Register-resident line solves: thread t owns E consecutive unknowns of the line and keeps the right-hand
// side and the three factor arrays in registers, so shared memory only carries the 2 x 32 warp totals of
// the scans.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double affine_exclusive_scan_dir(Affine m, double* sh, const bool rev) {
  // exclusive scan in thread order (rev: in reverse thread order); returns the incoming value
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int pos = rev ? 31 - lane : lane;       // position inside the warp in scan order
  const int wpos = rev ? nw - 1 - wid : wid;    // warp position in scan order
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double pa = rev ? __shfl_down_sync(0xffffffffu, m.a, o) : __shfl_up_sync(0xffffffffu, m.a, o);
    const double pb = rev ? __shfl_down_sync(0xffffffffu, m.b, o) : __shfl_up_sync(0xffffffffu, m.b, o);
    if (pos >= o) m = compose(m, Affine{pa, pb});
  }
  if (pos == 31) { sh[2 * wpos] = m.a; sh[2 * wpos + 1] = m.b; }
  __syncthreads();
  if (wid == 0) {
    Affine t = (lane < nw) ? Affine{sh[2 * lane], sh[2 * lane + 1]} : Affine{1.0, 0.0};
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double pa = __shfl_up_sync(0xffffffffu, t.a, o);
      const double pb = __shfl_up_sync(0xffffffffu, t.b, o);
      if (lane >= o) t = compose(t, Affine{pa, pb});
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

#define HT(k) do { if ((dbg & 32) && tid == 0 && l < 2048) g_zdbg[l * 8 + (k)] = gtime(); } while (0)

// ------------------------------------------------------------------------------------------------
// TMA line solve (right-hand side given): the four arrays of the next line arrive by four
// cp.async.bulk copies (UBLKCP) issued by one thread and tracked by an mbarrier -- no per-thread
// copy instructions at all (8-byte cp.async is issue-bound: ~2.6 us per line).  Lines start on
// 8-byte boundaries only, so the copy starts at the preceding even element and the shared-memory
// view is shifted by `sh`; an odd trailing element is fetched by a plain load.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, void* mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(mbar)) : "memory");
}

template <int E>
__global__ void __launch_bounds__(896, 1)
line_solve_tma_kernel(const int nx, const int ny, const int jfirst, const int nlines,
                      const double* __restrict__ lf, const double* __restrict__ iu,
                      const double* __restrict__ cu, const double* __restrict__ rhs,
                      double* __restrict__ x, const int dbg) {
  extern __shared__ __align__(16) double sm[];  // 4 input buffers of npad doubles + s_x
  __shared__ double sh[64];
  __shared__ __align__(8) unsigned long long mbar;
  const int n1 = nx + 1;
  const int npad = (n1 + 3) & ~1;
  double* buf[4] = {sm, sm + npad, sm + 2 * npad, sm + 3 * npad};
  double* s_x = sm + 4 * npad;
  const double* src[4] = {rhs, lf, iu, cu};
  const int T = blockDim.x, tid = threadIdx.x;
  const int i0 = tid * E;
  const int ne = max(0, min(E, n1 - i0));
  int l = blockIdx.x;
  if (l >= nlines) return;
  if (tid == 0) {
    // one phase = the four arrays of one line: four arrive.expect_tx (one per array) complete it
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 4;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  unsigned parity = 0;
  // per array: is the first element of the line on a 16-byte boundary? (the arrays themselves may
  // start on odd 8-byte offsets, e.g. the q-part of a split vector with an odd vertex count)
  auto shift_of = [&](int a, size_t base) -> int {
    return (int)((reinterpret_cast<unsigned long long>(src[a] + base) >> 3) & 1ull);
  };
  // array a of `line` -> buf[a]; called by all threads after a barrier that ended the reads of buf[a]
  auto issue_one = [&](int a, int line) {
    const size_t base = (size_t)(jfirst + 2 * line) * n1;
    const int shf = shift_of(a, base);
    const int cnt = (n1 + shf) & ~1;  // even number of elements copied by TMA
    if (tid == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"((unsigned)cnt * 8u) : "memory");
      tma_bulk_g2s(buf[a], src[a] + base - shf, (unsigned)cnt * 8u, &mbar);
    }
    if (tid == 32 && cnt < n1 + shf) buf[a][cnt] = src[a][base - shf + cnt];  // odd trailing element
  };
#pragma unroll
  for (int a = 0; a < 4; ++a) issue_one(a, l);
  for (; l < nlines; l += gridDim.x) {
    const size_t lbase = (size_t)(jfirst + 2 * l) * n1;
    const int sh0 = shift_of(0, lbase), sh1 = shift_of(1, lbase), sh2 = shift_of(2, lbase), sh3 = shift_of(3, lbase);
    const bool more = l + gridDim.x < nlines;
    const int lnext = l + gridDim.x;
    HT(0);
    {  // wait for the bulk copies of this line (bounded: a lost copy must not hang the GPU)
      unsigned done = 0;
      long long spins = 0;
      while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&mbar)), "r"(parity) : "memory");
        if (!done && ++spins > (1ll << 26)) __trap();
      }
      parity ^= 1u;
    }
    __syncthreads();  // the plain-load tail elements are visible too
    HT(1);
    // shared -> registers array by array; each buffer is refilled for the CTA's next line as soon as it
    // has been read, so the copies of the next line are in flight during the rest of this phase and the scans
    double r[E], lv[E], iv[E], cv[E];
#pragma unroll
    for (int e = 0; e < E; ++e)
      if (e < ne) r[e] = buf[0][sh0 + i0 + e];
    __syncthreads();
    if (more) issue_one(0, lnext);
#pragma unroll
    for (int e = 0; e < E; ++e)
      if (e < ne) lv[e] = buf[1][sh1 + i0 + e];
    __syncthreads();
    if (more) issue_one(1, lnext);
#pragma unroll
    for (int e = 0; e < E; ++e)
      if (e < ne) iv[e] = buf[2][sh2 + i0 + e];
    __syncthreads();
    if (more) issue_one(2, lnext);
#pragma unroll
    for (int e = 0; e < E; ++e)
      if (e < ne) cv[e] = buf[3][sh3 + i0 + e];
    __syncthreads();
    if (more) issue_one(3, lnext);
    HT(2);
    Affine m{1.0, 0.0};
#pragma unroll
    for (int e = 0; e < E; ++e)
      if (e < ne) { m.b = r[e] - lv[e] * m.b; m.a = -lv[e] * m.a; }
    double z = affine_exclusive_scan_dir(m, sh, false);
#pragma unroll
    for (int e = 0; e < E; ++e)
      if (e < ne) { z = r[e] - lv[e] * z; r[e] = z * iv[e]; }
    HT(3);
    m = Affine{1.0, 0.0};
#pragma unroll
    for (int e = E - 1; e >= 0; --e)
      if (e < ne) { m.b = r[e] - cv[e] * m.b; m.a = -cv[e] * m.a; }
    double xn = affine_exclusive_scan_dir(m, sh, true);
#pragma unroll
    for (int e = E - 1; e >= 0; --e)
      if (e < ne) { xn = r[e] - cv[e] * xn; s_x[i0 + e] = xn; }
    __syncthreads();
    HT(4);
    for (int i = tid; i < n1; i += T) x[lbase + i] = s_x[i];
    HT(5);
  }
}

// ------------------------------------------------------------------------------------------------
// coarsest level: dense inverse by Gauss-Jordan with partial pivoting (one CTA), then x = Ainv b
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
coarse_invert_kernel(const int nx, const int ny, const double* __restrict__ A, double* __restrict__ M,
                     double* __restrict__ Ainv, int* __restrict__ err) {
  // M: n x 2n work matrix in global memory ([A | I]), row-major; n <= 256
  const int n1 = nx + 1, n = n1 * (ny + 1), w = 2 * n;
  const int tid = threadIdx.x, T = blockDim.x;
  __shared__ int s_piv;
  __shared__ double s_pv;
  __shared__ double s_col[256];
  for (int e = tid; e < n * w; e += T) M[e] = 0.0;
  __syncthreads();
  for (int v = tid; v < n; v += T) {
    const int ix = v % n1, iy = v / n1;
    for (int k = 0; k < 7; ++k) {
      const int jx = ix + st_dx(k), jy = iy + st_dy(k);
      if (jx < 0 || jx > nx || jy < 0 || jy > ny) continue;
      M[v * w + jy * n1 + jx] = A[(size_t)k * n + v];
    }
    M[v * w + n + v] = 1.0;
  }
  __syncthreads();
  for (int c = 0; c < n; ++c) {
    if (tid == 0) {
      int p = c;
      double best = fabs(M[c * w + c]);
      for (int r = c + 1; r < n; ++r) {
        const double t = fabs(M[r * w + c]);
        if (t > best) { best = t; p = r; }
      }
      s_piv = p;
      s_pv = M[p * w + c];
      if (!(best > 0.0)) atomicExch(err, 3);
    }
    __syncthreads();
    const int p = s_piv;
    const double inv = 1.0 / s_pv;
    if (p != c) {
      for (int e = tid; e < w; e += T) {
        const double t = M[c * w + e];
        M[c * w + e] = M[p * w + e];
        M[p * w + e] = t;
      }
    }
    __syncthreads();
    for (int e = tid; e < w; e += T) M[c * w + e] *= inv;
    for (int r = tid; r < n; r += T) s_col[r] = (r == c) ? 0.0 : M[r * w + c];
    __syncthreads();
    for (int e = tid; e < n * w; e += T) {
      const int r = e / w, k = e - r * w;
      const double f = s_col[r];
      if (f != 0.0) M[e] -= f * M[c * w + k];
    }
    __syncthreads();
  }
  for (int e = tid; e < n * n; e += T) Ainv[e] = M[(e / n) * w + n + (e % n)];
}

// the same elimination with the n x 2n work matrix in shared memory and a warp-wide pivot search (n <= 112):
// ~0.4 us per column instead of ~3.5 us through global memory with a serial search (156 us for the 45 x 45 operator
// of the 16 Mi-cell hierarchy, once per Newton step on EVERY rank: 2 % of the 8-GPU step)
__global__ void __launch_bounds__(256)
coarse_invert_smem_kernel(const int nx, const int ny, const double* __restrict__ A, double* __restrict__ Ainv,
                          int* __restrict__ err) {
  extern __shared__ double sM[];   // n x (2n + 1): the odd row stride keeps column accesses conflict free
  const int n1 = nx + 1, n = n1 * (ny + 1), w = 2 * n, ld = w + 1;
  const int tid = threadIdx.x, T = blockDim.x;
  __shared__ int s_piv;
  __shared__ double s_col[128];
  for (int e = tid; e < n * ld; e += T) sM[e] = 0.0;
  __syncthreads();
  for (int v = tid; v < n; v += T) {
    const int ix = v % n1, iy = v / n1;
    for (int k = 0; k < 7; ++k) {
      const int jx = ix + st_dx(k), jy = iy + st_dy(k);
      if (jx < 0 || jx > nx || jy < 0 || jy > ny) continue;
      sM[v * ld + jy * n1 + jx] = A[(size_t)k * n + v];
    }
    sM[v * ld + n + v] = 1.0;
  }
  __syncthreads();
  for (int c = 0; c < n; ++c) {
    if (tid < 32) {   // partial pivoting: first row of maximal modulus (the order the serial search finds)
      double best = -1.0;
      int p = c;
      for (int r = c + tid; r < n; r += 32) {
        const double t = fabs(sM[r * ld + c]);
        if (t > best) { best = t; p = r; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_down_sync(0xffffffffu, best, o);
        const int op = __shfl_down_sync(0xffffffffu, p, o);
        if (ob > best || (ob == best && op < p)) { best = ob; p = op; }
      }
      if (tid == 0) {
        s_piv = p;
        if (!(best > 0.0)) atomicExch(err, 3);
      }
    }
    __syncthreads();
    const int p = s_piv;
    const double inv = 1.0 / sM[p * ld + c];
    __syncthreads();
    for (int e = tid; e < w; e += T) {   // swap rows c and p, scale the pivot row
      const double tp = sM[p * ld + e], tc = sM[c * ld + e];
      if (p != c) sM[p * ld + e] = tc;
      sM[c * ld + e] = tp * inv;
    }
    __syncthreads();
    for (int r = tid; r < n; r += T) s_col[r] = (r == c) ? 0.0 : sM[r * ld + c];
    __syncthreads();
    for (int r = tid >> 5; r < n; r += (T >> 5)) {   // one warp per row: no index division
      const double f = s_col[r];
      if (f != 0.0)
        for (int k = tid & 31; k < w; k += 32) sM[r * ld + k] -= f * sM[c * ld + k];
    }
    __syncthreads();
  }
  for (int r = tid >> 5; r < n; r += (T >> 5))
    for (int k = tid & 31; k < n; k += 32) Ainv[r * n + k] = sM[r * ld + n + k];
}

__global__ void __launch_bounds__(256)
coarse_apply_kernel(const int n, const double* __restrict__ Ainv, const double* __restrict__ b,
                    double* __restrict__ x) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  double s = 0.0;
  for (int c = 0; c < n; ++c) s += Ainv[(size_t)r * n + c] * b[c];
  x[r] = s;
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
static inline int grid_for(long long n) { return (int)((n + kSB - 1) / kSB); }

// Row ranges: every launcher works on grid rows [r0, r1) (r1 < 0: the whole grid).  Arrays are
// always indexed globally, so a strip-partitioned rank simply passes its own rows.
static inline void row_range(int ny, int n1, int r0, int r1, int& vb, int& ve) {
  if (r1 < 0) { r0 = 0; r1 = ny + 1; }
  vb = r0 * n1;
  ve = r1 * n1;
}

int st_lump_extras(int nx, int ny, const double* QQ, const double* QQx, double* QQl, cudaStream_t st) {
  const int nv = (nx + 1) * (ny + 1);
  lump_extras_kernel<<<grid_for(nv), kSB, 0, st>>>(nv, QQ, QQx, QQl);
  CM_BYTES((double)nv * 20 * 8.0);
  CM_LAUNCH_CHECK();
  return CM_OK;
}

int st_extract_blocks(int nx, int ny, const int* nrowptr, const int* ncol, const double* vals,
                      double* PP, double* PQ, double* QP, double* QQ, double* sp, double* sq,
                      int* err, cudaStream_t st, int r0, int r1, double* QQx, int* has_extra,
                      int* wide_host) {
  const int nv = (nx + 1) * (ny + 1);
  int vb, ve;
  row_range(ny, nx + 1, r0, r1, vb, ve);
  // interior-facet couples present?  (one 4-byte read of the total entry count, once per Jacobian):
  // without them the 6 extra q-q diagonals are neither zero-filled here nor read anywhere (has_extra = 0)
  int nnz = 0;
  CM_CUDA(cudaMemcpyAsync(&nnz, nrowptr + nv, sizeof(int), cudaMemcpyDeviceToHost, st));
  CM_CUDA(cudaStreamSynchronize(st));
  const long long nnz7 = (long long)nv + 2LL * ((long long)nx * (ny + 1) + (long long)ny * (nx + 1) + (long long)nx * ny);
  extract_blocks_kernel<<<grid_for(ve - vb), kSB, 0, st>>>(nx, nv, nrowptr, ncol, (const double2*)vals,
                                                           PP, PQ, QP, QQ, sp, sq, err, vb, ve,
                                                           nnz != nnz7 ? QQx : nullptr, has_extra);
  if (wide_host) *wide_host = (nnz != nnz7 && QQx != nullptr) ? 1 : 0;
  CM_BYTES((double)(ve - vb) * (2 * 28 * 8.0 + 16.0));  // CSR values in, 4 stencil blocks + scalings out
  CM_LAUNCH_CHECK();
  return CM_OK;
}

int st_split_scale(int nv, const double* b, const double* sp, const double* sq, double* out,
                   cudaStream_t st, int vb, int ve) {
  if (ve < 0) { vb = 0; ve = nv; }
  split_scale_kernel<<<grid_for(ve - vb), kSB, 0, st>>>(nv, (const double2*)b, sp, sq, out, vb, ve);
  CM_BYTES((double)(ve - vb) * 48.0);
  CM_LAUNCH_CHECK();
  return CM_OK;
}

int st_interleave(int nv, const double* xs, double* x, cudaStream_t st, int vb, int ve) {
  if (ve < 0) { vb = 0; ve = nv; }
  interleave_kernel<<<grid_for(ve - vb), kSB, 0, st>>>(nv, xs, (double2*)x, vb, ve);
  CM_BYTES((double)(ve - vb) * 32.0);
  CM_LAUNCH_CHECK();
  return CM_OK;
}

int st_pq_spmv(int nx, int ny, const double* PP, const double* PQ, const double* QP,
               const double* QQ, const double* sp, const double* sq, const double* x, double* y,
               cudaStream_t st, int r0, int r1, const double* QQx, const int* has_extra, const Halo* H,
               double* zcopy, long long zstride, const double* jptr, int halo_channels, const double* xscale) {
  int vb, ve;
  row_range(ny, nx + 1, r0, r1, vb, ve);
  prof_begin(kProfSpmv, (double)(ve - vb) * 288.0, st);
  CM_BYTES((double)(ve - vb) * 288.0);
  pq_stencil_spmv_kernel<<<grid_for(ve - vb), kSB, 0, st>>>(nx, ny, PP, PQ, QP, QQ, sp, sq, x, y, vb, ve,
                                                            QQx, has_extra, H ? *H : Halo{}, zcopy, zstride, jptr, halo_channels,
                                                            xscale);
  if (zcopy) CM_BYTES((double)(ve - vb) * 16.0);
  prof_end(kProfSpmv, st);
  CM_LAUNCH_CHECK();
  return CM_OK;
}

int st_unscale(int nx, int ny, const double* r, const double* sp, const double* sq, double* out,
               cudaStream_t st, int r0, int r1) {
  int vb, ve;
  row_range(ny, nx + 1, r0, r1, vb, ve);
  unscale_kernel<<<grid_for(ve - vb), kSB, 0, st>>>((nx + 1) * (ny + 1), r, sp, sq, out, vb, ve);
  CM_BYTES((double)(ve - vb) * 48.0);
  CM_LAUNCH_CHECK();
  return CM_OK;
}

int st_unscale_dev(int nx, int ny, const double* rbase, long long stride, const double* jptr, const double* vscale,
                   const double* sp, const double* sq, double* out, cudaStream_t st, int r0, int r1) {
  int vb, ve;
  row_range(ny, nx + 1, r0, r1, vb, ve);
  unscale_dev_kernel<<<grid_for(ve - vb), kSB, 0, st>>>((nx + 1) * (ny + 1), rbase, stride, jptr, vscale, sp, sq, out,
                                                       vb, ve);
  CM_BYTES((double)(ve - vb) * 48.0);
  CM_LAUNCH_CHECK();
  return CM_OK;
}

int st_residual(int nx, int ny, const double* A, const double* b, const double* x, double* r,
                cudaStream_t st, int r0, int r1) {
  int vb, ve;
  row_range(ny, nx + 1, r0, r1, vb, ve);
  stencil_residual_kernel<<<grid_for(ve - vb), kSB, 0, st>>>(nx, ny, A, b, x, r, vb, ve);
  CM_BYTES((double)(ve - vb) * 80.0);
  CM_LAUNCH_CHECK();
  return CM_OK;
}

// coarse rows [R0, R1)
int st_restrict(int nxc, int nyc, const double* rf, double* rc, cudaStream_t st, int R0, int R1) {
  int vb, ve;
  row_range(nyc, nxc + 1, R0, R1, vb, ve);
  restrict_kernel<<<grid_for(ve - vb), kSB, 0, st>>>(nxc, nyc, rf, rc, vb, ve);
  CM_BYTES((double)(ve - vb) * (4 * 8.0 + 8.0));  // fine residual read once, coarse rhs written
  CM_LAUNCH_CHECK();
  return CM_OK;
}

// fine rows [r0, r1)
int st_prolong_add(int nxc, int nyc, const double* ec, double* xf, cudaStream_t st, int r0, int r1) {
  int vb, ve;
  row_range(2 * nyc, 2 * nxc + 1, r0, r1, vb, ve);
  prolong_add_kernel<<<grid_for(ve - vb), kSB, 0, st>>>(nxc, nyc, ec, xf, vb, ve);
  CM_BYTES((double)(ve - vb) * (16.0 + 2.0));  // fine x read+write, coarse correction read once
  CM_LAUNCH_CHECK();
  return CM_OK;
}

int st_rap(int nxc, int nyc, const double* Af, double* Ac, cudaStream_t st, int R0, int R1) {
  int vb, ve;
  row_range(nyc, nxc + 1, R0, R1, vb, ve);
  rap_kernel<<<grid_for(ve - vb), kSB, 0, st>>>(nxc, nyc, Af, Ac, vb, ve);
  CM_BYTES((double)(ve - vb) * (4 * 56.0 + 56.0));  // fine operator read once, coarse operator written
  CM_LAUNCH_CHECK();
  return CM_OK;
}

// opt-in shared-memory sizes of the setup kernels, once per process and outside any stream capture (cm_pqs_create)
int st_init_attributes() {
  static std::atomic<bool> done{false};
  if (done.load()) return CM_OK;
  CM_CUDA(cudaFuncSetAttribute(line_factor_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CM_CUDA(cudaFuncSetAttribute(coarse_invert_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
  done.store(true);
  return CM_OK;
}

int st_line_factor(int nx, int ny, const double* A, double* lf, double* iu, double* cu, int* err,
                   cudaStream_t st, int r0, int r1) {
  if (r1 < 0) { r0 = 0; r1 = ny + 1; }
  if (r1 <= r0) return CM_OK;
  const int n1 = nx + 1;
  const size_t smem = 3 * (size_t)n1 * sizeof(double);
  if (smem > 200 * 1024) {  // very long lines: serial Thomas per line (single-GPU only)
    line_factor_kernel<<<(ny + 1 + 127) / 128, 128, 0, st>>>(nx, ny, A, lf, iu, cu, err);
    CM_LAUNCH_CHECK();
    return CM_OK;
  }
  static bool attr_set = false;
  if (!attr_set) {
    CM_CUDA(cudaFuncSetAttribute(line_factor_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 200 * 1024));
    attr_set = true;
  }
  const int T = (n1 > 1024) ? 256 : ((n1 > 128) ? 128 : 32);
  const int E = (n1 + T - 1) / T;
  line_factor_scan_kernel<<<r1 - r0, T, smem, st>>>(nx, ny, E, A, lf, iu, cu, err, r0);
  CM_BYTES((double)(r1 - r0) * n1 * 48.0);
  CM_LAUNCH_CHECK();
  return CM_OK;
}

int st_zebra(int nx, int ny, int colour, const double* A, const double* lf, const double* iu,
             const double* cu, const double* b, double* x, int x_zero, cudaStream_t st, int r0,
             int r1, double* tmp, const double* Ax, const int* has_extra, bool prof) {
  if (r1 < 0) { r0 = 0; r1 = ny + 1; }
  const int n1 = nx + 1;
  const int jfirst = r0 + (((r0 & 1) == colour) ? 0 : 1);
  const int nlines = (r1 - jfirst + 1) / 2;
  if (nlines <= 0) return CM_OK;
  if (tmp != nullptr && (n1 >= 1024 || Ax != nullptr) && n1 <= 5 * 896 &&
      (5 * (size_t)n1 + 16) * sizeof(double) <= 200 * 1024) {
    CM_BYTES((double)nlines * n1 * (x_zero ? 40.0 : 88.0));
    if (prof) prof_begin(kProfZebraFine, (double)nlines * n1 * (x_zero ? 40.0 : 88.0), st);
    const double* rhs = b;
    if (!x_zero) {
      const long long nt = (long long)nlines * n1;
      offline_rhs_kernel<<<(int)((nt + kSB - 1) / kSB), kSB, 0, st>>>(nx, ny, jfirst, nlines, A, b, x, tmp, Ax,
                                                                      has_extra);
      ++g_launches;
      rhs = tmp;
    }
    const char* edbg = getenv("CM_ZEBRA_DEBUG");   // diagnostics (tools/time_kernels.py); read per call, nothing cached
    const int dbg2 = edbg ? atoi(edbg) : 0;
    const int grid = nlines < 148 ? nlines : 148;
    static bool attr4 = false;
    if (!attr4) {
      CM_CUDA(cudaFuncSetAttribute(line_solve_tma_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   200 * 1024));
      attr4 = true;
    }
    int T = ((n1 + 4) / 5 + 31) / 32 * 32;
    if (T < 64) T = 64;
    const int npad = (n1 + 3) & ~1;
    line_solve_tma_kernel<5><<<grid, T, (4 * (size_t)npad + n1 + 1) * sizeof(double), st>>>(
        nx, ny, jfirst, nlines, lf, iu, cu, rhs, x, dbg2);
    if (prof) prof_end(kProfZebraFine, st);
    CM_LAUNCH_CHECK();
    return CM_OK;
  }
  const int T = (n1 > 2048) ? 512 : ((n1 > 1024) ? 256 : ((n1 > 128) ? 128 : 32));
  int E = (n1 + T - 1) / T;
  if ((E & 1) == 0) ++E;  // odd stride: conflict-free 64-bit shared-memory accesses per half warp
  const size_t smem = 2 * (size_t)n1 * sizeof(double);
  static bool attr_set = false;
  if (!attr_set) {
    CM_CUDA(cudaFuncSetAttribute(zebra_line_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 200 * 1024));
    CM_CUDA(cudaFuncSetAttribute(zebra_line_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 200 * 1024));
    CM_CUDA(cudaFuncSetAttribute(zebra_line_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    CM_CUDA(cudaFuncSetAttribute(zebra_line_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    attr_set = true;
  }
  if (smem > 200 * 1024) {
    set_error("zebra line relaxation: line of %d unknowns exceeds shared memory", n1);
    return CM_E_UNSUPPORTED;
  }
  const char* edbg1 = getenv("CM_ZEBRA_DEBUG");
  const int dbg = edbg1 ? atoi(edbg1) : 0;
  // algorithmic bytes of this launch: b, l, 1/u, c/u, x per node (+ 4 stencil diagonals and the two
  // neighbouring lines when the current iterate is used)
  CM_BYTES((double)nlines * n1 * (x_zero ? 40.0 : 88.0));
  if (prof) prof_begin(kProfZebraFine, (double)nlines * n1 * (x_zero ? 40.0 : 88.0), st);
  if (x_zero)
    zebra_line_kernel<true><<<nlines, T, smem, st>>>(nx, ny, colour, E, A, lf, iu, cu, b, x, dbg, jfirst, r1);
  else
    zebra_line_kernel<false><<<nlines, T, smem, st>>>(nx, ny, colour, E, A, lf, iu, cu, b, x, dbg, jfirst, r1);
  if (prof) prof_end(kProfZebraFine, st);
  CM_LAUNCH_CHECK();
  return CM_OK;
}

int st_coarse_invert(int nx, int ny, const double* A, double* M, double* Ainv, int* err,
                     cudaStream_t st) {
  const int n = (nx + 1) * (ny + 1);
  if (n <= 112) {
    static std::atomic<bool> attr{false};
    const size_t smem = (size_t)n * (2 * n + 1) * sizeof(double);
    if (smem > 48 * 1024 && !attr.load()) {
      CM_CUDA(cudaFuncSetAttribute(coarse_invert_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
      attr.store(true);
    }
    coarse_invert_smem_kernel<<<1, 256, smem, st>>>(nx, ny, A, Ainv, err);
  } else {
    coarse_invert_kernel<<<1, 256, 0, st>>>(nx, ny, A, M, Ainv, err);
  }
  CM_LAUNCH_CHECK();
  return CM_OK;
}

int st_coarse_apply(int n, const double* Ainv, const double* b, double* x, cudaStream_t st) {
  coarse_apply_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, Ainv, b, x);
  CM_BYTES((double)n * n * 8.0);
  CM_LAUNCH_CHECK();
  return CM_OK;
}

}  // namespace cm
