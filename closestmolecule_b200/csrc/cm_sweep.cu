// Structured-mesh assembly of the P-Q cell integrals ("sweep" kernel).
//
// On RectangleMesh-numbered meshes (create_flat_boundary.py:65; vertex coordinates arbitrary, so
// boundary-fitted production-like grids qualify) cells, neighbours and CSR slots follow from
// (ix,iy): no connectivity, gather table or cell->nnz map is read.  A CTA owns a strip of TX-1
// vertex columns and sweeps upwards over a segment of vertex rows.  One thread per CELL computes the
// cell's quadrature sums ONCE (registers) and adds its three element rows into shared-memory
// accumulators of the two vertex rows the quad row touches, in three barrier-separated phases with a
// fixed summation order per CSR slot -- deterministic, no atomics.  A finished vertex row is written
// out exactly once in the DOLFIN/PETSc-AIJ value layout.
// Replaces the same reference code as pq_rows_kernel (advection_diffusion_convergence.py:196-207).
#include "cm_common.cuh"
#include "cm_decl.cuh"

namespace cm {

constexpr int kTX = 128;       // threads per CTA = quads per strip row; the strip owns kTX-1 vertices
constexpr int kAccPerVertex = 30;  // 7 stencil slots x (pp,pq,qp,qq) + (Fp,Fq)

struct CellSums {
  double S0, SG, SGp[3], SGq[3];
  double M1[6], M2[6], Md[6];  // symmetric 3x3: (00,01,02,11,12,22)
};

__device__ __forceinline__ int sym(int i, int j) {
  // index into the packed symmetric 3x3
  return (i <= j) ? (i == 0 ? j : (i == 1 ? 2 + j : 5)) : (j == 0 ? i : (j == 1 ? 2 + i : 5));
}

template <int NQ, int GKIND, bool QGEN>
__device__ __forceinline__ void cell_sums(const cm_pq_params& P, const TriGeom& g, const double p0,
                                          const double p1, const double p2, const double q0,
                                          const double q1, const double q2, const bool trans,
                                          CellSums& S) {
  S.S0 = S.SG = 0.0;
#pragma unroll
  for (int i = 0; i < 3; ++i) S.SGp[i] = S.SGq[i] = 0.0;
#pragma unroll
  for (int i = 0; i < 6; ++i) S.M1[i] = S.M2[i] = S.Md[i] = 0.0;
  const bool need_md = QGEN && P.q_divr != 0.0;
#pragma unroll
  for (int k = 0; k < NQ; ++k) {
    double l[3], w;
    TriRule<NQ>::get(k, l[0], l[1], l[2], w);
    const double x = l[0] * g.x0 + l[1] * g.x1 + l[2] * g.x2;
    const double y = l[0] * g.y0 + l[1] * g.y1 + l[2] * g.y2;
    const double t = kFourPi * x * y;
    const double a = w * g.adet * (t * t);
    const double p = l[0] * p0 + l[1] * p1 + l[2] * p2;
    const double q = l[0] * q0 + l[1] * q1 + l[2] * q2;
    const double x2 = x * x;
    double G, Gp, Gq;
    eval_G<GKIND>(P, x2, p, q, G, Gp, Gq);
    S.S0 += a;
    S.SG += a * G;
    const double aGp = a * Gp, aGq = a * Gq, ax2 = a * x2;
    const double ad = need_md ? a * (2.0 / x) : 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      S.SGp[i] += aGp * l[i];
      S.SGq[i] += aGq * l[i];
      const double ali = a * l[i], axi = ax2 * l[i], adi = ad * l[i];
#pragma unroll
      for (int j = i; j < 3; ++j) {
        const int s = (i == 0) ? j : (i == 1 ? 2 + j : 5);
        S.M1[s] += ali * l[j];
        S.M2[s] += axi * l[j];
        if (need_md) S.Md[s] += adi * l[j];
      }
    }
  }
  (void)trans;
}

// Residual only (the convergence test of every Newton iteration): the three element rows of F straight
// from the quadrature points, without the mass-type matrices the Jacobian needs:
//   Fp_I = S0 (D2 gx_I p_x + D1 gy_I p_y) + D2 gx_I SG + inv_dt sum_k a dp_k l_I
//   Fq_I = sum_k a rho_k (l_I + tau gx_I) - q_ibp gx_I sum_k a q_k,
//   rho = q_react q + q_adv q_x + x^2 (q_x2p p + q_x2q q) - q_divr (2/x) q.
template <int NQ, int GKIND, bool QGEN>
__device__ __forceinline__ void cell_resid(const cm_pq_params& P, const TriGeom& g, const double* pe,
                                           const double* qe, const double* po, const bool trans,
                                           const double tau, double* Fp, double* Fq) {
  const double gx[3] = {g.gx0, g.gx1, g.gx2}, gy[3] = {g.gy0, g.gy1, g.gy2};
  const double px = pe[0] * gx[0] + pe[1] * gx[1] + pe[2] * gx[2];
  const double py = pe[0] * gy[0] + pe[1] * gy[1] + pe[2] * gy[2];
  const double qx = qe[0] * gx[0] + qe[1] * gx[1] + qe[2] * gx[2];
  const double dp[3] = {pe[0] - po[0], pe[1] - po[1], pe[2] - po[2]};
  double S0 = 0.0, SG = 0.0, R0 = 0.0, Q0 = 0.0;
  double Rq[3] = {0.0, 0.0, 0.0}, T[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int k = 0; k < NQ; ++k) {
    double l[3], w;
    TriRule<NQ>::get(k, l[0], l[1], l[2], w);
    const double x = l[0] * g.x0 + l[1] * g.x1 + l[2] * g.x2;
    const double y = l[0] * g.y0 + l[1] * g.y1 + l[2] * g.y2;
    const double t = kFourPi * x * y;
    const double a = w * g.adet * (t * t);
    const double p = l[0] * pe[0] + l[1] * pe[1] + l[2] * pe[2];
    const double q = l[0] * qe[0] + l[1] * qe[1] + l[2] * qe[2];
    const double x2 = x * x;
    double G, Gp, Gq;
    eval_G<GKIND>(P, x2, p, q, G, Gp, Gq);
    S0 += a;
    SG += a * G;
    double rho;
    if (QGEN) {
      rho = P.q_react * q + P.q_adv * qx + x2 * (P.q_x2p * p + P.q_x2q * q);
      if (P.q_divr != 0.0) rho -= P.q_divr * (2.0 / x) * q;
      if (P.q_ibp != 0.0) Q0 += a * q;
    } else {
      rho = qx + x2 * p;
    }
    const double ar = a * rho;
    R0 += ar;
#pragma unroll
    for (int i = 0; i < 3; ++i) Rq[i] += ar * l[i];
    if (trans) {
      const double ad = a * (l[0] * dp[0] + l[1] * dp[1] + l[2] * dp[2]);
#pragma unroll
      for (int i = 0; i < 3; ++i) T[i] += ad * l[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double dg = P.D2 * gx[i];
    Fp[i] = S0 * (dg * px + P.D1 * gy[i] * py) + dg * SG;
    if (trans) Fp[i] += P.inv_dt * T[i];
    Fq[i] = Rq[i];
    if (QGEN) Fq[i] += tau * gx[i] * R0 - P.q_ibp * gx[i] * Q0;
  }
}

// element rows of local vertex I: 3 column blocks (pp,pq,qp,qq) + (Fp,Fq)
template <int I, bool QGEN>
__device__ __forceinline__ void cell_row(const cm_pq_params& P, const TriGeom& g, const CellSums& S,
                                         const double* pe, const double* qe, const double* po,
                                         const bool trans, const double tau, double (*J)[4],
                                         double& Fp, double& Fq) {
  const double gx[3] = {g.gx0, g.gx1, g.gx2}, gy[3] = {g.gy0, g.gy1, g.gy2};
  const double px = pe[0] * gx[0] + pe[1] * gx[1] + pe[2] * gx[2];
  const double py = pe[0] * gy[0] + pe[1] * gy[1] + pe[2] * gy[2];
  const double dg = P.D2 * gx[I];
  const double ti = QGEN ? tau * gx[I] : 0.0;
  double S1[3], Mx[3], Mdx[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    S1[j] = S.M1[sym(0, j)] + S.M1[sym(1, j)] + S.M1[sym(2, j)];
    Mx[j] = S.M2[sym(0, j)] + S.M2[sym(1, j)] + S.M2[sym(2, j)];
    Mdx[j] = QGEN ? S.Md[sym(0, j)] + S.Md[sym(1, j)] + S.Md[sym(2, j)] : 0.0;
  }
  Fp = S.S0 * (dg * px + P.D1 * gy[I] * py) + dg * S.SG;
  Fq = 0.0;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const double K = dg * gx[j] + P.D1 * gy[I] * gy[j];
    double jpp = S.S0 * K + dg * S.SGp[j];
    if (trans) {
      jpp += P.inv_dt * S.M1[sym(I, j)];
      Fp += P.inv_dt * S.M1[sym(I, j)] * (pe[j] - po[j]);
    }
    const double jpq = dg * S.SGq[j];
    double jqp, jqq;
    if (QGEN) {
      const double m2 = S.M2[sym(I, j)] + ti * Mx[j];
      jqp = P.q_x2p * m2;
      jqq = P.q_react * (S.M1[sym(I, j)] + ti * S1[j]) + P.q_adv * gx[j] * (S1[I] + ti * S.S0) +
            P.q_x2q * m2 - P.q_divr * (S.Md[sym(I, j)] + ti * Mdx[j]) - P.q_ibp * gx[I] * S1[j];
    } else {
      jqp = S.M2[sym(I, j)];
      jqq = gx[j] * S1[I];
    }
    J[j][0] = jpp; J[j][1] = jpq; J[j][2] = jqp; J[j][3] = jqq;
    Fq += jqp * pe[j] + jqq * qe[j];
  }
}

// acc layout: [row buffer][entry 0..29][vertex column 0..kTX-1]
__device__ __forceinline__ void acc_add(double* acc, int col, int slot, const double* J4) {
  double* a = acc + (4 * slot) * kTX + col;
  a[0] += J4[0]; a[kTX] += J4[1]; a[2 * kTX] += J4[2]; a[3 * kTX] += J4[3];
}

template <int I, bool QGEN, bool WANTJ>
__device__ __forceinline__ void add_row(const cm_pq_params& P, const TriGeom& g, const CellSums& S,
                                        const double* pe, const double* qe, const double* po,
                                        const bool trans, const double tau, double* acc, const int col,
                                        const int s0, const int s1, const int s2, const double* Fpr,
                                        const double* Fqr) {
  if (WANTJ) {
    double J[3][4], Fp, Fq;
    cell_row<I, QGEN>(P, g, S, pe, qe, po, trans, tau, J, Fp, Fq);
    acc_add(acc, col, s0, J[0]);
    acc_add(acc, col, s1, J[1]);
    acc_add(acc, col, s2, J[2]);
    acc[28 * kTX + col] += Fp;
    acc[29 * kTX + col] += Fq;
  } else {  // residual only: the rows were computed by cell_resid
    acc[28 * kTX + col] += Fpr[I];
    acc[29 * kTX + col] += Fqr[I];
  }
}

// 2*kTX threads: thread t < kTX owns cell A = (v0,v1,v3) of quad (ix0-1+t, jq), thread kTX+t owns
// cell B = (v0,v2,v3) of the same quad.  Three barrier-separated phases add the three element rows
// of every cell; in each phase the A-threads and the B-threads target different row buffers or
// different columns, so there are no write conflicts and no atomics, and the summation order per
// CSR slot is fixed (B1,B2,A2 from the quad row below, then A1,A0,B0): bit-reproducible.
template <int NQ, int GKIND, bool QGEN, bool WANTJ>
__global__ void __launch_bounds__(2 * kTX, 2)
pq_sweep_kernel(const cm_pq_params P, const int nx, const int ny, const int rows_per_seg,
                const int row_begin, const int row_end,
                const double2* __restrict__ coords, const int* __restrict__ nrowptr,
                const double2* __restrict__ u, const double* __restrict__ p_old,
                const double2* __restrict__ load, double* __restrict__ vals,
                double2* __restrict__ b, const cm_pq_blocks blk) {
  extern __shared__ double sm[];  // 2 row buffers x kAccPerVertex x kTX
  const int tid = threadIdx.x;
  const bool isB = tid >= kTX;
  const int t = isB ? tid - kTX : tid;
  const int n1 = nx + 1;
  const int ix0 = blockIdx.x * (kTX - 1);          // first owned vertex column
  const int jy0 = row_begin + blockIdx.y * rows_per_seg;  // first owned vertex row
  const int jy1 = min(jy0 + rows_per_seg, row_end);       // one past the last owned vertex row
  const int qx = ix0 - 1 + t;                      // this thread's quad column
  const bool quad_ok = qx >= 0 && qx < nx;
  const int colL = t - 1, colR = t;                // accumulator columns of the quad's left/right vertices
  const bool ownL = quad_ok && colL >= 0;          // left vertex (qx) is owned by this strip
  const bool ownR = quad_ok && colR <= kTX - 2;    // right vertex (qx+1)
  const bool trans = P.inv_dt != 0.0;
  const int vx = ix0 + t;                          // vertex written out by this thread (p-row: A half, q-row: B half)
  const bool vown = t <= kTX - 2 && vx <= nx;

  // zero both row buffers once; afterwards the write-out zeroes what it reads
  for (int e = tid; e < 2 * kAccPerVertex * kTX; e += 2 * kTX) sm[e] = 0.0;

  for (int jq = jy0 - 1; jq < jy1; ++jq) {
    double* accLo = sm + ((jq & 1) * kAccPerVertex) * kTX;        // vertex row jq
    double* accUp = sm + (((jq + 1) & 1) * kAccPerVertex) * kTX;  // vertex row jq+1
    const bool row_ok = jq >= 0 && jq < ny;                       // quad row exists
    const bool vrow_lo = jq >= jy0 && jq < jy1;                   // vertex row jq exists and is owned
    const bool do_lo = row_ok && vrow_lo;
    const bool do_up = row_ok && jq + 1 < jy1;                    // vertex row jq+1 is owned
    // pull the vertex row the NEXT step adds (jq+2) towards L1 while this step computes
    if (quad_ok && jq + 2 <= ny) {
      const int vn = (jq + 2) * n1 + qx + (isB ? 0 : 1);
      asm volatile("prefetch.global.L1 [%0];" ::"l"(coords + vn));
      asm volatile("prefetch.global.L1 [%0];" ::"l"(u + vn));
    }
    TriGeom g;
    CellSums S;
    double tau = 0.0;
    double pe[3], qe[3], po[3] = {0.0, 0.0, 0.0};
    double Fpr[3], Fqr[3];  // residual-only pass: element rows of F
    const bool active = quad_ok && row_ok;
    if (active) {
      const int v0 = jq * n1 + qx;
      const int vm = isB ? v0 + n1 : v0 + 1;  // middle vertex: v1 (A) or v2 (B)
      const int v3 = v0 + n1 + 1;
      const double2 X0 = coords[v0], X1 = coords[vm], X2 = coords[v3];
      const double2 U0 = u[v0], U1 = u[vm], U2 = u[v3];
      g.x0 = X0.x; g.y0 = X0.y; g.x1 = X1.x; g.y1 = X1.y; g.x2 = X2.x; g.y2 = X2.y;
      g.build();
      pe[0] = U0.x; pe[1] = U1.x; pe[2] = U2.x;
      qe[0] = U0.y; qe[1] = U1.y; qe[2] = U2.y;
      if (trans) { po[0] = p_old[v0]; po[1] = p_old[vm]; po[2] = p_old[v3]; }
      if (QGEN && P.supg_stab != 0.0) tau = 0.5 * P.supg_stab * g.hmax();
      if (WANTJ)
        cell_sums<NQ, GKIND, QGEN>(P, g, pe[0], pe[1], pe[2], qe[0], qe[1], qe[2], trans, S);
      else
        cell_resid<NQ, GKIND, QGEN>(P, g, pe, qe, po, trans, tau, Fpr, Fqr);
    }
    __syncthreads();  // previous write-out (which re-zeroed its buffer) done
    // ---- phase 1: A row 1 -> (row jq, right vertex); B row 1 -> (row jq+1, left vertex) ----------
    if (active) {
      if (!isB) {
        if (do_lo && ownR) add_row<1, QGEN, WANTJ>(P, g, S, pe, qe, po, trans, tau, accLo, colR, 2, 3, 5, Fpr, Fqr);
      } else {
        if (do_up && ownL) add_row<1, QGEN, WANTJ>(P, g, S, pe, qe, po, trans, tau, accUp, colL, 1, 3, 4, Fpr, Fqr);
      }
    }
    __syncthreads();
    // ---- phase 2: A row 0 -> (row jq, left vertex); B row 2 -> (row jq+1, right vertex) ----------
    if (active) {
      if (!isB) {
        if (do_lo && ownL) add_row<0, QGEN, WANTJ>(P, g, S, pe, qe, po, trans, tau, accLo, colL, 3, 4, 6, Fpr, Fqr);
      } else {
        if (do_up && ownR) add_row<2, QGEN, WANTJ>(P, g, S, pe, qe, po, trans, tau, accUp, colR, 0, 2, 3, Fpr, Fqr);
      }
    }
    __syncthreads();
    // ---- phase 3: A row 2 -> (row jq+1, right vertex); B row 0 -> (row jq, left vertex) ----------
    if (active) {
      if (!isB) {
        if (do_up && ownR) add_row<2, QGEN, WANTJ>(P, g, S, pe, qe, po, trans, tau, accUp, colR, 0, 1, 3, Fpr, Fqr);
      } else {
        if (do_lo && ownL) add_row<0, QGEN, WANTJ>(P, g, S, pe, qe, po, trans, tau, accLo, colL, 3, 5, 6, Fpr, Fqr);
      }
    }
    __syncthreads();
    // vertex row jq is complete: write it out (every CSR slot exactly once); the A half of the CTA
    // writes the p-rows (and b), the B half the q-rows
    if (vrow_lo && vown) {
      const int v = jq * n1 + vx;
      const int rp = nrowptr[v];
      const int deg = nrowptr[v + 1] - rp;
      if (!isB) {
        double2 out = make_double2(accLo[28 * kTX + t], accLo[29 * kTX + t]);
        accLo[28 * kTX + t] = 0.0;
        accLo[29 * kTX + t] = 0.0;
        if (load != nullptr) { const double2 l = load[v]; out.x -= l.x; out.y -= l.y; }
        b[v] = out;
      }
      if (WANTJ) {
        double2* row = reinterpret_cast<double2*>(vals + 4 * (size_t)rp) + (isB ? deg : 0);
        const int e0 = isB ? 2 : 0;
        int k = 0;
        // optional second copy of the row in the stencil form of the linear solver (cm_pq_blocks: diagonals
        // [7][nv], slot s = this loop's s) + the row equilibration 1 / max |entry|: what cm_pqs_ctx_setup would
        // otherwise extract from `vals` in a pass of its own
        const bool wblk = blk.PP != nullptr;
        const size_t nvb = (size_t)n1 * (ny + 1);
        double* __restrict__ B0 = isB ? blk.QP : blk.PP;
        double* __restrict__ B1 = isB ? blk.QQ : blk.PQ;
        double rmax = 0.0;
#pragma unroll
        for (int s = 0; s < 7; ++s) {
          const int dx = (s == 0 || s == 2) ? -1 : ((s == 4 || s == 6) ? 1 : 0);
          const int dy = (s < 2) ? -1 : (s > 4 ? 1 : 0);
          const int jx = vx + dx, jy = jq + dy;
          if (jx < 0 || jx > nx || jy < 0 || jy > ny) {
            if (wblk) { B0[(size_t)s * nvb + v] = 0.0; B1[(size_t)s * nvb + v] = 0.0; }
            continue;
          }
          const double a0 = accLo[(4 * s + e0) * kTX + t], a1 = accLo[(4 * s + e0 + 1) * kTX + t];
          row[k] = make_double2(a0, a1);
          if (wblk) {
            B0[(size_t)s * nvb + v] = a0;
            B1[(size_t)s * nvb + v] = a1;
            rmax = fmax(rmax, fmax(fabs(a0), fabs(a1)));
          }
          accLo[(4 * s + e0) * kTX + t] = 0.0;
          accLo[(4 * s + e0 + 1) * kTX + t] = 0.0;
          ++k;
        }
        if (wblk) (isB ? blk.sq : blk.sp)[v] = rmax > 0.0 ? 1.0 / rmax : 1.0;
      }
    }
  }
}

template <int NQ, int GKIND, bool QGEN>
static int launch_sweep(const cm_pq_params& P, int nx, int ny, const double* coords,
                        const int* nrowptr, const double* u, const double* p_old,
                        const double* load, double* vals, double* b, cudaStream_t st, int row_begin,
                        int row_end, const cm_pq_blocks& blk) {
  if (row_end < 0) { row_begin = 0; row_end = ny + 1; }
  const int nrows = row_end - row_begin;
  if (nrows <= 0) return CM_OK;
  const int strips = (nx + 1 + kTX - 2) / (kTX - 1);
  // y-segments of at least 32 rows each (every segment recomputes the quad row below its first
  // vertex row); their number is chosen so that the grid fills a whole number of waves of
  // 148 SMs x 2 resident CTAs as exactly as possible (no nearly empty last wave)
  int rows = 32, segs = (nrows + rows - 1) / rows;
  {
    const int slots = 148 * 2;
    double best = 1e300;
    int best_rows = rows;
    for (int r = 32; r <= 96 && r <= (nrows > 32 ? nrows : 32); ++r) {
      const int sg = (nrows + r - 1) / r;
      const long long ctas = (long long)sg * strips;
      const long long waves = (ctas + slots - 1) / slots;
      const double cost = (double)waves * (r + 1);  // time ~ waves x (rows + redundant start row)
      if (cost < best) { best = cost; best_rows = r; }
    }
    rows = best_rows;
    segs = (nrows + rows - 1) / rows;
  }
  const size_t smem = (size_t)2 * kAccPerVertex * kTX * sizeof(double);
  dim3 grid(strips, segs);
  // algorithmic bytes: 148.05 B per cell of the assembled rows
  // J+F: 148.05 B/cell; F only: coordinates + state + residual = 12 + 16 + 16 + 16 B per vertex-ish = 36 B/cell
  CM_BYTES((vals != nullptr ? 148.05 : 36.0) * 2.0 * nx * (double)nrows);
  if (vals != nullptr) prof_begin(kProfAssembly, 148.05 * 2.0 * nx * (double)nrows, st);
  if (vals != nullptr) {
    auto kern = pq_sweep_kernel<NQ, GKIND, QGEN, true>;
    CM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 2 * kTX, smem, st>>>(P, nx, ny, rows, row_begin, row_end, (const double2*)coords, nrowptr, (const double2*)u,
                                  p_old, (const double2*)load, vals, (double2*)b, blk);
    if (blk.PP != nullptr) CM_BYTES((double)nrows * (nx + 1) * (28 * 8.0 + 16.0));   // stencil blocks + scalings
    prof_end(kProfAssembly, st);
  } else {
    auto kern = pq_sweep_kernel<NQ, GKIND, QGEN, false>;
    CM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 2 * kTX, smem, st>>>(P, nx, ny, rows, row_begin, row_end, (const double2*)coords, nrowptr, (const double2*)u,
                                  p_old, (const double2*)load, vals, (double2*)b, cm_pq_blocks{});
  }
  CM_LAUNCH_CHECK();
  return CM_OK;
}

int assemble_pq_sweep(const cm_pq_params& P, int nx, int ny, const double* coords,
                      const int* nrowptr, const double* u, const double* p_old, const double* load,
                      double* vals, double* b, cudaStream_t st, int row_begin, int row_end,
                      const cm_pq_blocks* blocks) {
  const int nq = nq_of_degree(P.degree);
  const cm_pq_blocks blk = (blocks != nullptr && vals != nullptr) ? *blocks : cm_pq_blocks{};
  const bool plain = P.q_react == 0.0 && P.q_adv == 1.0 && P.q_x2p == 1.0 && P.q_x2q == 0.0 &&
                     P.q_divr == 0.0 && P.q_ibp == 0.0 && P.supg_stab == 0.0;
#define CM_SW(NQ_, GK_)                                                                          \
  if (nq == NQ_ && P.gkind == GK_)                                                               \
    return plain ? launch_sweep<NQ_, GK_, false>(P, nx, ny, coords, nrowptr, u, p_old, load, vals, b, st, row_begin, row_end, blk) \
                 : launch_sweep<NQ_, GK_, true>(P, nx, ny, coords, nrowptr, u, p_old, load, vals, b, st, row_begin, row_end, blk);
  CM_SW(6, CM_G_NONE)
  CM_SW(6, CM_G_POWER)
  CM_SW(6, CM_G_GSTAR)
  CM_SW(12, CM_G_NONE)
  CM_SW(12, CM_G_POWER)
  CM_SW(12, CM_G_GSTAR)
#undef CM_SW
  set_error("assemble_pq_cells_structured: unsupported quadrature degree %d / gkind %d", P.degree,
            P.gkind);
  return CM_E_UNSUPPORTED;
}

}  // namespace cm
