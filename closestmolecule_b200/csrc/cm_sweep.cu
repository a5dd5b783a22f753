// Structured-mesh assembly of the P-Q cell integrals ("sweep" kernel).
//
// On RectangleMesh-numbered meshes (create_flat_boundary.py:65; vertex coordinates arbitrary, so
// boundary-fitted production-like grids qualify) cells, neighbours and CSR slots follow from
// (ix,iy): no connectivity, gather table or cell->nnz map is read.  A CTA owns a strip of TX-1
// vertex columns and sweeps upwards over a segment of vertex rows.  Thread t owns quad
// (ix0-1+t, jq): it computes the quadrature sums of the quad's two cells ONCE (registers) and adds
// the element rows into shared-memory accumulators of the two vertex rows the quad touches, in four
// barrier-separated phases ordered so that every CSR slot receives its contributions in ascending
// cell index -- the summation order of DOLFIN's serial assembler -- without atomics.  A finished
// vertex row is written out exactly once in the DOLFIN/PETSc-AIJ value layout.
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

template <int NQ, int GKIND, bool QGEN, bool WANTJ>
__global__ void __launch_bounds__(kTX)
pq_sweep_kernel(const cm_pq_params P, const int nx, const int ny, const int rows_per_seg,
                const double2* __restrict__ coords, const int* __restrict__ nrowptr,
                const double2* __restrict__ u, const double* __restrict__ p_old,
                const double2* __restrict__ load, double* __restrict__ vals,
                double2* __restrict__ b) {
  extern __shared__ double sm[];  // 2 row buffers x kAccPerVertex x kTX
  const int t = threadIdx.x;
  const int n1 = nx + 1;
  const int ix0 = blockIdx.x * (kTX - 1);          // first owned vertex column
  const int jy0 = blockIdx.y * rows_per_seg;       // first owned vertex row
  const int jy1 = min(jy0 + rows_per_seg, ny + 1); // one past the last owned vertex row
  const int qx = ix0 - 1 + t;                      // this thread's quad column
  const bool quad_ok = qx >= 0 && qx < nx;
  const int colL = t - 1, colR = t;                // accumulator columns of the quad's left/right vertices
  const bool ownL = quad_ok && colL >= 0;          // left vertex (qx) is owned by this strip
  const bool ownR = quad_ok && colR <= kTX - 2;    // right vertex (qx+1)
  const bool trans = P.inv_dt != 0.0;
  // owned vertex of this thread for the write-out
  const int vx = ix0 + t;
  const bool vown = t <= kTX - 2 && vx <= nx;

  // lower vertex row of the first quad row
  double2 X0, X1, U0, U1;
  double O0 = 0.0, O1 = 0.0;
  const int jq_first = max(jy0 - 1, 0);
  if (quad_ok) {
    const int v0 = jq_first * n1 + qx;
    X0 = coords[v0]; X1 = coords[v0 + 1];
    U0 = u[v0]; U1 = u[v0 + 1];
    if (trans) { O0 = p_old[v0]; O1 = p_old[v0 + 1]; }
  }
  // zero the buffer of the first owned vertex row
  for (int e = 0; e < kAccPerVertex; ++e) sm[((jy0 & 1) * kAccPerVertex + e) * kTX + t] = 0.0;
  __syncthreads();

  for (int jq = jy0 - 1; jq < jy1; ++jq) {
    double* accLo = sm + ((jq & 1) * kAccPerVertex) * kTX;        // vertex row jq
    double* accUp = sm + (((jq + 1) & 1) * kAccPerVertex) * kTX;  // vertex row jq+1
    const bool row_ok = jq >= 0 && jq < ny;                       // quad row exists
    const bool vrow_lo = jq >= jy0 && jq < jy1;                   // vertex row jq exists and is owned
    const bool do_lo = row_ok && vrow_lo;
    const bool do_up = row_ok && jq + 1 < jy1;                    // vertex row jq+1 is owned
    if (jq >= jy0 - 1) {
      // fresh accumulators for the upper vertex row (the lower one carries the previous step)
      if (jq + 1 != jy0)
        for (int e = 0; e < kAccPerVertex; ++e) accUp[e * kTX + t] = 0.0;
    }
    double2 X2, X3, U2, U3;
    double O2 = 0.0, O3 = 0.0;
    TriGeom gA, gB;
    CellSums SA, SB;
    double tauA = 0.0, tauB = 0.0;
    double pA[3], qA[3], oA[3], pB[3], qB[3], oB[3];
    const bool active = quad_ok && row_ok;
    if (active) {
      const int v2 = (jq + 1) * n1 + qx;
      X2 = coords[v2]; X3 = coords[v2 + 1];
      U2 = u[v2]; U3 = u[v2 + 1];
      if (trans) { O2 = p_old[v2]; O3 = p_old[v2 + 1]; }
      // cell A = (v0, v1, v3), cell B = (v0, v2, v3)
      gA.x0 = X0.x; gA.y0 = X0.y; gA.x1 = X1.x; gA.y1 = X1.y; gA.x2 = X3.x; gA.y2 = X3.y; gA.build();
      gB.x0 = X0.x; gB.y0 = X0.y; gB.x1 = X2.x; gB.y1 = X2.y; gB.x2 = X3.x; gB.y2 = X3.y; gB.build();
      pA[0] = U0.x; pA[1] = U1.x; pA[2] = U3.x; qA[0] = U0.y; qA[1] = U1.y; qA[2] = U3.y;
      pB[0] = U0.x; pB[1] = U2.x; pB[2] = U3.x; qB[0] = U0.y; qB[1] = U2.y; qB[2] = U3.y;
      oA[0] = O0; oA[1] = O1; oA[2] = O3; oB[0] = O0; oB[1] = O2; oB[2] = O3;
      cell_sums<NQ, GKIND, QGEN>(P, gA, pA[0], pA[1], pA[2], qA[0], qA[1], qA[2], trans, SA);
      cell_sums<NQ, GKIND, QGEN>(P, gB, pB[0], pB[1], pB[2], qB[0], qB[1], qB[2], trans, SB);
      if (QGEN && P.supg_stab != 0.0) {
        tauA = 0.5 * P.supg_stab * gA.hmax();
        tauB = 0.5 * P.supg_stab * gB.hmax();
      }
    }
    __syncthreads();  // zeroing done
    double J[3][4], Fp, Fq;
    // ---- phase 1: lower-right vertex v1 (row jq): cell A local 1 ------------------------------
    if (active && do_lo && ownR) {
      cell_row<1, QGEN>(P, gA, SA, pA, qA, oA, trans, tauA, J, Fp, Fq);
      if (WANTJ) { acc_add(accLo, colR, 2, J[0]); acc_add(accLo, colR, 3, J[1]); acc_add(accLo, colR, 5, J[2]); }
      accLo[28 * kTX + colR] += Fp; accLo[29 * kTX + colR] += Fq;
    }
    __syncthreads();
    // ---- phase 2: lower-left vertex v0 (row jq): cells A then B, local 0 -----------------------
    if (active && do_lo && ownL) {
      cell_row<0, QGEN>(P, gA, SA, pA, qA, oA, trans, tauA, J, Fp, Fq);
      if (WANTJ) { acc_add(accLo, colL, 3, J[0]); acc_add(accLo, colL, 4, J[1]); acc_add(accLo, colL, 6, J[2]); }
      accLo[28 * kTX + colL] += Fp; accLo[29 * kTX + colL] += Fq;
      cell_row<0, QGEN>(P, gB, SB, pB, qB, oB, trans, tauB, J, Fp, Fq);
      if (WANTJ) { acc_add(accLo, colL, 3, J[0]); acc_add(accLo, colL, 5, J[1]); acc_add(accLo, colL, 6, J[2]); }
      accLo[28 * kTX + colL] += Fp; accLo[29 * kTX + colL] += Fq;
    }
    __syncthreads();
    // vertex row jq is complete: write it out (every CSR slot exactly once)
    if (vrow_lo && vown) {
      const int v = jq * n1 + vx;
      const int rp = nrowptr[v];
      const int deg = nrowptr[v + 1] - rp;
      double2 out = make_double2(accLo[28 * kTX + t], accLo[29 * kTX + t]);
      if (load != nullptr) { const double2 l = load[v]; out.x -= l.x; out.y -= l.y; }
      b[v] = out;
      if (WANTJ) {
        double2* prow = reinterpret_cast<double2*>(vals + 4 * (size_t)rp);
        double2* qrow = prow + deg;
        int k = 0;
#pragma unroll
        for (int s = 0; s < 7; ++s) {
          const int dx = (s == 0 || s == 2) ? -1 : ((s == 4 || s == 6) ? 1 : 0);
          const int dy = (s < 2) ? -1 : (s > 4 ? 1 : 0);
          const int jx = vx + dx, jy = jq + dy;
          if (jx < 0 || jx > nx || jy < 0 || jy > ny) continue;
          prow[k] = make_double2(accLo[(4 * s) * kTX + t], accLo[(4 * s + 1) * kTX + t]);
          qrow[k] = make_double2(accLo[(4 * s + 2) * kTX + t], accLo[(4 * s + 3) * kTX + t]);
          ++k;
        }
      }
    }
    // ---- phase 3: upper-right vertex v3 (row jq+1): cells A then B, local 2 -------------------
    if (active && do_up && ownR) {
      cell_row<2, QGEN>(P, gA, SA, pA, qA, oA, trans, tauA, J, Fp, Fq);
      if (WANTJ) { acc_add(accUp, colR, 0, J[0]); acc_add(accUp, colR, 1, J[1]); acc_add(accUp, colR, 3, J[2]); }
      accUp[28 * kTX + colR] += Fp; accUp[29 * kTX + colR] += Fq;
      cell_row<2, QGEN>(P, gB, SB, pB, qB, oB, trans, tauB, J, Fp, Fq);
      if (WANTJ) { acc_add(accUp, colR, 0, J[0]); acc_add(accUp, colR, 2, J[1]); acc_add(accUp, colR, 3, J[2]); }
      accUp[28 * kTX + colR] += Fp; accUp[29 * kTX + colR] += Fq;
    }
    __syncthreads();
    // ---- phase 4: upper-left vertex v2 (row jq+1): cell B local 1 -----------------------------
    if (active && do_up && ownL) {
      cell_row<1, QGEN>(P, gB, SB, pB, qB, oB, trans, tauB, J, Fp, Fq);
      if (WANTJ) { acc_add(accUp, colL, 1, J[0]); acc_add(accUp, colL, 3, J[1]); acc_add(accUp, colL, 4, J[2]); }
      accUp[28 * kTX + colL] += Fp; accUp[29 * kTX + colL] += Fq;
    }
    __syncthreads();
    if (active) { X0 = X2; X1 = X3; U0 = U2; U1 = U3; O0 = O2; O1 = O3; }
  }
}

template <int NQ, int GKIND, bool QGEN>
static int launch_sweep(const cm_pq_params& P, int nx, int ny, const double* coords,
                        const int* nrowptr, const double* u, const double* p_old,
                        const double* load, double* vals, double* b, cudaStream_t st) {
  const int strips = (nx + 1 + kTX - 2) / (kTX - 1);
  // enough y-segments to fill the machine several times over, at least 32 rows each
  int segs = (148 * 8 + strips - 1) / strips;
  int rows = (ny + 1 + segs - 1) / segs;
  if (rows < 32) rows = 32;
  segs = (ny + 1 + rows - 1) / rows;
  const size_t smem = (size_t)2 * kAccPerVertex * kTX * sizeof(double);
  dim3 grid(strips, segs);
  if (vals != nullptr) {
    auto kern = pq_sweep_kernel<NQ, GKIND, QGEN, true>;
    CM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kTX, smem, st>>>(P, nx, ny, rows, (const double2*)coords, nrowptr, (const double2*)u,
                                  p_old, (const double2*)load, vals, (double2*)b);
  } else {
    auto kern = pq_sweep_kernel<NQ, GKIND, QGEN, false>;
    CM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kTX, smem, st>>>(P, nx, ny, rows, (const double2*)coords, nrowptr, (const double2*)u,
                                  p_old, (const double2*)load, vals, (double2*)b);
  }
  CM_LAUNCH_CHECK();
  return CM_OK;
}

int assemble_pq_sweep(const cm_pq_params& P, int nx, int ny, const double* coords,
                      const int* nrowptr, const double* u, const double* p_old, const double* load,
                      double* vals, double* b, cudaStream_t st) {
  const int nq = nq_of_degree(P.degree);
  const bool plain = P.q_react == 0.0 && P.q_adv == 1.0 && P.q_x2p == 1.0 && P.q_x2q == 0.0 &&
                     P.q_divr == 0.0 && P.q_ibp == 0.0 && P.supg_stab == 0.0;
#define CM_SW(NQ_, GK_)                                                                          \
  if (nq == NQ_ && P.gkind == GK_)                                                               \
    return plain ? launch_sweep<NQ_, GK_, false>(P, nx, ny, coords, nrowptr, u, p_old, load, vals, b, st) \
                 : launch_sweep<NQ_, GK_, true>(P, nx, ny, coords, nrowptr, u, p_old, load, vals, b, st);
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
