// Row-owner ("gather") assembly kernels for unstructured P1 meshes.
//
// One thread owns one vertex row: it walks the vertex's incident cells in ascending cell index,
// recomputes geometry and quadrature in registers and accumulates the rows of its own dofs, so
// every CSR slot is written exactly once, in DOLFIN's serial summation order, without atomics.
// Replaces FFC tabulate_tensor + DOLFIN Assembler::assemble_cells [ext] for the forms declared at
// advection_diffusion_convergence.py:196-207 (P-Q primal) and
// SphericalAdvectionDiffusionReaction_convergence.py:92-102 (scalar).
#include "cm_common.cuh"

namespace cm {

constexpr int kRowBlock = 128;

// ------------------------------------------------------------------------------------------------
// P-Q cell integrals
// ------------------------------------------------------------------------------------------------
template <int NQ, int GKIND, bool QGEN, bool WANTJ>
__global__ void __launch_bounds__(kRowBlock)
pq_rows_kernel(const cm_pq_params P, const int nv, const int* __restrict__ cells,
               const double2* __restrict__ coords, const int* __restrict__ v2c_ptr,
               const int* __restrict__ v2c_ent, const int* __restrict__ v2c_slot,
               const int* __restrict__ nrowptr, const int* __restrict__ ncol,
               const double2* __restrict__ u, const double* __restrict__ p_old,
               const double2* __restrict__ load, const double* __restrict__ Kq,
               const double* __restrict__ cq, double* __restrict__ vals, double2* __restrict__ b,
               const int accumulate) {
  extern __shared__ double acc[];  // [4*maxdeg][kRowBlock], column = thread
  const int tid = threadIdx.x;
  const int v = blockIdx.x * kRowBlock + tid;
  if (v >= nv) return;
  const int rp = nrowptr[v];
  const int deg = nrowptr[v + 1] - rp;
  if (WANTJ)
    for (int k = 0; k < 4 * deg; ++k) acc[k * kRowBlock + tid] = 0.0;
  double Fp = 0.0, Fq = 0.0;
  const bool trans = P.inv_dt != 0.0;
  const double2 X0 = coords[v];
  const double2 U0 = u[v];
  const double po0 = trans ? p_old[v] : 0.0;

  const int e1 = v2c_ptr[v + 1];
  for (int e = v2c_ptr[v]; e < e1; ++e) {
    const int ent = v2c_ent[e];
    const int c = ent >> 2, i = ent & 3;
    const int ia = (i == 2) ? 0 : i + 1, ib = (i == 0) ? 2 : i - 1;
    const int va = cells[3 * c + ia], vb = cells[3 * c + ib];
    const double2 X1 = coords[va], X2 = coords[vb];
    const double2 U1 = u[va], U2 = u[vb];
    TriGeom g;
    g.x0 = X0.x; g.y0 = X0.y; g.x1 = X1.x; g.y1 = X1.y; g.x2 = X2.x; g.y2 = X2.y;
    g.build();
    const double p0 = U0.x, p1 = U1.x, p2 = U2.x, q0 = U0.y, q1 = U1.y, q2 = U2.y;
    const double t0 = (P.supg_stab != 0.0) ? 0.5 * P.supg_stab * g.hmax() * g.gx0 : 0.0;

    double S0 = 0, SG = 0, SGp0 = 0, SGp1 = 0, SGp2 = 0, SGq0 = 0, SGq1 = 0, SGq2 = 0;
    double Mm0 = 0, Mm1 = 0, Mm2 = 0;
    double A0 = 0, A10 = 0, A11 = 0, A12 = 0, A20 = 0, A21 = 0, A22 = 0;
    double A30 = 0, A31 = 0, A32 = 0, B10 = 0, B11 = 0, B12 = 0;
    double Rq = 0, Bq = 0;  // residual-only sums
    const double qx = q0 * g.gx0 + q1 * g.gx1 + q2 * g.gx2;
#pragma unroll
    for (int k = 0; k < NQ; ++k) {
      double l0, l1, l2, w;
      TriRule<NQ>::get(k, l0, l1, l2, w);
      const double x = l0 * g.x0 + l1 * g.x1 + l2 * g.x2;
      const double y = l0 * g.y0 + l1 * g.y1 + l2 * g.y2;
      const double t = kFourPi * x * y;
      const double a = w * g.adet * (t * t);
      const double p = l0 * p0 + l1 * p1 + l2 * p2;
      const double q = l0 * q0 + l1 * q1 + l2 * q2;
      const double x2 = x * x;
      double G, Gp, Gq;
      eval_G<GKIND>(P, x2, p, q, G, Gp, Gq);
      S0 += a;
      SG += a * G;
      if (WANTJ) {
        const double aGp = a * Gp, aGq = a * Gq;
        SGp0 += aGp * l0; SGp1 += aGp * l1; SGp2 += aGp * l2;
        SGq0 += aGq * l0; SGq1 += aGq * l1; SGq2 += aGq * l2;
      }
      if (trans) {
        const double al0 = a * l0;
        Mm0 += al0 * l0; Mm1 += al0 * l1; Mm2 += al0 * l2;
      }
      const double cw = a * (l0 + t0);
      if (WANTJ) {
        A0 += cw;
        const double cx2 = cw * x2;
        A20 += cx2 * l0; A21 += cx2 * l1; A22 += cx2 * l2;
        if (QGEN) {
          A10 += cw * l0; A11 += cw * l1; A12 += cw * l2;
          if (P.q_divr != 0.0) {
            const double cd = cw * (2.0 / x);
            A30 += cd * l0; A31 += cd * l1; A32 += cd * l2;
          }
          if (P.q_ibp != 0.0) { B10 += a * l0; B11 += a * l1; B12 += a * l2; }
        }
      } else {
        double r;
        if (QGEN) {
          r = P.q_react * q + P.q_adv * qx + x2 * (P.q_x2p * p + P.q_x2q * q);
          if (P.q_divr != 0.0) r -= P.q_divr * (2.0 / x) * q;
          Bq += a * q;
        } else {
          r = qx + x2 * p;
        }
        Rq += cw * r;
      }
    }
    const double px = p0 * g.gx0 + p1 * g.gx1 + p2 * g.gx2;
    const double py = p0 * g.gy0 + p1 * g.gy1 + p2 * g.gy2;
    const double dg = P.D2 * g.gx0;
    double fp = S0 * (dg * px + P.D1 * g.gy0 * py) + dg * SG;
    if (trans) {
      const double pa = p_old[va], pb = p_old[vb];
      fp += P.inv_dt * (Mm0 * (p0 - po0) + Mm1 * (p1 - pa) + Mm2 * (p2 - pb));
    }
    Fp += fp;
    if (WANTJ) {
      const double K0 = dg * g.gx0 + P.D1 * g.gy0 * g.gy0;
      const double K1 = dg * g.gx1 + P.D1 * g.gy0 * g.gy1;
      const double K2 = dg * g.gx2 + P.D1 * g.gy0 * g.gy2;
      double Jpp0 = S0 * K0 + dg * SGp0, Jpp1 = S0 * K1 + dg * SGp1, Jpp2 = S0 * K2 + dg * SGp2;
      if (trans) { Jpp0 += P.inv_dt * Mm0; Jpp1 += P.inv_dt * Mm1; Jpp2 += P.inv_dt * Mm2; }
      const double Jpq0 = dg * SGq0, Jpq1 = dg * SGq1, Jpq2 = dg * SGq2;
      double Jqp0, Jqp1, Jqp2, Jqq0, Jqq1, Jqq2;
      if (QGEN) {
        Jqp0 = P.q_x2p * A20; Jqp1 = P.q_x2p * A21; Jqp2 = P.q_x2p * A22;
        const double gi = P.q_ibp * g.gx0;
        Jqq0 = P.q_react * A10 + P.q_adv * g.gx0 * A0 + P.q_x2q * A20 - P.q_divr * A30 - gi * B10;
        Jqq1 = P.q_react * A11 + P.q_adv * g.gx1 * A0 + P.q_x2q * A21 - P.q_divr * A31 - gi * B11;
        Jqq2 = P.q_react * A12 + P.q_adv * g.gx2 * A0 + P.q_x2q * A22 - P.q_divr * A32 - gi * B12;
      } else {
        Jqp0 = A20; Jqp1 = A21; Jqp2 = A22;
        Jqq0 = g.gx0 * A0; Jqq1 = g.gx1 * A0; Jqq2 = g.gx2 * A0;
      }
      Fq += (Jqp0 * p0 + Jqp1 * p1 + Jqp2 * p2) + (Jqq0 * q0 + Jqq1 * q1 + Jqq2 * q2);
      const int sl = v2c_slot[e];
      const int k0 = sl & 255, k1 = (sl >> 8) & 255, k2 = (sl >> 16) & 255;
      double* a0 = acc + (4 * k0) * kRowBlock + tid;
      double* a1 = acc + (4 * k1) * kRowBlock + tid;
      double* a2 = acc + (4 * k2) * kRowBlock + tid;
      a0[0] += Jpp0; a0[kRowBlock] += Jpq0; a0[2 * kRowBlock] += Jqp0; a0[3 * kRowBlock] += Jqq0;
      a1[0] += Jpp1; a1[kRowBlock] += Jpq1; a1[2 * kRowBlock] += Jqp1; a1[3 * kRowBlock] += Jqq1;
      a2[0] += Jpp2; a2[kRowBlock] += Jpq2; a2[2 * kRowBlock] += Jqp2; a2[3 * kRowBlock] += Jqq2;
    } else {
      Fq += Rq - P.q_ibp * g.gx0 * Bq;
    }
  }
  // constant facet operator of the C0IP q-equation (assembled once per mesh)
  if (Kq != nullptr) {
    double s = cq ? cq[v] : 0.0;
    for (int k = 0; k < deg; ++k) {
      const double kv = Kq[rp + k];
      s += kv * u[ncol[rp + k]].y;
      if (WANTJ) acc[(4 * k + 3) * kRowBlock + tid] += kv;
    }
    Fq += s;
  }
  double2 out = make_double2(Fp, Fq);
  if (load != nullptr) { const double2 l = load[v]; out.x -= l.x; out.y -= l.y; }
  if (accumulate) { const double2 o = b[v]; out.x += o.x; out.y += o.y; }
  b[v] = out;
  if (WANTJ) {
    double2* prow = reinterpret_cast<double2*>(vals + 4 * (size_t)rp);
    double2* qrow = prow + deg;
    for (int k = 0; k < deg; ++k) {
      double2 pv = make_double2(acc[(4 * k) * kRowBlock + tid], acc[(4 * k + 1) * kRowBlock + tid]);
      double2 qv = make_double2(acc[(4 * k + 2) * kRowBlock + tid], acc[(4 * k + 3) * kRowBlock + tid]);
      if (accumulate) {
        const double2 o = prow[k], r = qrow[k];
        pv.x += o.x; pv.y += o.y; qv.x += r.x; qv.y += r.y;
      }
      prow[k] = pv;
      qrow[k] = qv;
    }
  }
}

template <int NQ, int GKIND, bool QGEN>
static int launch_pq_rows(const cm_pq_params& P, int nv, const int* cells, const double* coords,
                          const int* v2c_ptr, const int* v2c_ent, const int* v2c_slot,
                          const int* nrowptr, const int* ncol, const double* u, const double* p_old,
                          const double* load, const double* Kq, const double* cq, double* vals,
                          double* b, int accumulate, int maxdeg, cudaStream_t st) {
  const int grid = (nv + kRowBlock - 1) / kRowBlock;
  CM_BYTES((vals != nullptr ? 148.05 : 36.0) * 2.0 * (double)nv);  // ~2 cells per vertex
  if (vals != nullptr) {
    const size_t smem = (size_t)4 * maxdeg * kRowBlock * sizeof(double);
    auto kern = pq_rows_kernel<NQ, GKIND, QGEN, true>;
    CM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kRowBlock, smem, st>>>(P, nv, cells, (const double2*)coords, v2c_ptr, v2c_ent,
                                        v2c_slot, nrowptr, ncol, (const double2*)u, p_old,
                                        (const double2*)load, Kq, cq, vals, (double2*)b, accumulate);
  } else {
    pq_rows_kernel<NQ, GKIND, QGEN, false><<<grid, kRowBlock, 0, st>>>(
        P, nv, cells, (const double2*)coords, v2c_ptr, v2c_ent, v2c_slot, nrowptr, ncol,
        (const double2*)u, p_old, (const double2*)load, Kq, cq, vals, (double2*)b, accumulate);
  }
  CM_LAUNCH_CHECK();
  return CM_OK;
}

static bool q_is_plain(const cm_pq_params& P) {
  return P.q_react == 0.0 && P.q_adv == 1.0 && P.q_x2p == 1.0 && P.q_x2q == 0.0 &&
         P.q_divr == 0.0 && P.q_ibp == 0.0 && P.supg_stab == 0.0;
}

int assemble_pq_rows(const cm_pq_params& P, int nv, const int* cells, const double* coords,
                     const int* v2c_ptr, const int* v2c_ent, const int* v2c_slot,
                     const int* nrowptr, const int* ncol, const double* u, const double* p_old,
                     const double* load, const double* Kq, const double* cq, double* vals,
                     double* b, int accumulate, int maxdeg, cudaStream_t st) {
  const int nq = nq_of_degree(P.degree);
  const bool plain = q_is_plain(P);
#define CM_DISPATCH(NQ_, GK_)                                                                  \
  if (nq == NQ_ && P.gkind == GK_) {                                                           \
    return plain ? launch_pq_rows<NQ_, GK_, false>(P, nv, cells, coords, v2c_ptr, v2c_ent,     \
                                                   v2c_slot, nrowptr, ncol, u, p_old, load, Kq, \
                                                   cq, vals, b, accumulate, maxdeg, st)        \
                 : launch_pq_rows<NQ_, GK_, true>(P, nv, cells, coords, v2c_ptr, v2c_ent,      \
                                                  v2c_slot, nrowptr, ncol, u, p_old, load, Kq, \
                                                  cq, vals, b, accumulate, maxdeg, st);        \
  }
  CM_DISPATCH(6, CM_G_NONE)
  CM_DISPATCH(6, CM_G_POWER)
  CM_DISPATCH(6, CM_G_GSTAR)
  CM_DISPATCH(12, CM_G_NONE)
  CM_DISPATCH(12, CM_G_POWER)
  CM_DISPATCH(12, CM_G_GSTAR)
#undef CM_DISPATCH
  set_error("assemble_pq_cells: unsupported quadrature degree %d / gkind %d", P.degree, P.gkind);
  return CM_E_UNSUPPORTED;
}

// ------------------------------------------------------------------------------------------------
// scalar P1 forms (F-A / F-B)
// ------------------------------------------------------------------------------------------------
template <int NQ>
__global__ void __launch_bounds__(kRowBlock)
scalar_rows_kernel(const cm_scalar_params P, const int nv, const int* __restrict__ cells,
                   const double2* __restrict__ coords, const int* __restrict__ v2c_ptr,
                   const int* __restrict__ v2c_ent, const int* __restrict__ v2c_slot,
                   const int* __restrict__ nrowptr, const double* __restrict__ u,
                   const double* __restrict__ load, double* __restrict__ vals,
                   double* __restrict__ b) {
  extern __shared__ double acc[];  // [maxdeg][kRowBlock]
  const int tid = threadIdx.x;
  const int v = blockIdx.x * kRowBlock + tid;
  if (v >= nv) return;
  const int rp = nrowptr[v];
  const int deg = nrowptr[v + 1] - rp;
  const bool wantj = vals != nullptr;
  if (wantj)
    for (int k = 0; k < deg; ++k) acc[k * kRowBlock + tid] = 0.0;
  double F = 0.0;
  const double2 X0 = coords[v];
  const double u0 = u[v];
  const int e1 = v2c_ptr[v + 1];
  for (int e = v2c_ptr[v]; e < e1; ++e) {
    const int ent = v2c_ent[e];
    const int c = ent >> 2, i = ent & 3;
    const int ia = (i == 2) ? 0 : i + 1, ib = (i == 0) ? 2 : i - 1;
    const int va = cells[3 * c + ia], vb = cells[3 * c + ib];
    const double2 X1 = coords[va], X2 = coords[vb];
    const double u1 = u[va], u2 = u[vb];
    TriGeom g;
    g.x0 = X0.x; g.y0 = X0.y; g.x1 = X1.x; g.y1 = X1.y; g.x2 = X2.x; g.y2 = X2.y;
    g.build();
    const double ux = u0 * g.gx0 + u1 * g.gx1 + u2 * g.gx2;
    const double uy = u0 * g.gy0 + u1 * g.gy1 + u2 * g.gy2;
    double J0 = 0, J1 = 0, J2 = 0, f = 0;
    const double K0 = P.D2 * g.gx0 * g.gx0 + P.D1 * g.gy0 * g.gy0;
    const double K1 = P.D2 * g.gx1 * g.gx0 + P.D1 * g.gy1 * g.gy0;
    const double K2 = P.D2 * g.gx2 * g.gx0 + P.D1 * g.gy2 * g.gy0;
#pragma unroll
    for (int k = 0; k < NQ; ++k) {
      double l0, l1, l2, w;
      TriRule<NQ>::get(k, l0, l1, l2, w);
      const double x = l0 * g.x0 + l1 * g.x1 + l2 * g.x2;
      const double y = l0 * g.y0 + l1 * g.y1 + l2 * g.y2;
      const double a = w * g.adet;
      const double uq = l0 * u0 + l1 * u1 + l2 * u2;
      const double ix = 2.0 / x, iy = 2.0 / y;
      const double gq = P.gcoef * x * x * uq * uq;
      const double gp = 2.0 * P.gcoef * x * x * uq;
      // test-function factor for the reaction-like terms: phi_0; advective: d phi_0/dx
      const double cx = P.D2 * ix, cy = P.D1 * iy;
      f += a * (uq * l0 + (P.D2 * ux * g.gx0 + P.D1 * uy * g.gy0) - cx * ux * l0 - cy * uy * l0 +
                gq * g.gx0 - ix * gq * l0);
      if (wantj) {
        const double m = l0 - ix * gp * l0 + gp * g.gx0;  // multiplies phi_j
        J0 += a * (l0 * m + K0 - (cx * g.gx0 + cy * g.gy0) * l0);
        J1 += a * (l1 * m + K1 - (cx * g.gx1 + cy * g.gy1) * l0);
        J2 += a * (l2 * m + K2 - (cx * g.gx2 + cy * g.gy2) * l0);
      }
    }
    F += f;
    if (wantj) {
      const int sl = v2c_slot[e];
      acc[(sl & 255) * kRowBlock + tid] += J0;
      acc[((sl >> 8) & 255) * kRowBlock + tid] += J1;
      acc[((sl >> 16) & 255) * kRowBlock + tid] += J2;
    }
  }
  b[v] = F - (load ? load[v] : 0.0);
  if (wantj)
    for (int k = 0; k < deg; ++k) vals[rp + k] = acc[k * kRowBlock + tid];
}

int assemble_scalar_rows(const cm_scalar_params& P, int nv, const int* cells, const double* coords,
                         const int* v2c_ptr, const int* v2c_ent, const int* v2c_slot,
                         const int* nrowptr, const double* u, const double* load, double* vals,
                         double* b, int maxdeg, cudaStream_t st) {
  const int grid = (nv + kRowBlock - 1) / kRowBlock;
  const size_t smem = vals ? (size_t)maxdeg * kRowBlock * sizeof(double) : 0;
  const int nq = nq_of_degree(P.degree);
#define CM_SC(NQ_)                                                                            \
  if (nq == NQ_) {                                                                            \
    auto kern = scalar_rows_kernel<NQ_>;                                                      \
    CM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kern<<<grid, kRowBlock, smem, st>>>(P, nv, cells, (const double2*)coords, v2c_ptr, v2c_ent, \
                                        v2c_slot, nrowptr, u, load, vals, b);                 \
    CM_LAUNCH_CHECK();                                                                        \
    return CM_OK;                                                                             \
  }
  CM_SC(3)
  CM_SC(6)
  CM_SC(12)
#undef CM_SC
  return CM_E_UNSUPPORTED;
}

// ------------------------------------------------------------------------------------------------
// state-independent load vector from forcing values at the quadrature points
// ------------------------------------------------------------------------------------------------
template <int NQ>
__global__ void __launch_bounds__(kRowBlock)
load_rows_kernel(const int ncomp, const int weighted, const double supg_stab, const int nv,
                 const int* __restrict__ cells, const double2* __restrict__ coords,
                 const int* __restrict__ v2c_ptr, const int* __restrict__ v2c_ent,
                 const double* __restrict__ fq, double* __restrict__ load) {
  const int v = blockIdx.x * kRowBlock + threadIdx.x;
  if (v >= nv) return;
  double L0 = 0.0, L1 = 0.0;
  const int e1 = v2c_ptr[v + 1];
  for (int e = v2c_ptr[v]; e < e1; ++e) {
    const int ent = v2c_ent[e];
    const int c = ent >> 2, i = ent & 3;
    // canonical frame of the cell (fq is stored in the cell's own point order)
    const int a0 = cells[3 * c], a1 = cells[3 * c + 1], a2 = cells[3 * c + 2];
    const double2 X0 = coords[a0], X1 = coords[a1], X2 = coords[a2];
    TriGeom g;
    g.x0 = X0.x; g.y0 = X0.y; g.x1 = X1.x; g.y1 = X1.y; g.x2 = X2.x; g.y2 = X2.y;
    g.build();
    const double gxi = (i == 0) ? g.gx0 : (i == 1 ? g.gx1 : g.gx2);
    const double ti = (supg_stab != 0.0) ? 0.5 * supg_stab * g.hmax() * gxi : 0.0;
    const double* f = fq + (size_t)c * NQ * ncomp;
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int k = 0; k < NQ; ++k) {
      double l0, l1, l2, w;
      TriRule<NQ>::get(k, l0, l1, l2, w);
      double a = w * g.adet;
      if (weighted) {
        const double x = l0 * g.x0 + l1 * g.x1 + l2 * g.x2;
        const double y = l0 * g.y0 + l1 * g.y1 + l2 * g.y2;
        const double t = kFourPi * x * y;
        a *= t * t;
      }
      const double li = (i == 0) ? l0 : (i == 1 ? l1 : l2);
      s0 += a * f[k * ncomp] * li;
      if (ncomp == 2) s1 += a * f[k * ncomp + 1] * (li + ti);
    }
    L0 += s0;
    L1 += s1;
  }
  if (ncomp == 2) {
    reinterpret_cast<double2*>(load)[v] = make_double2(L0, L1);
  } else {
    load[v] = L0;
  }
}

int assemble_load(int ncomp, int weighted, double supg_stab, int degree, int nv, const int* cells,
                  const double* coords, const int* v2c_ptr, const int* v2c_ent, const double* fq,
                  double* load, cudaStream_t st) {
  const int grid = (nv + kRowBlock - 1) / kRowBlock;
  const int nq = nq_of_degree(degree);
#define CM_LD(NQ_)                                                                             \
  if (nq == NQ_) {                                                                             \
    load_rows_kernel<NQ_><<<grid, kRowBlock, 0, st>>>(ncomp, weighted, supg_stab, nv, cells,   \
                                                      (const double2*)coords, v2c_ptr, v2c_ent, \
                                                      fq, load);                               \
    CM_LAUNCH_CHECK();                                                                         \
    return CM_OK;                                                                              \
  }
  CM_LD(3)
  CM_LD(6)
  CM_LD(12)
#undef CM_LD
  return CM_E_UNSUPPORTED;
}

// ------------------------------------------------------------------------------------------------
// C0IP facet operator (state independent): interior jump penalty and exterior-facet terms
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int find_slot(const int* __restrict__ ncol, int rp, int deg, int w) {
  for (int k = 0; k < deg; ++k)
    if (ncol[rp + k] == w) return k;
  return -1;
}

__device__ __forceinline__ double facet_int_W(int degree, double2 A, double2 B) {
  double s[4], w[4];
  const int n = facet_rule(degree, s, w);
  double acc = 0.0;
  for (int k = 0; k < n; ++k) {
    const double x = A.x + s[k] * (B.x - A.x), y = A.y + s[k] * (B.y - A.y);
    const double t = kFourPi * x * y;
    acc += w[k] * t * t;
  }
  return acc;
}

__global__ void __launch_bounds__(kRowBlock)
c0ip_rows_kernel(const cm_pq_params P, const int nv, const int* __restrict__ cells,
                 const double2* __restrict__ coords, const int* __restrict__ fverts,
                 const int* __restrict__ fcell, const int* __restrict__ v2f_ptr,
                 const int* __restrict__ v2f_ent, const int* __restrict__ nrowptr,
                 const int* __restrict__ ncol, double* __restrict__ Kq) {
  const int v = blockIdx.x * kRowBlock + threadIdx.x;
  if (v >= nv) return;
  const int rp = nrowptr[v], deg = nrowptr[v + 1] - rp;
  const int e1 = v2f_ptr[v + 1];
  for (int e = v2f_ptr[v]; e < e1; ++e) {
    const int f = v2f_ent[e];
    const int c0 = fcell[2 * f], c1 = fcell[2 * f + 1];
    const int fa = fverts[2 * f], fb = fverts[2 * f + 1];
    const double2 A = coords[fa], B = coords[fb];
    const double tx = B.x - A.x, ty = B.y - A.y;
    const double L = sqrt(tx * tx + ty * ty);
    double nx = ty / L, ny = -tx / L;
    int mv[6];
    double d[6];
    double hsum = 0.0;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const int c = s ? c1 : c0;
      const int a0 = cells[3 * c], a1 = cells[3 * c + 1], a2 = cells[3 * c + 2];
      const double2 X0 = coords[a0], X1 = coords[a1], X2 = coords[a2];
      TriGeom g;
      g.x0 = X0.x; g.y0 = X0.y; g.x1 = X1.x; g.y1 = X1.y; g.x2 = X2.x; g.y2 = X2.y;
      g.build();
      if (s == 0) {
        // orient n outward from cell0: the vertex opposite the facet lies inside
        const int opp = (a0 != fa && a0 != fb) ? 0 : ((a1 != fa && a1 != fb) ? 1 : 2);
        const double2 O = opp == 0 ? X0 : (opp == 1 ? X1 : X2);
        if ((O.x - A.x) * nx + (O.y - A.y) * ny > 0.0) { nx = -nx; ny = -ny; }
      }
      const double sg = s ? -1.0 : 1.0;
      mv[3 * s] = a0; mv[3 * s + 1] = a1; mv[3 * s + 2] = a2;
      d[3 * s] = sg * (g.gx0 * nx + g.gy0 * ny);
      d[3 * s + 1] = sg * (g.gx1 * nx + g.gy1 * ny);
      d[3 * s + 2] = sg * (g.gx2 * nx + g.gy2 * ny);
      hsum += g.hmax();
    }
    const double hbar = (P.c0ip_h == 1) ? L : 0.5 * hsum;
    const double coef = P.c0ip_gamma * hbar * hbar * facet_int_W(P.degree, A, B) * L;
    for (int i = 0; i < 6; ++i) {
      if (mv[i] != v) continue;
      for (int j = 0; j < 6; ++j) {
        const int k = find_slot(ncol, rp, deg, mv[j]);
        Kq[rp + k] += coef * d[i] * d[j];
      }
    }
  }
}

__global__ void __launch_bounds__(kRowBlock)
ext_rows_kernel(const cm_pq_params P, const int nv, const double2* __restrict__ coords,
                const int* __restrict__ efverts, const int* __restrict__ efopp,
                const int* __restrict__ efkind, const double* __restrict__ qdata,
                const int* __restrict__ v2e_ptr, const int* __restrict__ v2e_ent,
                const int* __restrict__ nrowptr, const int* __restrict__ ncol,
                double* __restrict__ Kq, double* __restrict__ cq) {
  const int v = blockIdx.x * kRowBlock + threadIdx.x;
  if (v >= nv) return;
  const int rp = nrowptr[v], deg = nrowptr[v + 1] - rp;
  const int e1 = v2e_ptr[v + 1];
  double cv = 0.0;
  for (int e = v2e_ptr[v]; e < e1; ++e) {
    const int f = v2e_ent[e];
    const int fa = efverts[2 * f], fb = efverts[2 * f + 1];
    const double2 A = coords[fa], B = coords[fb], O = coords[efopp[f]];
    const double tx = B.x - A.x, ty = B.y - A.y;
    const double L = sqrt(tx * tx + ty * ty);
    double nx = ty / L, ny = -tx / L;
    if ((O.x - A.x) * nx + (O.y - A.y) * ny > 0.0) { nx = -nx; ny = -ny; }
    const double bn = nx * P.r2x + ny * P.r2y;
    const double ng = 0.5 * (fabs(bn) - bn);
    const int kind = efkind[f];
    const double pen = (kind & CM_EF_PENALTY) ? P.pen_stab2 * (L * ng * ng + 8.0 / L) : 0.0;
    const double kadv = (kind & CM_EF_ADV) ? bn : 0.0;
    const double kdat = (kind & CM_EF_DATA) ? bn : 0.0;
    double s[4], w[4];
    const int n = facet_rule(P.degree, s, w);
    const int ka = find_slot(ncol, rp, deg, fa), kb = find_slot(ncol, rp, deg, fb);
    double Ja = 0.0, Jb = 0.0;
    for (int k = 0; k < n; ++k) {
      const double x = A.x + s[k] * tx, y = A.y + s[k] * ty;
      const double t = kFourPi * x * y;
      const double W = t * t;
      const double ds = w[k] * L;
      const double pa = 1.0 - s[k], pb = s[k];
      const double pv = (v == fa) ? pa : pb;
      const double cqq = kadv * W + pen;
      const double cd = kdat * W - pen;
      Ja += ds * cqq * pv * pa;
      Jb += ds * cqq * pv * pb;
      cv += ds * cd * (qdata ? qdata[(size_t)f * n + k] : 0.0) * pv;
    }
    Kq[rp + ka] += Ja;
    Kq[rp + kb] += Jb;
  }
  if (e1 > v2e_ptr[v]) cq[v] += cv;
}

int assemble_c0ip(const cm_pq_params& P, int nv, const int* cells, const double* coords,
                  const int* fverts, const int* fcell, const int* v2f_ptr, const int* v2f_ent,
                  const int* nrowptr, const int* ncol, double* Kq, cudaStream_t st) {
  const int grid = (nv + kRowBlock - 1) / kRowBlock;
  c0ip_rows_kernel<<<grid, kRowBlock, 0, st>>>(P, nv, cells, (const double2*)coords, fverts, fcell,
                                               v2f_ptr, v2f_ent, nrowptr, ncol, Kq);
  CM_LAUNCH_CHECK();
  return CM_OK;
}

int assemble_ext(const cm_pq_params& P, int nv, const double* coords, const int* efverts,
                 const int* efopp, const int* efkind, const double* qdata, const int* v2e_ptr,
                 const int* v2e_ent, const int* nrowptr, const int* ncol, double* Kq, double* cq,
                 cudaStream_t st) {
  const int grid = (nv + kRowBlock - 1) / kRowBlock;
  ext_rows_kernel<<<grid, kRowBlock, 0, st>>>(P, nv, (const double2*)coords, efverts, efopp, efkind,
                                              qdata, v2e_ptr, v2e_ent, nrowptr, ncol, Kq, cq);
  CM_LAUNCH_CHECK();
  return CM_OK;
}

// ------------------------------------------------------------------------------------------------
// Dirichlet rows: A[row,:]=0, A[row,row]=1, b[row]=x[row]-g   (DirichletBC::apply [ext], B.7)
// ------------------------------------------------------------------------------------------------
__global__ void dirichlet_kernel(const int bs, const int nrows, const int* __restrict__ rows,
                                 const int* __restrict__ diag_slot, const double* __restrict__ g,
                                 const double* __restrict__ x, const int* __restrict__ nrowptr,
                                 double* __restrict__ vals, double* __restrict__ b) {
  // one warp per BC row
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= nrows) return;
  const int dof = rows[r];
  const int node = (bs == 2) ? (dof >> 1) : dof;
  const int comp = (bs == 2) ? (dof & 1) : 0;
  const int rp = nrowptr[node], deg = nrowptr[node + 1] - rp;
  if (vals != nullptr) {
    const size_t base = (size_t)bs * bs * rp + (size_t)comp * bs * deg;
    const int len = bs * deg;
    const int dpos = bs * diag_slot[r] + comp;
    for (int k = lane; k < len; k += 32) vals[base + k] = (k == dpos) ? 1.0 : 0.0;
  }
  if (lane == 0) b[dof] = x[dof] - g[r];
}

// the same on the stencil form of the P-Q Jacobian (cm_pq_blocks): identity row, unit row scaling
__global__ void dirichlet_blocks_kernel(const int nrows, const int* __restrict__ rows, const size_t nv,
                                        const cm_pq_blocks blk) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  const int dof = rows[r];
  const size_t v = (size_t)(dof >> 1);
  double* __restrict__ B0 = (dof & 1) ? blk.QP : blk.PP;   // coupling to p
  double* __restrict__ B1 = (dof & 1) ? blk.QQ : blk.PQ;   // coupling to q
#pragma unroll
  for (int s = 0; s < 7; ++s) {
    B0[(size_t)s * nv + v] = (!(dof & 1) && s == 3) ? 1.0 : 0.0;
    B1[(size_t)s * nv + v] = ((dof & 1) && s == 3) ? 1.0 : 0.0;
  }
  ((dof & 1) ? blk.sq : blk.sp)[v] = 1.0;
}

int apply_dirichlet_blocks(int nrows, const int* rows, int nv, const cm_pq_blocks& blk, cudaStream_t st) {
  if (nrows == 0) return CM_OK;
  dirichlet_blocks_kernel<<<(nrows + 127) / 128, 128, 0, st>>>(nrows, rows, (size_t)nv, blk);
  CM_LAUNCH_CHECK();
  return CM_OK;
}

int apply_dirichlet(int bs, int nrows, const int* rows, const int* diag_slot, const double* g,
                    const double* x, const int* nrowptr, double* vals, double* b, cudaStream_t st) {
  if (nrows == 0) return CM_OK;
  const int threads = 128;
  const int grid = (nrows * 32 + threads - 1) / threads;
  dirichlet_kernel<<<grid, threads, 0, st>>>(bs, nrows, rows, diag_slot, g, x, nrowptr, vals, b);
  CM_LAUNCH_CHECK();
  return CM_OK;
}

}  // namespace cm
