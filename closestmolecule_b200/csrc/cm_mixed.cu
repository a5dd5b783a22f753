// Lowest-order mixed-primal system RT1 x DG0 x CG1 (deg = 0) of convergence/test_for_paper_convergence.py:68,108-167
// (SURVEY section 8f, row 1): element tensors of the cell integrals (:137-141, 150-151), the deterministic
// CSR gather that replaces DOLFIN's serial add_local, and the weighted error functionals (:175-177).
// Unknowns per cell: 3 edge fluxes (RT basis sigma_i (x - X_i)/(2|K|) on the facet opposite local vertex i),
// 1 cell value of p, 3 vertex values of q.  The q-q facet terms (:142-145,152-153) come from the same
// kernels as the P-Q path (cm_assemble_c0ip_facets / cm_assemble_exterior_facets).
#include "cm_common.cuh"
#include "cm_decl.cuh"

namespace cm {

constexpr int kMxThreads = 128;
constexpr int kMxRule = 40;
constexpr int kMxBlocks = 592;

struct MxRule {
  int nq;
  double l1[kMxRule], l2[kMxRule], w[kMxRule];
};

// Je: [49][ncells] (entry-major, coalesced over cells), Fe: [7][ncells]
__global__ void __launch_bounds__(kMxThreads)
mixed_m0_cells_kernel(const MxRule R, const double D1, const double D2, const double geps,
                      const int ncells, const int* __restrict__ cells,
                      const double2* __restrict__ coords, const int* __restrict__ cell_edges,
                      const double* __restrict__ sign, const int nE, const double* __restrict__ u,
                      const double* __restrict__ f2q, const double* __restrict__ f3q,
                      double* __restrict__ Je, double* __restrict__ Fe) {
  const int c = blockIdx.x * kMxThreads + threadIdx.x;
  if (c >= ncells) return;
  double X[3], Y[3], se[3], sg[3], qe[3];
  int v[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    v[i] = cells[3 * (size_t)c + i];
    const double2 P = coords[v[i]];
    X[i] = P.x;
    Y[i] = P.y;
    sg[i] = sign[3 * (size_t)c + i];
    se[i] = u[cell_edges[3 * (size_t)c + i]];
  }
  const double pc = u[nE + c];
#pragma unroll
  for (int i = 0; i < 3; ++i) qe[i] = u[nE + ncells + v[i]];
  const double j00 = X[1] - X[0], j01 = X[2] - X[0], j10 = Y[1] - Y[0], j11 = Y[2] - Y[0];
  const double det = j00 * j11 - j01 * j10;
  const double adet = fabs(det), area = 0.5 * adet, id = 1.0 / det;
  double gx[3];
  gx[1] = j11 * id;
  gx[2] = -j10 * id;
  gx[0] = -(gx[1] + gx[2]);
  double rs[3], dv[3], dRx[3];  // RT scale sign/(2|K|), divergence sign/|K|, d/dx of the x component
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    rs[i] = sg[i] / (2.0 * area);
    dv[i] = sg[i] / area;
    dRx[i] = rs[i];
  }
  double F[7], J[49];
#pragma unroll
  for (int i = 0; i < 7; ++i) F[i] = 0.0;
#pragma unroll
  for (int i = 0; i < 49; ++i) J[i] = 0.0;
  for (int k = 0; k < R.nq; ++k) {
    const double l1 = R.l1[k], l2 = R.l2[k], l0 = 1.0 - l1 - l2;
    const double phi[3] = {l0, l1, l2};
    const double x = l0 * X[0] + l1 * X[1] + l2 * X[2];
    const double y = l0 * Y[0] + l1 * Y[1] + l2 * Y[2];
    const double t = kFourPi * x * y;
    const double a = R.w[k] * adet * (t * t);
    double Rx[3], Ry[3], dr[3], drD[3];
    double sx = 0.0, sy = 0.0, drDs = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      Rx[i] = (x - X[i]) * rs[i];
      Ry[i] = (y - Y[i]) * rs[i];
      dr[i] = dv[i] + 2.0 * Rx[i] / x + 2.0 * Ry[i] / y;
      drD[i] = (D2 + D1) * dRx[i] + 2.0 * D2 * Rx[i] / x + 2.0 * D1 * Ry[i] / y;
      sx += se[i] * Rx[i];
      sy += se[i] * Ry[i];
      drDs += drD[i] * se[i];
    }
    const double q = l0 * qe[0] + l1 * qe[1] + l2 * qe[2];
    const double x2 = x * x;
    const double iq = 1.0 / (q + geps);
    const double tg = x2 * pc * iq;
    const double G = tg * pc, Gp = 2.0 * tg, Gq = -G * iq;
    const double f2 = f2q[(size_t)c * R.nq + k], f3 = f3q[(size_t)c * R.nq + k];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      F[i] += a * (sx * Rx[i] + sy * Ry[i] + G * Rx[i] - pc * dr[i]);
      F[4 + i] += a * ((-(2.0 / x) * q + x2 * pc - f3) * phi[i] - q * gx[i]);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        J[7 * i + j] += a * (Rx[i] * Rx[j] + Ry[i] * Ry[j]);
        J[7 * i + 4 + j] += a * Gq * Rx[i] * phi[j];
        J[7 * (4 + i) + 4 + j] -= a * ((2.0 / x) * phi[i] * phi[j] + phi[j] * gx[i]);
      }
      J[7 * i + 3] += a * (Gp * Rx[i] - dr[i]);
      J[7 * 3 + i] -= a * drD[i];
      J[7 * (4 + i) + 3] += a * x2 * phi[i];
    }
    F[3] += a * (-drDs + f2);
  }
#pragma unroll
  for (int i = 0; i < 7; ++i) Fe[(size_t)i * ncells + c] = F[i];
  if (Je != nullptr) {
#pragma unroll
    for (int i = 0; i < 49; ++i) Je[(size_t)i * ncells + c] = J[i];
  }
}

// out[s] (+)= sum_{k in [ptr[s], ptr[s+1])} src[idx[k]] in list order (ascending cell = DOLFIN's serial
// insertion order): the precomputed cell->nnz map turned into a per-slot gather list; no atomics.
__global__ void __launch_bounds__(256)
csr_gather_kernel(const int nslots, const int* __restrict__ ptr, const int* __restrict__ idx,
                  const double* __restrict__ src, double* __restrict__ out, const int accumulate) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nslots) return;
  double acc = 0.0;
  for (int k = ptr[s]; k < ptr[s + 1]; ++k) acc += src[idx[k]];
  out[s] = accumulate ? out[s] + acc : acc;
}

// dst[map[k]] += src[k]  (map is injective)
__global__ void __launch_bounds__(256)
add_mapped_kernel(const int n, const int* __restrict__ map, const double* __restrict__ src,
                  double* __restrict__ dst) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) dst[map[k]] += src[k];
}

// error functionals of :175-177: sum a [(s_ex - s_h)^2 + (div_rad s_ex - div_rad s_h)^2], sum a (p_ex - p_h)^2,
// sum a [(q_ex - q_h)^2 + (dq_ex/dx - dq_h/dx)^2]; exact data exq[ncells*nq*6] = (sx, sy, drs, p, q, qx)
__global__ void __launch_bounds__(kMxThreads)
mixed_m0_errors_kernel(const MxRule R, const int ncells, const int* __restrict__ cells,
                       const double2* __restrict__ coords, const int* __restrict__ cell_edges,
                       const double* __restrict__ sign, const int nE, const double* __restrict__ u,
                       const double* __restrict__ exq, double* __restrict__ partial) {
  double e0 = 0.0, e1 = 0.0, e2 = 0.0;
  for (int c = blockIdx.x * kMxThreads + threadIdx.x; c < ncells; c += kMxBlocks * kMxThreads) {
    double X[3], Y[3], se[3], rs[3], dv[3], qe[3];
    int v[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      v[i] = cells[3 * (size_t)c + i];
      const double2 P = coords[v[i]];
      X[i] = P.x;
      Y[i] = P.y;
      se[i] = u[cell_edges[3 * (size_t)c + i]];
    }
    const double pc = u[nE + c];
#pragma unroll
    for (int i = 0; i < 3; ++i) qe[i] = u[nE + ncells + v[i]];
    const double j00 = X[1] - X[0], j01 = X[2] - X[0], j10 = Y[1] - Y[0], j11 = Y[2] - Y[0];
    const double det = j00 * j11 - j01 * j10;
    const double adet = fabs(det), area = 0.5 * adet, id = 1.0 / det;
    const double g1 = j11 * id, g2 = -j10 * id;
    const double qhx = -(g1 + g2) * qe[0] + g1 * qe[1] + g2 * qe[2];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const double sg = sign[3 * (size_t)c + i];
      rs[i] = sg / (2.0 * area);
      dv[i] = sg / area;
    }
    double c0 = 0.0, c1 = 0.0, c2 = 0.0;
    for (int k = 0; k < R.nq; ++k) {
      const double l1 = R.l1[k], l2 = R.l2[k], l0 = 1.0 - l1 - l2;
      const double x = l0 * X[0] + l1 * X[1] + l2 * X[2];
      const double y = l0 * Y[0] + l1 * Y[1] + l2 * Y[2];
      const double t = kFourPi * x * y;
      const double a = R.w[k] * adet * (t * t);
      double sx = 0.0, sy = 0.0, drs = 0.0;
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const double Rx = (x - X[i]) * rs[i], Ry = (y - Y[i]) * rs[i];
        sx += se[i] * Rx;
        sy += se[i] * Ry;
        drs += se[i] * (dv[i] + 2.0 * Rx / x + 2.0 * Ry / y);
      }
      const double* ex = exq + ((size_t)c * R.nq + k) * 6;
      const double qh = l0 * qe[0] + l1 * qe[1] + l2 * qe[2];
      const double d0 = ex[0] - sx, d1 = ex[1] - sy, d2 = ex[2] - drs, dp = ex[3] - pc,
                   dq = ex[4] - qh, dqx = ex[5] - qhx;
      c0 += a * (d0 * d0 + d1 * d1 + d2 * d2);
      c1 += a * (dp * dp);
      c2 += a * (dq * dq + dqx * dqx);
    }
    e0 += c0;
    e1 += c1;
    e2 += c2;
  }
  __shared__ double sh[3][kMxThreads];
  sh[0][threadIdx.x] = e0;
  sh[1][threadIdx.x] = e1;
  sh[2][threadIdx.x] = e2;
  __syncthreads();
  for (int o = kMxThreads / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o)
      for (int m = 0; m < 3; ++m) sh[m][threadIdx.x] += sh[m][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0)
    for (int m = 0; m < 3; ++m) partial[3 * blockIdx.x + m] = sh[m][0];
}

__global__ void __launch_bounds__(32)
mixed_final_sum_kernel(const int nparts, const double* __restrict__ partial, double* __restrict__ out) {
  if (threadIdx.x < 3) {
    double s = 0.0;
    for (int i = 0; i < nparts; ++i) s += partial[3 * i + threadIdx.x];
    out[threadIdx.x] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// k = 1: RT2 x DG1 x CG2 (test_for_paper_convergence.py with deg = 1; recorded table :211-217).
// Bases: RT2 = { lambda_j N_i, lambda_k N_i : facet i opposite local vertex i with end points j < k }
// + { lambda_1 N_1, lambda_2 N_2 } (unsigned) with the RT0 functions N_i = sigma_i (x - X_i)/(2|K|);
// DG1 = barycentric coordinates; CG2 = lambda_i (2 lambda_i - 1), 4 lambda_j lambda_k.  17 local dofs
// [s:8 | p:3 | q:6]; global dofs [2 per edge | 2 per cell | 3 per cell | vertices | edges].
// One thread per (test function, cell): it evaluates all bases at the quadrature points and accumulates its
// row of the 17 x 17 element Jacobian and its residual entry (entry-major storage, coalesced over cells).
// ------------------------------------------------------------------------------------------------
struct M1Cell {
  double X[3], Y[3], gx[3], gy[3], adet, hK, rs[3];  // rs = sigma_i / (2|K|); hK = CellDiameter (longest edge)
  double se[8], pe[3], qe[6];
};

struct M1Point {
  double a, x, y;            // weight * |det| * W, coordinates
  double Rx[8], Ry[8], dr[8], drD[8], chi[3], phi[6], phix[6], phiy[6];
};

__device__ __forceinline__ void m1_load_cell(const int c, const int ncells, const int nE, const int nV,
                                             const int* __restrict__ cells,
                                             const double2* __restrict__ coords,
                                             const int* __restrict__ cell_edges,
                                             const double* __restrict__ sign, const double* __restrict__ u,
                                             M1Cell& K) {
  int v[3], e[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    v[i] = cells[3 * (size_t)c + i];
    e[i] = cell_edges[3 * (size_t)c + i];
    const double2 P = coords[v[i]];
    K.X[i] = P.x;
    K.Y[i] = P.y;
  }
  const double j00 = K.X[1] - K.X[0], j01 = K.X[2] - K.X[0], j10 = K.Y[1] - K.Y[0], j11 = K.Y[2] - K.Y[0];
  const double det = j00 * j11 - j01 * j10, id = 1.0 / det;
  K.adet = fabs(det);
  {
    const double e2 = (K.X[2] - K.X[1]) * (K.X[2] - K.X[1]) + (K.Y[2] - K.Y[1]) * (K.Y[2] - K.Y[1]);
    K.hK = sqrt(fmax(fmax(j00 * j00 + j10 * j10, j01 * j01 + j11 * j11), e2));
  }
  K.gx[1] = j11 * id;  K.gy[1] = -j01 * id;
  K.gx[2] = -j10 * id; K.gy[2] = j00 * id;
  K.gx[0] = -(K.gx[1] + K.gx[2]);
  K.gy[0] = -(K.gy[1] + K.gy[2]);
#pragma unroll
  for (int i = 0; i < 3; ++i) K.rs[i] = sign[3 * (size_t)c + i] / K.adet;   // 2|K| = |det|
  if (u != nullptr) {
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      K.se[2 * i] = u[2 * e[i]];
      K.se[2 * i + 1] = u[2 * e[i] + 1];
      K.pe[i] = u[2 * nE + 2 * ncells + 3 * c + i];
    }
    K.se[6] = u[2 * nE + 2 * c];
    K.se[7] = u[2 * nE + 2 * c + 1];
    const int q0 = 2 * nE + 5 * ncells;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      K.qe[i] = u[q0 + v[i]];
      K.qe[3 + i] = u[q0 + nV + e[i]];
    }
  }
}

__device__ __forceinline__ void m1_eval_point(const M1Cell& K, const double l1, const double l2, const double w,
                                              const double D1, const double D2, M1Point& Q) {
  const double lam[3] = {1.0 - l1 - l2, l1, l2};
  Q.x = lam[0] * K.X[0] + lam[1] * K.X[1] + lam[2] * K.X[2];
  Q.y = lam[0] * K.Y[0] + lam[1] * K.Y[1] + lam[2] * K.Y[2];
  const double t = kFourPi * Q.x * Q.y;
  Q.a = w * K.adet * (t * t);
  // RT2: (facet i, end point a, scale) for the 6 facet functions, then the two interior ones (unsigned scale)
  const int fi[8] = {0, 0, 1, 1, 2, 2, 1, 2};
  const int fa[8] = {1, 2, 0, 2, 0, 1, 1, 2};
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const int i = fi[m], a = fa[m];
    const double c = (m < 6) ? K.rs[i] : 1.0 / K.adet;
    const double dx = Q.x - K.X[i], dy = Q.y - K.Y[i];
    Q.Rx[m] = lam[a] * dx * c;
    Q.Ry[m] = lam[a] * dy * c;
    const double dxx = c * (K.gx[a] * dx + lam[a]);
    const double dyy = c * (K.gy[a] * dy + lam[a]);
    Q.dr[m] = dxx + dyy + 2.0 * Q.Rx[m] / Q.x + 2.0 * Q.Ry[m] / Q.y;
    Q.drD[m] = D2 * (dxx + 2.0 * Q.Rx[m] / Q.x) + D1 * (dyy + 2.0 * Q.Ry[m] / Q.y);
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    Q.chi[i] = lam[i];
    Q.phi[i] = lam[i] * (2.0 * lam[i] - 1.0);
    Q.phix[i] = (4.0 * lam[i] - 1.0) * K.gx[i];
    Q.phiy[i] = (4.0 * lam[i] - 1.0) * K.gy[i];
  }
  const int ej[3] = {1, 0, 0}, ek[3] = {2, 2, 1};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    Q.phi[3 + i] = 4.0 * lam[ej[i]] * lam[ek[i]];
    Q.phix[3 + i] = 4.0 * (lam[ej[i]] * K.gx[ek[i]] + lam[ek[i]] * K.gx[ej[i]]);
    Q.phiy[3 + i] = 4.0 * (lam[ej[i]] * K.gy[ek[i]] + lam[ek[i]] * K.gy[ej[i]]);
  }
}

// Je: [289][ncells] (row-major local (test, trial), entry-major over cells), Fe: [17][ncells]
__global__ void __launch_bounds__(kMxThreads)
mixed_m1_cells_kernel(const MxRule R, const double D1, const double D2, const double geps, const double gcoef,
                      const int qmode, const double q_react, const double q_x2p, const double q_x2q,
                      const double supg_stab, const int ncells,
                      const int nE, const int nV, const int* __restrict__ cells,
                      const double2* __restrict__ coords, const int* __restrict__ cell_edges,
                      const double* __restrict__ sign, const double* __restrict__ u,
                      const double* __restrict__ f2q, const double* __restrict__ f3q,
                      double* __restrict__ Je, double* __restrict__ Fe) {
  const long long t = (long long)blockIdx.x * kMxThreads + threadIdx.x;
  if (t >= 17ll * ncells) return;
  const int row = (int)(t / ncells), c = (int)(t % ncells);
  M1Cell K;
  m1_load_cell(c, ncells, nE, nV, cells, coords, cell_edges, sign, u, K);
  double J[17], F = 0.0;
#pragma unroll
  for (int b = 0; b < 17; ++b) J[b] = 0.0;
  for (int k = 0; k < R.nq; ++k) {
    M1Point Q;
    m1_eval_point(K, R.l1[k], R.l2[k], R.w[k], D1, D2, Q);
    double sx = 0.0, sy = 0.0, drDs = 0.0, p = 0.0, q = 0.0, qx = 0.0;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      sx += K.se[m] * Q.Rx[m];
      sy += K.se[m] * Q.Ry[m];
      drDs += K.se[m] * Q.drD[m];
    }
#pragma unroll
    for (int b = 0; b < 3; ++b) p += K.pe[b] * Q.chi[b];
#pragma unroll
    for (int m = 0; m < 6; ++m) {
      q += K.qe[m] * Q.phi[m];
      qx += K.qe[m] * Q.phix[m];
    }
    const double x2 = Q.x * Q.x;
    const double iq = 1.0 / (q + geps);
    const double tg = gcoef * x2 * p * iq;
    const double G = tg * p, Gp = 2.0 * tg, Gq = -G * iq;
    const double a = Q.a;
    if (row < 8) {
      const double rx = Q.Rx[row], ry = Q.Ry[row], dra = Q.dr[row];
      F += a * (sx * rx + sy * ry + G * rx - p * dra);
#pragma unroll
      for (int b = 0; b < 8; ++b) J[b] += a * (rx * Q.Rx[b] + ry * Q.Ry[b]);
#pragma unroll
      for (int b = 0; b < 3; ++b) J[8 + b] += a * (Gp * rx - dra) * Q.chi[b];
#pragma unroll
      for (int m = 0; m < 6; ++m) J[11 + m] += a * Gq * rx * Q.phi[m];
    } else if (row < 11) {
      const double ch = Q.chi[row - 8];
      F += a * (-drDs + f2q[(size_t)c * R.nq + k]) * ch;
#pragma unroll
      for (int b = 0; b < 8; ++b) J[b] -= a * ch * Q.drD[b];
    } else {
      const double ph = Q.phi[row - 11], phx = Q.phix[row - 11];
      const double f3 = f3q[(size_t)c * R.nq + k];
      if (qmode != 1) {
        // test_for_paper_convergence.py:140-141: (-div_rad(r2vec) q + r2^2 p) w W - (r2vec.grad w) q W  (q_react = 0,
        // q_x2p = 1); C0IP_pure_advection_spherical_convergence.py:109-110: (q - div_rad(r2vec) q) w W - ... (1, 0);
        // qmode 2: the same with r2vec = (-1,0) (..._C0IP_otherBCs.py:58,131-132)
        const double dir = qmode == 2 ? -1.0 : 1.0;
#pragma unroll
        for (int b = 0; b < 3; ++b) J[8 + b] += a * x2 * q_x2p * Q.chi[b] * ph;
        const double cr = q_react - dir * 2.0 / Q.x;
        F += a * ((cr * q + x2 * q_x2p * p - f3) * ph - dir * q * phx);
#pragma unroll
        for (int m = 0; m < 6; ++m) J[11 + m] += a * (cr * Q.phi[m] * ph - dir * phx * Q.phi[m]);
      } else {
        // (q_react q + r2vec.grad q + r2^2 (q_x2p p + q_x2q q) - f3) (w + tau r2vec.grad w) W, tau = supg_stab hK/2:
        // ...MixedPrimal_convergence.py:123 (1,1,0,0) and ..._SUPG.py:118-129 (1,0,1,1; "r2**2*q" as written there)
        const double tw = ph + 0.5 * supg_stab * K.hK * phx;
        F += a * (q_react * q + qx + x2 * (q_x2p * p + q_x2q * q) - f3) * tw;
#pragma unroll
        for (int b = 0; b < 3; ++b) J[8 + b] += a * x2 * q_x2p * Q.chi[b] * tw;
#pragma unroll
        for (int m = 0; m < 6; ++m) J[11 + m] += a * ((q_react + x2 * q_x2q) * Q.phi[m] + Q.phix[m]) * tw;
      }
    }
  }
  Fe[(size_t)row * ncells + c] = F;
  if (Je != nullptr) {
#pragma unroll
    for (int b = 0; b < 17; ++b) Je[((size_t)row * 17 + b) * ncells + c] = J[b];
  }
}

// squares of e_div(s), e_0(p), e_1(q) for k = 1; exq[ncells*nq*7] = (s_x, s_y, div_rad s, p, q, dq/dx, dq/dy);
// full_grad: the q error takes the whole gradient (...MixedPrimal_convergence.py:148), else d/dx only (:175)
__global__ void __launch_bounds__(kMxThreads)
mixed_m1_errors_kernel(const MxRule R, const int full_grad, const int ncells, const int nE, const int nV,
                       const int* __restrict__ cells, const double2* __restrict__ coords,
                       const int* __restrict__ cell_edges, const double* __restrict__ sign,
                       const double* __restrict__ u, const double* __restrict__ exq,
                       double* __restrict__ partial) {
  double e0 = 0.0, e1 = 0.0, e2 = 0.0;
  for (int c = blockIdx.x * kMxThreads + threadIdx.x; c < ncells; c += kMxBlocks * kMxThreads) {
    M1Cell K;
    m1_load_cell(c, ncells, nE, nV, cells, coords, cell_edges, sign, u, K);
    double c0 = 0.0, c1 = 0.0, c2 = 0.0;
    for (int k = 0; k < R.nq; ++k) {
      M1Point Q;
      m1_eval_point(K, R.l1[k], R.l2[k], R.w[k], 1.0, 1.0, Q);
      double sx = 0.0, sy = 0.0, drs = 0.0, p = 0.0, q = 0.0, qx = 0.0, qy = 0.0;
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        sx += K.se[m] * Q.Rx[m];
        sy += K.se[m] * Q.Ry[m];
        drs += K.se[m] * Q.dr[m];
      }
#pragma unroll
      for (int b = 0; b < 3; ++b) p += K.pe[b] * Q.chi[b];
#pragma unroll
      for (int m = 0; m < 6; ++m) {
        q += K.qe[m] * Q.phi[m];
        qx += K.qe[m] * Q.phix[m];
        qy += K.qe[m] * Q.phiy[m];
      }
      const double* ex = exq + ((size_t)c * R.nq + k) * 7;
      const double d0 = ex[0] - sx, d1 = ex[1] - sy, d2 = ex[2] - drs, dp = ex[3] - p, dq = ex[4] - q,
                   dqx = ex[5] - qx, dqy = full_grad ? ex[6] - qy : 0.0;
      c0 += Q.a * (d0 * d0 + d1 * d1 + d2 * d2);
      c1 += Q.a * (dp * dp);
      c2 += Q.a * (dq * dq + dqx * dqx + dqy * dqy);
    }
    e0 += c0;
    e1 += c1;
    e2 += c2;
  }
  __shared__ double sh[3][kMxThreads];
  sh[0][threadIdx.x] = e0;
  sh[1][threadIdx.x] = e1;
  sh[2][threadIdx.x] = e2;
  __syncthreads();
  for (int o = kMxThreads / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o)
      for (int m = 0; m < 3; ++m) sh[m][threadIdx.x] += sh[m][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0)
    for (int m = 0; m < 3; ++m) partial[3 * blockIdx.x + m] = sh[m][0];
}

// ---- CG2 facet operators of the q-equation for k = 1 (state independent: once per mesh) ---------------------
struct FacetRule {
  int nq;
  double s[8], w[8];  // Gauss-Legendre on [0,1]
};

__device__ __forceinline__ void cg2_values(const double* lam, double* phi) {
#pragma unroll
  for (int i = 0; i < 3; ++i) phi[i] = lam[i] * (2.0 * lam[i] - 1.0);
  phi[3] = 4.0 * lam[1] * lam[2];
  phi[4] = 4.0 * lam[0] * lam[2];
  phi[5] = 4.0 * lam[0] * lam[1];
}

// d phi_k / dn for the 6 CG2 functions of a cell with barycentric gradients (gx, gy)
__device__ __forceinline__ void cg2_normal_derivs(const double* lam, const double* gx, const double* gy,
                                                  const double nx, const double ny, double* d) {
  double gn[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) gn[i] = gx[i] * nx + gy[i] * ny;
#pragma unroll
  for (int i = 0; i < 3; ++i) d[i] = (4.0 * lam[i] - 1.0) * gn[i];
  d[3] = 4.0 * (lam[1] * gn[2] + lam[2] * gn[1]);
  d[4] = 4.0 * (lam[0] * gn[2] + lam[2] * gn[0]);
  d[5] = 4.0 * (lam[0] * gn[1] + lam[1] * gn[0]);
}

// exterior facets: Jx[36][nfe] = sum ds cq phi_k phi_l, cx[6][nfe] = sum ds cd q_ex phi_k,
// nat[2][nfe] = natural p data on the two flux dofs of the facet (kind bit 4), with
//   cq = [kind&1] (r2vec.n) W + stab2 (hK neg(r2vec.n)^2 + 8/hK),  cd = [kind&2] (r2vec.n) W - stab2 (...)
// (test_for_paper_convergence.py:142-145,148-149,152-153; hK = FacetArea, r2vec = (1,0))
__global__ void __launch_bounds__(kMxThreads)
cg2_exterior_facets_kernel(const FacetRule R, const double stab2, const int nfe,
                           const int* __restrict__ fverts, const int* __restrict__ fcell,
                           const int* __restrict__ kind, const int* __restrict__ cells,
                           const double2* __restrict__ coords, const double* __restrict__ qex,
                           const double* __restrict__ pex, double* __restrict__ Jx,
                           double* __restrict__ cx, double* __restrict__ nat) {
  const int f = blockIdx.x * kMxThreads + threadIdx.x;
  if (f >= nfe) return;
  const int va = fverts[2 * f], vb = fverts[2 * f + 1], c = fcell[f], kd = kind[f];
  int la = 0, lb = 0;
  double2 X[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int v = cells[3 * (size_t)c + i];
    X[i] = coords[v];
    if (v == va) la = i;
    if (v == vb) lb = i;
  }
  const int lo = 3 - la - lb;  // local vertex opposite the facet
  const double tx = X[lb].x - X[la].x, ty = X[lb].y - X[la].y;
  const double L = sqrt(tx * tx + ty * ty);
  const double gnx = ty / L, gny = -tx / L;  // global normal: tangent (low -> high vertex) rotated clockwise
  double nx = gnx, ny = gny;
  if (nx * (X[lo].x - X[la].x) + ny * (X[lo].y - X[la].y) > 0.0) { nx = -nx; ny = -ny; }
  const double sgn = (nx * gnx + ny * gny) > 0.0 ? 1.0 : -1.0;
  const double bn = (kd & 8) ? -nx : nx;  // r2vec = (1,0), or (-1,0) with kind bit 8 (..._C0IP_otherBCs.py:58)
  const double neg = 0.5 * (fabs(bn) - bn);
  const double pen = stab2 * (L * neg * neg + 8.0 / L);
  double J[36], cv[6], n0 = 0.0, n1 = 0.0;
#pragma unroll
  for (int i = 0; i < 36; ++i) J[i] = 0.0;
#pragma unroll
  for (int i = 0; i < 6; ++i) cv[i] = 0.0;
  for (int k = 0; k < R.nq; ++k) {
    const double s = R.s[k];
    double lam[3] = {0.0, 0.0, 0.0};
    lam[la] = 1.0 - s;
    lam[lb] = s;
    const double x = X[la].x + s * tx, y = X[la].y + s * ty;
    const double t = kFourPi * x * y, W = t * t;
    const double ds = R.w[k] * L;
    const double cq = ((kd & 1) ? bn * W : 0.0) + pen;
    const double cd = ((kd & 2) ? bn * W : 0.0) - pen;
    double phi[6];
    cg2_values(lam, phi);
    const double dq = ds * cd * qex[(size_t)f * R.nq + k];
#pragma unroll
    for (int a = 0; a < 6; ++a) {
      cv[a] += dq * phi[a];
#pragma unroll
      for (int b = 0; b < 6; ++b) J[6 * a + b] += ds * cq * phi[a] * phi[b];
    }
    if (kd & 4) {
      const double pw = ds * W * pex[(size_t)f * R.nq + k];
      n0 += pw * (1.0 - s);
      n1 += pw * s;
    }
  }
#pragma unroll
  for (int i = 0; i < 36; ++i) Jx[(size_t)i * nfe + f] = J[i];
#pragma unroll
  for (int i = 0; i < 6; ++i) cx[(size_t)i * nfe + f] = cv[i];
  nat[f] = n0 * sgn / L;
  nat[(size_t)nfe + f] = n1 * sgn / L;
}

// interior facets: Ji[144][nfi], row-major over the 12 macro dofs (6 of cell0 = '+', then 6 of cell1):
// gamma avg(hK)^2 (jump(grad q).n+)(jump(grad w).n+) W dS with hK = FacetArea (:144).  One thread per (row, facet).
__global__ void __launch_bounds__(kMxThreads)
cg2_interior_facets_kernel(const FacetRule R, const double gamma, const int nfi,
                           const int* __restrict__ fverts, const int* __restrict__ fcell,
                           const int* __restrict__ cells, const double2* __restrict__ coords,
                           const double* __restrict__ hbar, double* __restrict__ Ji) {
  const long long t = (long long)blockIdx.x * kMxThreads + threadIdx.x;
  if (t >= 12ll * nfi) return;
  const int row = (int)(t / nfi), f = (int)(t % nfi);
  const int va = fverts[2 * f], vb = fverts[2 * f + 1];
  int la[2], lb[2];
  double gx[2][3], gy[2][3];
  double2 A = make_double2(0.0, 0.0), B = A, O = A;
#pragma unroll
  for (int side = 0; side < 2; ++side) {
    const int c = fcell[2 * f + side];
    double2 X[3];
    la[side] = lb[side] = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const int v = cells[3 * (size_t)c + i];
      X[i] = coords[v];
      if (v == va) la[side] = i;
      if (v == vb) lb[side] = i;
    }
    const double j00 = X[1].x - X[0].x, j01 = X[2].x - X[0].x, j10 = X[1].y - X[0].y, j11 = X[2].y - X[0].y;
    const double id = 1.0 / (j00 * j11 - j01 * j10);
    gx[side][1] = j11 * id;  gy[side][1] = -j01 * id;
    gx[side][2] = -j10 * id; gy[side][2] = j00 * id;
    gx[side][0] = -(gx[side][1] + gx[side][2]);
    gy[side][0] = -(gy[side][1] + gy[side][2]);
    if (side == 0) {
      A = X[la[0]];
      B = X[lb[0]];
      O = X[3 - la[0] - lb[0]];
    }
  }
  const double tx = B.x - A.x, ty = B.y - A.y;
  const double L = sqrt(tx * tx + ty * ty);
  double nx = ty / L, ny = -tx / L;
  if (nx * (O.x - A.x) + ny * (O.y - A.y) > 0.0) { nx = -nx; ny = -ny; }  // n+ = outward normal of cell0
  double J[12];
#pragma unroll
  for (int b = 0; b < 12; ++b) J[b] = 0.0;
  for (int k = 0; k < R.nq; ++k) {
    const double s = R.s[k];
    const double x = A.x + s * tx, y = A.y + s * ty;
    const double tt = kFourPi * x * y;
    const double hb = hbar ? hbar[f] : L;  // avg(hK): FacetArea (= L) or the mean CellDiameter of the two cells
    const double coef = gamma * hb * hb * R.w[k] * L * (tt * tt);
    double d[12];
#pragma unroll
    for (int side = 0; side < 2; ++side) {
      double lam[3] = {0.0, 0.0, 0.0};
      lam[la[side]] = 1.0 - s;
      lam[lb[side]] = s;
      cg2_normal_derivs(lam, gx[side], gy[side], nx, ny, d + 6 * side);
    }
#pragma unroll
    for (int b = 6; b < 12; ++b) d[b] = -d[b];
    const double dr = coef * d[row];
#pragma unroll
    for (int b = 0; b < 12; ++b) J[b] += dr * d[b];
  }
#pragma unroll
  for (int b = 0; b < 12; ++b) Ji[((size_t)row * 12 + b) * nfi + f] = J[b];
}

static int facet_rule_arg(int nq, const double* rule_host, FacetRule& R) {
  if (nq < 1 || nq > 8 || rule_host == nullptr) return CM_E_INVALID;
  R.nq = nq;
  for (int k = 0; k < nq; ++k) {
    R.s[k] = rule_host[2 * k];
    R.w[k] = rule_host[2 * k + 1];
  }
  return CM_OK;
}

int cg2_exterior_facets(int nq, const double* rule_host, double stab2, int nfe, const int* fverts,
                        const int* fcell, const int* kind, const int* cells, const double* coords,
                        const double* qex, const double* pex, double* Jx, double* cx, double* nat,
                        cudaStream_t st) {
  FacetRule R;
  if (facet_rule_arg(nq, rule_host, R)) {
    set_error("cm_cg2_exterior_facets: bad facet rule (1..8 points)");
    return CM_E_INVALID;
  }
  if (nfe <= 0) return CM_OK;
  cg2_exterior_facets_kernel<<<(nfe + kMxThreads - 1) / kMxThreads, kMxThreads, 0, st>>>(
      R, stab2, nfe, fverts, fcell, kind, cells, (const double2*)coords, qex, pex, Jx, cx, nat);
  CM_LAUNCH_CHECK();
  return CM_OK;
}

int cg2_interior_facets(int nq, const double* rule_host, double gamma, int nfi, const int* fverts,
                        const int* fcell, const int* cells, const double* coords, const double* hbar, double* Ji,
                        cudaStream_t st) {
  FacetRule R;
  if (facet_rule_arg(nq, rule_host, R)) {
    set_error("cm_cg2_interior_facets: bad facet rule (1..8 points)");
    return CM_E_INVALID;
  }
  if (nfi <= 0) return CM_OK;
  const long long nt = 12ll * nfi;
  cg2_interior_facets_kernel<<<(unsigned)((nt + kMxThreads - 1) / kMxThreads), kMxThreads, 0, st>>>(
      R, gamma, nfi, fverts, fcell, cells, (const double2*)coords, hbar, Ji);
  CM_LAUNCH_CHECK();
  return CM_OK;
}

// ------------------------------------------------------------------------------------------------
// Production mixed formulation DG2 x RT3 x CG3 (advection_diffusion_mixed.py:147-150,196-216).
// 31 local dofs [p: 6 DG2 | s: 15 RT3 | q: 10 CG3]; bases:
//   DG2: lambda_i (2 lambda_i - 1), 4 lambda_j lambda_k;
//   RT3: per facet i (end points j < k): lambda_j^2 N_i, lambda_j lambda_k N_i, lambda_k^2 N_i with
//        N_i = sigma_i (x - X_i)/(2|K|); interior: lambda_i lambda_m N_i (unsigned), i = 1,2, m = 0,1,2;
//   CG3: Lagrange P3 (vertices; per facet the node nearer the lower vertex, then the other; bubble).
// Global dofs [6 per cell | 3 per edge, 6 per cell | vertices, 2 per edge, 1 per cell].
// One thread per (test function, cell), as for k = 1.
// ------------------------------------------------------------------------------------------------
struct P2Cell {
  double X[3], Y[3], gx[3], gy[3], adet, rs[3];
  double pe[6], se[15], qe[10], po[6];
};

struct P2Point {
  double a, x, y;
  double Rx[15], Ry[15], dr[15], drD[15], chi[6], phi[10], phix[10];
};

__device__ __forceinline__ void p2m_load_cell(const int c, const int C, const int E, const int V,
                                              const int* __restrict__ cells,
                                              const double2* __restrict__ coords,
                                              const int* __restrict__ cell_edges,
                                              const double* __restrict__ sign, const double* __restrict__ u,
                                              const double* __restrict__ p_old, P2Cell& K) {
  int v[3], e[3];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    v[i] = cells[3 * (size_t)c + i];
    e[i] = cell_edges[3 * (size_t)c + i];
    const double2 P = coords[v[i]];
    K.X[i] = P.x;
    K.Y[i] = P.y;
  }
  const double j00 = K.X[1] - K.X[0], j01 = K.X[2] - K.X[0], j10 = K.Y[1] - K.Y[0], j11 = K.Y[2] - K.Y[0];
  const double det = j00 * j11 - j01 * j10, id = 1.0 / det;
  K.adet = fabs(det);
  K.gx[1] = j11 * id;  K.gy[1] = -j01 * id;
  K.gx[2] = -j10 * id; K.gy[2] = j00 * id;
  K.gx[0] = -(K.gx[1] + K.gx[2]);
  K.gy[0] = -(K.gy[1] + K.gy[2]);
#pragma unroll
  for (int i = 0; i < 3; ++i) K.rs[i] = sign[3 * (size_t)c + i] / K.adet;
  const size_t s0 = 6 * (size_t)C, q0 = s0 + 3 * (size_t)E + 6 * (size_t)C;
#pragma unroll
  for (int b = 0; b < 6; ++b) {
    K.pe[b] = u[6 * (size_t)c + b];
    K.po[b] = p_old ? p_old[6 * (size_t)c + b] : 0.0;
    K.se[9 + b] = u[s0 + 3 * (size_t)E + 6 * (size_t)c + b];
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int m = 0; m < 3; ++m) K.se[3 * i + m] = u[s0 + 3 * (size_t)e[i] + m];
    K.qe[i] = u[q0 + v[i]];
    K.qe[3 + 2 * i] = u[q0 + V + 2 * (size_t)e[i]];
    K.qe[4 + 2 * i] = u[q0 + V + 2 * (size_t)e[i] + 1];
  }
  K.qe[9] = u[q0 + V + 2 * (size_t)E + c];
}

__device__ __forceinline__ void p2m_eval_point(const P2Cell& K, const double l1, const double l2, const double w,
                                               const double D1, const double D2, P2Point& Q) {
  const double lam[3] = {1.0 - l1 - l2, l1, l2};
  Q.x = lam[0] * K.X[0] + lam[1] * K.X[1] + lam[2] * K.X[2];
  Q.y = lam[0] * K.Y[0] + lam[1] * K.Y[1] + lam[2] * K.Y[2];
  const double t = kFourPi * Q.x * Q.y;
  Q.a = w * K.adet * (t * t);
  // RT3: (facet i, quadratic lambda_a lambda_b); facets: (j,j), (j,k), (k,k); interior: i in {1,2}, (i, m)
  const int fi[15] = {0, 0, 0, 1, 1, 1, 2, 2, 2, 1, 1, 1, 2, 2, 2};
  const int fa[15] = {1, 1, 2, 0, 0, 2, 0, 0, 1, 1, 1, 1, 2, 2, 2};
  const int fb[15] = {1, 2, 2, 0, 2, 2, 0, 1, 1, 0, 1, 2, 0, 1, 2};
#pragma unroll
  for (int m = 0; m < 15; ++m) {
    const int i = fi[m], a = fa[m], b = fb[m];
    const double c = (m < 9) ? K.rs[i] : 1.0 / K.adet;
    const double dx = Q.x - K.X[i], dy = Q.y - K.Y[i];
    const double qv = lam[a] * lam[b];
    const double qgx = lam[a] * K.gx[b] + lam[b] * K.gx[a];
    const double qgy = lam[a] * K.gy[b] + lam[b] * K.gy[a];
    Q.Rx[m] = qv * dx * c;
    Q.Ry[m] = qv * dy * c;
    const double dxx = c * (qgx * dx + qv);
    const double dyy = c * (qgy * dy + qv);
    Q.dr[m] = dxx + dyy + 2.0 * Q.Rx[m] / Q.x + 2.0 * Q.Ry[m] / Q.y;
    Q.drD[m] = D2 * (dxx + 2.0 * Q.Rx[m] / Q.x) + D1 * (dyy + 2.0 * Q.Ry[m] / Q.y);
  }
  const int ej[3] = {1, 0, 0}, ek[3] = {2, 2, 1};
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const double l = lam[i];
    Q.chi[i] = l * (2.0 * l - 1.0);
    Q.chi[3 + i] = 4.0 * lam[ej[i]] * lam[ek[i]];
    Q.phi[i] = 0.5 * l * (3.0 * l - 1.0) * (3.0 * l - 2.0);
    Q.phix[i] = 0.5 * (27.0 * l * l - 18.0 * l + 2.0) * K.gx[i];
    const double lj = lam[ej[i]], lk = lam[ek[i]], gj = K.gx[ej[i]], gk = K.gx[ek[i]];
    Q.phi[3 + 2 * i] = 4.5 * lj * lk * (3.0 * lj - 1.0);
    Q.phix[3 + 2 * i] = 4.5 * ((gj * lk + lj * gk) * (3.0 * lj - 1.0) + lj * lk * 3.0 * gj);
    Q.phi[4 + 2 * i] = 4.5 * lk * lj * (3.0 * lk - 1.0);
    Q.phix[4 + 2 * i] = 4.5 * ((gk * lj + lk * gj) * (3.0 * lk - 1.0) + lk * lj * 3.0 * gk);
  }
  Q.phi[9] = 27.0 * lam[0] * lam[1] * lam[2];
  Q.phix[9] = 27.0 * (K.gx[0] * lam[1] * lam[2] + lam[0] * K.gx[1] * lam[2] + lam[0] * lam[1] * K.gx[2]);
}

// Je: [961][ncells] (31 x 31 row-major, entry-major over cells), Fe: [31][ncells]
__global__ void __launch_bounds__(kMxThreads)
mixed_p2_cells_kernel(const MxRule R, const double D1, const double D2, const double conc, const double inv_dt,
                      const int ncells, const int nE, const int nV, const int* __restrict__ cells,
                      const double2* __restrict__ coords, const int* __restrict__ cell_edges,
                      const double* __restrict__ sign, const double* __restrict__ u,
                      const double* __restrict__ p_old, double* __restrict__ Je, double* __restrict__ Fe) {
  const long long t = (long long)blockIdx.x * kMxThreads + threadIdx.x;
  if (t >= 31ll * ncells) return;
  const int row = (int)(t / ncells), c = (int)(t % ncells);
  P2Cell K;
  p2m_load_cell(c, ncells, nE, nV, cells, coords, cell_edges, sign, u, p_old, K);
  double J[31], F = 0.0;
#pragma unroll
  for (int b = 0; b < 31; ++b) J[b] = 0.0;
  const double gs_scale = 1.0 / (4.0 * conc * 3.14159265358979323846), gs_lin = kFourPi * conc;
  for (int k = 0; k < R.nq; ++k) {
    P2Point Q;
    p2m_eval_point(K, R.l1[k], R.l2[k], R.w[k], D1, D2, Q);
    double p = 0.0, po = 0.0, sx = 0.0, sy = 0.0, drDs = 0.0, q = 0.0, qx = 0.0;
#pragma unroll
    for (int b = 0; b < 6; ++b) {
      p += K.pe[b] * Q.chi[b];
      po += K.po[b] * Q.chi[b];
    }
#pragma unroll
    for (int m = 0; m < 15; ++m) {
      sx += K.se[m] * Q.Rx[m];
      sy += K.se[m] * Q.Ry[m];
      drDs += K.se[m] * Q.drD[m];
    }
#pragma unroll
    for (int m = 0; m < 10; ++m) {
      q += K.qe[m] * Q.phi[m];
      qx += K.qe[m] * Q.phix[m];
    }
    const double x2 = Q.x * Q.x;
    // Gstar (advection_diffusion_mixed.py:17-23): conditional(|p/(4 c pi) - q| > 1e-3, (p/q) p r2^2, 4 pi c p r2^2)
    const bool cond = fabs(gs_scale * p - q) > 1e-3;
    const double qs = cond ? q : 1.0;
    const double G = cond ? p / qs * p * x2 : gs_lin * p * x2;
    const double Gp = cond ? 2.0 * p / qs * x2 : gs_lin * x2;
    const double Gq = cond ? -p * p / (qs * qs) * x2 : 0.0;
    const double a = Q.a;
    if (row < 6) {  // v = chi_row: div_rad(D s) v W (+ (p - p_old)/dt v W)
      const double ch = Q.chi[row];
      F += a * (drDs + inv_dt * (p - po)) * ch;
#pragma unroll
      for (int b = 0; b < 6; ++b) J[b] += a * inv_dt * ch * Q.chi[b];
#pragma unroll
      for (int m = 0; m < 15; ++m) J[6 + m] += a * ch * Q.drD[m];
    } else if (row < 21) {  // tau = Psi: (s + G e_x).tau W - p div_rad(tau) W
      const int r = row - 6;
      const double rx = Q.Rx[r], ry = Q.Ry[r], dra = Q.dr[r];
      F += a * (sx * rx + sy * ry + G * rx - p * dra);
#pragma unroll
      for (int b = 0; b < 6; ++b) J[b] += a * (Gp * rx - dra) * Q.chi[b];
#pragma unroll
      for (int m = 0; m < 15; ++m) J[6 + m] += a * (rx * Q.Rx[m] + ry * Q.Ry[m]);
#pragma unroll
      for (int m = 0; m < 10; ++m) J[21 + m] += a * Gq * rx * Q.phi[m];
    } else {  // w = phi: (dq/dx + x^2 p) w W
      const double ph = Q.phi[row - 21];
      F += a * (qx + x2 * p) * ph;
#pragma unroll
      for (int b = 0; b < 6; ++b) J[b] += a * x2 * Q.chi[b] * ph;
#pragma unroll
      for (int m = 0; m < 10; ++m) J[21 + m] += a * ph * Q.phix[m];
    }
  }
  Fe[(size_t)row * ncells + c] = F;
  if (Je != nullptr) {
#pragma unroll
    for (int b = 0; b < 31; ++b) Je[((size_t)row * 31 + b) * ncells + c] = J[b];
  }
}

static int mx_rule(int nq, const double* rule_host, MxRule& R) {
  if (nq < 1 || nq > kMxRule || rule_host == nullptr) return CM_E_INVALID;
  R.nq = nq;
  for (int k = 0; k < nq; ++k) {
    R.l1[k] = rule_host[3 * k];
    R.l2[k] = rule_host[3 * k + 1];
    R.w[k] = rule_host[3 * k + 2];
  }
  return CM_OK;
}

int mixed_m0_cells(int nq, const double* rule_host, double D1, double D2, double geps, int ncells,
                   const int* cells, const double* coords, const int* cell_edges, const double* sign,
                   int nE, const double* u, const double* f2q, const double* f3q, double* Je, double* Fe,
                   cudaStream_t st) {
  MxRule R;
  if (mx_rule(nq, rule_host, R)) {
    set_error("cm_mixed_m0_cells: bad quadrature rule (1..%d points)", kMxRule);
    return CM_E_INVALID;
  }
  mixed_m0_cells_kernel<<<(ncells + kMxThreads - 1) / kMxThreads, kMxThreads, 0, st>>>(
      R, D1, D2, geps, ncells, cells, (const double2*)coords, cell_edges, sign, nE, u, f2q, f3q, Je, Fe);
  CM_LAUNCH_CHECK();
  CM_BYTES((double)ncells * (12.0 + 12.0 + 24.0 + 56.0 + 16.0 * nq + 56.0 + (Je ? 392.0 : 0.0)));
  return CM_OK;
}

int csr_gather(int nslots, const int* ptr, const int* idx, const double* src, double* out, int accumulate,
               cudaStream_t st) {
  if (nslots <= 0) return CM_OK;
  csr_gather_kernel<<<(nslots + 255) / 256, 256, 0, st>>>(nslots, ptr, idx, src, out, accumulate);
  CM_LAUNCH_CHECK();
  return CM_OK;
}

int add_mapped(int n, const int* map, const double* src, double* dst, cudaStream_t st) {
  if (n <= 0) return CM_OK;
  add_mapped_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, map, src, dst);
  CM_LAUNCH_CHECK();
  return CM_OK;
}

size_t mixed_workspace_doubles() { return 4 + 3 * (size_t)kMxBlocks; }

int mixed_m0_errors(int nq, const double* rule_host, int ncells, const int* cells, const double* coords,
                    const int* cell_edges, const double* sign, int nE, const double* u, const double* exq,
                    double* work, cudaStream_t st) {
  MxRule R;
  if (mx_rule(nq, rule_host, R)) {
    set_error("cm_mixed_m0_errors: bad quadrature rule (1..%d points)", kMxRule);
    return CM_E_INVALID;
  }
  mixed_m0_errors_kernel<<<kMxBlocks, kMxThreads, 0, st>>>(R, ncells, cells, (const double2*)coords,
                                                          cell_edges, sign, nE, u, exq, work + 4);
  CM_LAUNCH_CHECK();
  mixed_final_sum_kernel<<<1, 32, 0, st>>>(kMxBlocks, work + 4, work);
  CM_LAUNCH_CHECK();
  return CM_OK;
}

int mixed_m1_cells(int nq, const double* rule_host, double D1, double D2, double geps, double gcoef, int qmode,
                   double q_react, double q_x2p, double q_x2q, double supg_stab, int ncells, int nE,
                   int nV, const int* cells, const double* coords, const int* cell_edges, const double* sign,
                   const double* u, const double* f2q, const double* f3q, double* Je, double* Fe,
                   cudaStream_t st) {
  MxRule R;
  if (mx_rule(nq, rule_host, R)) {
    set_error("cm_mixed_m1_cells: bad quadrature rule (1..%d points)", kMxRule);
    return CM_E_INVALID;
  }
  const long long nt = 17ll * ncells;
  mixed_m1_cells_kernel<<<(unsigned)((nt + kMxThreads - 1) / kMxThreads), kMxThreads, 0, st>>>(
      R, D1, D2, geps, gcoef, qmode, q_react, q_x2p, q_x2q, supg_stab, ncells, nE, nV, cells, (const double2*)coords,
      cell_edges, sign, u, f2q, f3q, Je, Fe);
  CM_LAUNCH_CHECK();
  CM_BYTES((double)ncells * (12.0 + 12.0 + 24.0 + 136.0 + 16.0 * nq + 136.0 + (Je ? 2312.0 : 0.0)));
  return CM_OK;
}

int mixed_m1_errors(int nq, const double* rule_host, int full_grad, int ncells, int nE, int nV, const int* cells,
                    const double* coords, const int* cell_edges, const double* sign, const double* u,
                    const double* exq, double* work, cudaStream_t st) {
  MxRule R;
  if (mx_rule(nq, rule_host, R)) {
    set_error("cm_mixed_m1_errors: bad quadrature rule (1..%d points)", kMxRule);
    return CM_E_INVALID;
  }
  mixed_m1_errors_kernel<<<kMxBlocks, kMxThreads, 0, st>>>(R, full_grad, ncells, nE, nV, cells, (const double2*)coords,
                                                          cell_edges, sign, u, exq, work + 4);
  CM_LAUNCH_CHECK();
  mixed_final_sum_kernel<<<1, 32, 0, st>>>(kMxBlocks, work + 4, work);
  CM_LAUNCH_CHECK();
  return CM_OK;
}

int mixed_p2_cells(int nq, const double* rule_host, double D1, double D2, double conc, double inv_dt, int ncells,
                   int nE, int nV, const int* cells, const double* coords, const int* cell_edges,
                   const double* sign, const double* u, const double* p_old, double* Je, double* Fe,
                   cudaStream_t st) {
  MxRule R;
  if (mx_rule(nq, rule_host, R)) {
    set_error("cm_mixed_p2_cells: bad quadrature rule (1..%d points)", kMxRule);
    return CM_E_INVALID;
  }
  const long long nt = 31ll * ncells;
  mixed_p2_cells_kernel<<<(unsigned)((nt + kMxThreads - 1) / kMxThreads), kMxThreads, 0, st>>>(
      R, D1, D2, conc, inv_dt, ncells, nE, nV, cells, (const double2*)coords, cell_edges, sign, u, p_old, Je, Fe);
  CM_LAUNCH_CHECK();
  CM_BYTES((double)ncells * (12.0 + 12.0 + 24.0 + 248.0 + 248.0 + (Je ? 7688.0 : 0.0)));
  return CM_OK;
}

}  // namespace cm
