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
constexpr int kMxRule = 16;
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

}  // namespace cm
