// Scalar functionals and the flux post-processing of the drivers:
//   error norms    assemble((u_ex-u_h)**2*[W]*dx), assemble((grad(u_ex)-grad(u_h))**2*[W]*dx)
//                  (SphericalAdvectionDiffusionReaction_convergence.py:117-118, ...MixedPrimal_convergence.py:148-150)
//   DG0 flux       project(dMatrix*grad(p_h) + D2*r2*r2*p_h/q_h*p_h*r2_vec, VectorFunctionSpace(mesh,"DG",0))
//                  (convergence/advection_diffusion_convergence.py:222-223)
//   boundary flux  assemble(dot(flux,n)*4*pi*r1**2*ds(bottom)) and its r1 / r2 parts
//                  (advection_diffusion_mixed.py:261-263)
// One thread per cell / facet; reductions are two-stage with a fixed grid and a fixed summation order
// (bit-reproducible, no atomics).
#include "cm_common.cuh"
#include "cm_decl.cuh"

namespace cm {

constexpr int kFnBlocks = 592;  // 148 SMs x 4
constexpr int kFnThreads = 256;
constexpr int kMaxRule = 16;

struct RuleArg {
  int nq;
  double l1[kMaxRule], l2[kMaxRule], w[kMaxRule];
};

struct CellGeo {
  double x[3], y[3], gx[3], gy[3], adet;
};

__device__ __forceinline__ void load_cell(const int* __restrict__ cells, const double2* __restrict__ coords,
                                          const int c, int v[3], CellGeo& g) {
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    v[i] = cells[3 * (size_t)c + i];
    const double2 X = coords[v[i]];
    g.x[i] = X.x;
    g.y[i] = X.y;
  }
  const double j00 = g.x[1] - g.x[0], j01 = g.x[2] - g.x[0];
  const double j10 = g.y[1] - g.y[0], j11 = g.y[2] - g.y[0];
  const double det = j00 * j11 - j01 * j10;
  const double id = 1.0 / det;
  g.gx[1] = j11 * id;
  g.gy[1] = -j01 * id;
  g.gx[2] = -j10 * id;
  g.gy[2] = j00 * id;
  g.gx[0] = -(g.gx[1] + g.gx[2]);
  g.gy[0] = -(g.gy[1] + g.gy[2]);
  g.adet = fabs(det);
}

// fixed-order block reduction of two values; result valid in thread 0
__device__ __forceinline__ void block_reduce2(double& a, double& b) {
  __shared__ double sa[kFnThreads], sb[kFnThreads];
  sa[threadIdx.x] = a;
  sb[threadIdx.x] = b;
  __syncthreads();
  for (int o = kFnThreads / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      sa[threadIdx.x] += sa[threadIdx.x + o];
      sb[threadIdx.x] += sb[threadIdx.x + o];
    }
    __syncthreads();
  }
  a = sa[0];
  b = sb[0];
}

__global__ void __launch_bounds__(kFnThreads)
error_norms_kernel(const RuleArg R, const int weighted, const int xonly, const int ncells,
                   const int* __restrict__ cells, const double2* __restrict__ coords,
                   const double* __restrict__ uh, const int stride, const double* __restrict__ exq,
                   const double2* __restrict__ grq, double* __restrict__ partial) {
  double s0 = 0.0, s1 = 0.0;
  for (int c = blockIdx.x * kFnThreads + threadIdx.x; c < ncells; c += kFnBlocks * kFnThreads) {
    int v[3];
    CellGeo g;
    load_cell(cells, coords, c, v, g);
    const double u0 = uh[(size_t)v[0] * stride], u1 = uh[(size_t)v[1] * stride],
                 u2 = uh[(size_t)v[2] * stride];
    const double ux = u0 * g.gx[0] + u1 * g.gx[1] + u2 * g.gx[2];
    const double uy = u0 * g.gy[0] + u1 * g.gy[1] + u2 * g.gy[2];
    double c0 = 0.0, c1 = 0.0;
    for (int k = 0; k < R.nq; ++k) {
      const double l1 = R.l1[k], l2 = R.l2[k], l0 = 1.0 - l1 - l2;
      double a = R.w[k] * g.adet;
      if (weighted) {
        const double x = l0 * g.x[0] + l1 * g.x[1] + l2 * g.x[2];
        const double y = l0 * g.y[0] + l1 * g.y[1] + l2 * g.y[2];
        const double t = kFourPi * x * y;
        a *= t * t;
      }
      const double d = exq[(size_t)c * R.nq + k] - (l0 * u0 + l1 * u1 + l2 * u2);
      const double2 ge = grq[(size_t)c * R.nq + k];
      const double dx = ge.x - ux, dy = ge.y - uy;
      c0 += a * (d * d);
      c1 += a * (xonly ? dx * dx : dx * dx + dy * dy);
    }
    s0 += c0;
    s1 += c1;
  }
  block_reduce2(s0, s1);
  if (threadIdx.x == 0) {
    partial[2 * blockIdx.x] = s0;
    partial[2 * blockIdx.x + 1] = s1;
  }
}

__global__ void __launch_bounds__(kFnThreads)
final_sum_kernel(const int nparts, const int ncomp, const double* __restrict__ partial,
                 double* __restrict__ out) {
  for (int comp = 0; comp < ncomp; comp += 2) {
    double a = 0.0, b = 0.0;
    for (int i = threadIdx.x; i < nparts; i += kFnThreads) {
      a += partial[(size_t)ncomp * i + comp];
      if (comp + 1 < ncomp) b += partial[(size_t)ncomp * i + comp + 1];
    }
    block_reduce2(a, b);
    if (threadIdx.x == 0) {
      out[comp] = a;
      if (comp + 1 < ncomp) out[comp + 1] = b;
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kFnThreads)
flux_dg0_kernel(const RuleArg R, const double D1, const double D2, const int ncells,
                const int* __restrict__ cells, const double2* __restrict__ coords,
                const double2* __restrict__ u, double2* __restrict__ flux) {
  const int c = blockIdx.x * kFnThreads + threadIdx.x;
  if (c >= ncells) return;
  int v[3];
  CellGeo g;
  load_cell(cells, coords, c, v, g);
  const double2 U0 = u[v[0]], U1 = u[v[1]], U2 = u[v[2]];
  const double px = U0.x * g.gx[0] + U1.x * g.gx[1] + U2.x * g.gx[2];
  const double py = U0.x * g.gy[0] + U1.x * g.gy[1] + U2.x * g.gy[2];
  double s = 0.0;  // sum_k w_k x^2 p^2 / q  (the weights sum to 1/2: the DG0 projection is 2*s)
  for (int k = 0; k < R.nq; ++k) {
    const double l1 = R.l1[k], l2 = R.l2[k], l0 = 1.0 - l1 - l2;
    const double x = l0 * g.x[0] + l1 * g.x[1] + l2 * g.x[2];
    const double p = l0 * U0.x + l1 * U1.x + l2 * U2.x;
    const double q = l0 * U0.y + l1 * U1.y + l2 * U2.y;
    s += R.w[k] * (x * x * p / q * p);
  }
  flux[c] = make_double2(D2 * (px + 2.0 * s), D1 * py);
}

// out partial layout: (r2 part, r1 part) = int (f_x n_x, f_y n_y) 4 pi r1^2 ds
__global__ void __launch_bounds__(kFnThreads)
boundary_flux_kernel(const int nfe, const int2* __restrict__ efverts, const int* __restrict__ efopp,
                     const int* __restrict__ efcell, const double2* __restrict__ coords,
                     const double2* __restrict__ flux, double* __restrict__ out) {
  double s0 = 0.0, s1 = 0.0;
  for (int f = threadIdx.x; f < nfe; f += kFnThreads) {
    const int2 ab = efverts[f];
    const double2 A = coords[ab.x], B = coords[ab.y], O = coords[efopp[f]];
    const double tx = B.x - A.x, ty = B.y - A.y;
    const double len = sqrt(tx * tx + ty * ty);
    double nx = ty / len, ny = -tx / len;
    if (nx * (O.x - A.x) + ny * (O.y - A.y) > 0.0) {  // the normal points away from the opposite vertex
      nx = -nx;
      ny = -ny;
    }
    // int_facet 4 pi r1^2 ds, exact for the linear r1 (FEniCS: 2-point Gauss, exact as well)
    const double wgt = kFourPi * len * (A.y * A.y + A.y * B.y + B.y * B.y) * (1.0 / 3.0);
    const double2 F = flux[efcell[f]];
    s0 += F.x * nx * wgt;
    s1 += F.y * ny * wgt;
  }
  block_reduce2(s0, s1);
  if (threadIdx.x == 0) {
    out[0] = s0;
    out[1] = s1;
    out[2] = s0 + s1;
  }
}

static int make_rule(int nq, const double* rule_host, RuleArg& R) {
  if (nq < 1 || nq > kMaxRule || rule_host == nullptr) return CM_E_INVALID;
  R.nq = nq;
  for (int k = 0; k < nq; ++k) {
    R.l1[k] = rule_host[3 * k];
    R.l2[k] = rule_host[3 * k + 1];
    R.w[k] = rule_host[3 * k + 2];
  }
  return CM_OK;
}

size_t functional_workspace_doubles() { return 4 + 2 * (size_t)kFnBlocks; }

int functional_error_norms(int nq, const double* rule_host, int weighted, int xonly, int ncells,
                           const int* cells, const double* coords, const double* uh, int stride,
                           const double* exq, const double* grq, double* work, cudaStream_t st) {
  RuleArg R;
  if (make_rule(nq, rule_host, R)) {
    set_error("cm_functional_error_norms: bad quadrature rule (1..%d points)", kMaxRule);
    return CM_E_INVALID;
  }
  error_norms_kernel<<<kFnBlocks, kFnThreads, 0, st>>>(R, weighted, xonly, ncells, cells,
                                                      (const double2*)coords, uh, stride, exq,
                                                      (const double2*)grq, work + 4);
  CM_LAUNCH_CHECK();
  final_sum_kernel<<<1, kFnThreads, 0, st>>>(kFnBlocks, 2, work + 4, work);
  CM_LAUNCH_CHECK();
  CM_BYTES((12.0 + 24.0 * nq) * ncells);
  return CM_OK;
}

int project_flux_dg0(int nq, const double* rule_host, double D1, double D2, int ncells, const int* cells,
                     const double* coords, const double* u, double* flux, cudaStream_t st) {
  RuleArg R;
  if (make_rule(nq, rule_host, R)) {
    set_error("cm_project_flux_dg0: bad quadrature rule (1..%d points)", kMaxRule);
    return CM_E_INVALID;
  }
  flux_dg0_kernel<<<(ncells + kFnThreads - 1) / kFnThreads, kFnThreads, 0, st>>>(
      R, D1, D2, ncells, cells, (const double2*)coords, (const double2*)u, (double2*)flux);
  CM_LAUNCH_CHECK();
  CM_BYTES(28.0 * ncells);
  return CM_OK;
}

int functional_boundary_flux(int nfe, const int* efverts, const int* efopp, const int* efcell,
                             const double* coords, const double* flux, double* out3, cudaStream_t st) {
  boundary_flux_kernel<<<1, kFnThreads, 0, st>>>(nfe, (const int2*)efverts, efopp, efcell,
                                                 (const double2*)coords, (const double2*)flux, out3);
  CM_LAUNCH_CHECK();
  return CM_OK;
}

}  // namespace cm
