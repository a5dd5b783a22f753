// SpMV on the node-level pattern.  bs=2: the DOLFIN-interleaved P-Q value layout (see the header);
// bs=1: scalar node CSR.  A group of LPN lanes owns one node row: value and column loads of the four
// (resp. one) scalar rows of consecutive nodes are contiguous, partial sums meet by shuffles.
#include "cm_common.cuh"
#include "cm_decl.cuh"

namespace cm {

template <int LPN>
__global__ void __launch_bounds__(256)
spmv_pq_kernel(const int nv, const int* __restrict__ nrowptr, const int* __restrict__ ncol,
               const double2* __restrict__ vals2, const double2* __restrict__ x,
               double2* __restrict__ y, const double alpha, const double2* __restrict__ z) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int v = gid / LPN;
  const int lane = gid % LPN;
  double sp = 0.0, sq = 0.0;
  if (v < nv) {
    const int rp = nrowptr[v];
    const int deg = nrowptr[v + 1] - rp;
    const double2* prow = vals2 + 2 * (size_t)rp;  // 4*rp doubles
    const double2* qrow = prow + deg;
    for (int k = lane; k < deg; k += LPN) {
      const double2 xv = x[ncol[rp + k]];
      const double2 a = prow[k], c = qrow[k];
      sp += a.x * xv.x + a.y * xv.y;
      sq += c.x * xv.x + c.y * xv.y;
    }
  }
#pragma unroll
  for (int o = LPN / 2; o > 0; o >>= 1) {
    sp += __shfl_down_sync(0xffffffffu, sp, o, LPN);
    sq += __shfl_down_sync(0xffffffffu, sq, o, LPN);
  }
  if (v < nv && lane == 0) {
    double2 r = make_double2(sp, sq);
    if (z != nullptr) {  // y = z + alpha*A*x
      const double2 zz = z[v];
      r.x = zz.x + alpha * r.x;
      r.y = zz.y + alpha * r.y;
    }
    y[v] = r;
  }
}

template <int LPN>
__global__ void __launch_bounds__(256)
spmv_scalar_kernel(const int nv, const int* __restrict__ nrowptr, const int* __restrict__ ncol,
                   const double* __restrict__ vals, const double* __restrict__ x,
                   double* __restrict__ y, const double alpha, const double* __restrict__ z) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int v = gid / LPN;
  const int lane = gid % LPN;
  double s = 0.0;
  if (v < nv) {
    const int rp = nrowptr[v];
    const int e = nrowptr[v + 1];
    for (int k = rp + lane; k < e; k += LPN) s += vals[k] * x[ncol[k]];
  }
#pragma unroll
  for (int o = LPN / 2; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o, LPN);
  if (v < nv && lane == 0) y[v] = (z != nullptr) ? z[v] + alpha * s : s;
}

int spmv(int bs, int nv, const int* nrowptr, const int* ncol, const double* vals, const double* x,
         double* y, double alpha, const double* z, cudaStream_t st) {
  constexpr int LPN = 8;
  const long long threads = (long long)nv * LPN;
  const int grid = (int)((threads + 255) / 256);
  if (bs == 2) {
    spmv_pq_kernel<LPN><<<grid, 256, 0, st>>>(nv, nrowptr, ncol, (const double2*)vals,
                                             (const double2*)x, (double2*)y, alpha,
                                             (const double2*)z);
  } else {
    spmv_scalar_kernel<LPN><<<grid, 256, 0, st>>>(nv, nrowptr, ncol, vals, x, y, alpha, z);
  }
  CM_LAUNCH_CHECK();
  return CM_OK;
}

}  // namespace cm
