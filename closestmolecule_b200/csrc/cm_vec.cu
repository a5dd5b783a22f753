// Fused vector kernels of the Krylov solvers: grouped dot products (classical Gram-Schmidt in one
// pass over w), grouped axpy with the norm of the result, all with deterministic two-stage
// reductions (per-CTA partials in a fixed order, then one CTA) -- no atomics.
#include "cm_common.cuh"
#include "cm_decl.cuh"

namespace cm {

constexpr int kVecBlock = 256;
constexpr int kVecGrid = 148 * 4;  // persistent: 4 CTAs per SM
constexpr int kGroup = 8;

__device__ __forceinline__ double block_sum(double v, double* sh) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    r = (l < (blockDim.x >> 5)) ? sh[l] : 0.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
  }
  __syncthreads();
  return r;  // valid in thread 0
}

// partial[blockIdx.x * kGroup + k] = sum over this CTA's slice of V[k][i] * w[i]
template <int K>
__global__ void __launch_bounds__(kVecBlock)
dot_group_kernel(const long long n, const double* __restrict__ Vbase, const long long vstride,
                 const double* __restrict__ w, double* __restrict__ partial) {
  __shared__ double sh[32];
  double acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = 0.0;
  const long long n2 = n >> 1;
  const double2* w2 = reinterpret_cast<const double2*>(w);
  for (long long i = blockIdx.x * (long long)kVecBlock + threadIdx.x; i < n2;
       i += (long long)gridDim.x * kVecBlock) {
    const double2 ww = w2[i];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const double2 vv = reinterpret_cast<const double2*>(Vbase + k * vstride)[i];
      acc[k] += vv.x * ww.x + vv.y * ww.y;
    }
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] += Vbase[k * vstride + n - 1] * w[n - 1];
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const double s = block_sum(acc[k], sh);
    if (threadIdx.x == 0) partial[blockIdx.x * kGroup + k] = s;
  }
}

// out[k] = sum_b partial[b*kGroup + k]   (one CTA, fixed order)
__global__ void __launch_bounds__(kVecBlock)
reduce_partials_kernel(const int nblocks, const int K, const double* __restrict__ partial,
                       double* __restrict__ out) {
  __shared__ double sh[32];
  for (int k = 0; k < K; ++k) {
    double v = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += kVecBlock) v += partial[b * kGroup + k];
    const double s = block_sum(v, sh);
    if (threadIdx.x == 0) out[k] = s;
  }
}

// w -= sum_k h[k] * V[k]; if want_norm: partial[blockIdx.x*kGroup] = sum w_new^2 over the slice
template <int K>
__global__ void __launch_bounds__(kVecBlock)
axpy_group_kernel(const long long n, const double* __restrict__ Vbase, const long long vstride,
                  const double* __restrict__ h, const double sign, double* __restrict__ w,
                  double* __restrict__ partial, const int want_norm) {
  __shared__ double sh[32];
  double hk[K];
#pragma unroll
  for (int k = 0; k < K; ++k) hk[k] = sign * h[k];
  double nrm = 0.0;
  const long long n2 = n >> 1;
  double2* w2 = reinterpret_cast<double2*>(w);
  for (long long i = blockIdx.x * (long long)kVecBlock + threadIdx.x; i < n2;
       i += (long long)gridDim.x * kVecBlock) {
    double2 ww = w2[i];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const double2 vv = reinterpret_cast<const double2*>(Vbase + k * vstride)[i];
      ww.x += hk[k] * vv.x;
      ww.y += hk[k] * vv.y;
    }
    w2[i] = ww;
    nrm += ww.x * ww.x + ww.y * ww.y;
  }
  if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    double ww = w[n - 1];
#pragma unroll
    for (int k = 0; k < K; ++k) ww += hk[k] * Vbase[k * vstride + n - 1];
    w[n - 1] = ww;
    nrm += ww * ww;
  }
  if (want_norm) {
    const double s = block_sum(nrm, sh);
    if (threadIdx.x == 0) partial[blockIdx.x * kGroup] = s;
  }
}

// y = a*x (+ y_in*b)
__global__ void __launch_bounds__(kVecBlock)
scale_kernel(const long long n, const double* __restrict__ x, const double* __restrict__ a_dev,
             const double a_host, const int invert, double* __restrict__ y) {
  double a = a_dev ? *a_dev : a_host;
  if (invert) a = 1.0 / a;
  for (long long i = blockIdx.x * (long long)kVecBlock + threadIdx.x; i < n;
       i += (long long)gridDim.x * kVecBlock)
    y[i] = a * x[i];
}

static int vec_grid(long long n) {
  long long g = (n / 2 + kVecBlock - 1) / kVecBlock;
  if (g < 1) g = 1;
  return (int)(g < kVecGrid ? g : kVecGrid);
}

#define CM_K_SWITCH(K, STMT) \
  switch (K) {               \
    case 1: { constexpr int KK = 1; STMT; } break; \
    case 2: { constexpr int KK = 2; STMT; } break; \
    case 3: { constexpr int KK = 3; STMT; } break; \
    case 4: { constexpr int KK = 4; STMT; } break; \
    case 5: { constexpr int KK = 5; STMT; } break; \
    case 6: { constexpr int KK = 6; STMT; } break; \
    case 7: { constexpr int KK = 7; STMT; } break; \
    default: { constexpr int KK = 8; STMT; } break; \
  }

// out[0..nvec) = <V_k, w>, vectors V_k = Vbase + k*vstride.  scratch: kVecGrid*kGroup doubles.
int multi_dot(long long n, const double* Vbase, long long vstride, int nvec, const double* w,
              double* out, double* scratch, cudaStream_t st) {
  const int grid = vec_grid(n);
  for (int k0 = 0; k0 < nvec; k0 += kGroup) {
    const int K = (nvec - k0 < kGroup) ? nvec - k0 : kGroup;
    CM_K_SWITCH(K, (dot_group_kernel<KK><<<grid, kVecBlock, 0, st>>>(n, Vbase + k0 * vstride,
                                                                    vstride, w, scratch)));
    reduce_partials_kernel<<<1, kVecBlock, 0, st>>>(grid, K, scratch, out + k0);
    g_launches += 2;
    CM_BYTES((double)n * 8.0 * (K + 1));
  }
  --g_launches;
  CM_LAUNCH_CHECK();
  return CM_OK;
}

// w += sign * sum_k h[k] V_k ; if norm2_out: *norm2_out = ||w||^2 afterwards
int multi_axpy(long long n, const double* Vbase, long long vstride, int nvec, const double* h,
               double sign, double* w, double* norm2_out, double* scratch, cudaStream_t st) {
  const int grid = vec_grid(n);
  for (int k0 = 0; k0 < nvec; k0 += kGroup) {
    const int K = (nvec - k0 < kGroup) ? nvec - k0 : kGroup;
    const int last = (k0 + kGroup >= nvec) && norm2_out != nullptr;
    CM_K_SWITCH(K, (axpy_group_kernel<KK><<<grid, kVecBlock, 0, st>>>(
                       n, Vbase + k0 * vstride, vstride, h + k0, sign, w, scratch, last)));
    if (last) reduce_partials_kernel<<<1, kVecBlock, 0, st>>>(grid, 1, scratch, norm2_out);
    g_launches += last ? 2 : 1;
    CM_BYTES((double)n * 8.0 * (K + 2));
  }
  --g_launches;
  CM_LAUNCH_CHECK();
  return CM_OK;
}

// y = a*x with a on the device (a_dev != null) or host; invert: y = x/a
int vec_scale(long long n, const double* x, const double* a_dev, double a_host, int invert,
              double* y, cudaStream_t st) {
  long long g = (n + kVecBlock - 1) / kVecBlock;
  const int grid = (int)(g < kVecGrid ? g : kVecGrid);
  scale_kernel<<<grid, kVecBlock, 0, st>>>(n, x, a_dev, a_host, invert, y);
  CM_BYTES((double)n * 16.0);
  CM_LAUNCH_CHECK();
  return CM_OK;
}


// ------------------------------------------------------------------------------------------------
// slice variants for the strip-partitioned (multi-GPU) path: the vector is the union of two
// contiguous index ranges [off0, off0+len) and [off1, off1+len) (p-part and q-part rows of this
// rank in the split layout).  Scalar loads: the offsets need not be 16-byte aligned.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ long long slice_index(long long i, const VecSlice s) {
  return i < s.len ? s.off0 + i : s.off1 + (i - s.len);
}

template <int K>
__global__ void __launch_bounds__(kVecBlock)
dot_group_slice_kernel(const VecSlice sl, const double* __restrict__ Vbase, const long long vstride,
                       const double* __restrict__ w, double* __restrict__ partial) {
  __shared__ double sh[32];
  double acc[K];
#pragma unroll
  for (int k = 0; k < K; ++k) acc[k] = 0.0;
  const long long n = 2 * sl.len;
  for (long long i = blockIdx.x * (long long)kVecBlock + threadIdx.x; i < n;
       i += (long long)gridDim.x * kVecBlock) {
    const long long j = slice_index(i, sl);
    const double ww = w[j];
#pragma unroll
    for (int k = 0; k < K; ++k) acc[k] += Vbase[k * vstride + j] * ww;
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const double s = block_sum(acc[k], sh);
    if (threadIdx.x == 0) partial[blockIdx.x * kGroup + k] = s;
  }
}

template <int K>
__global__ void __launch_bounds__(kVecBlock)
axpy_group_slice_kernel(const VecSlice sl, const double* __restrict__ Vbase, const long long vstride,
                        const double* __restrict__ h, const double sign, double* __restrict__ w,
                        double* __restrict__ partial, const int want_norm) {
  __shared__ double sh[32];
  double hk[K];
#pragma unroll
  for (int k = 0; k < K; ++k) hk[k] = sign * h[k];
  double nrm = 0.0;
  const long long n = 2 * sl.len;
  for (long long i = blockIdx.x * (long long)kVecBlock + threadIdx.x; i < n;
       i += (long long)gridDim.x * kVecBlock) {
    const long long j = slice_index(i, sl);
    double ww = w[j];
#pragma unroll
    for (int k = 0; k < K; ++k) ww += hk[k] * Vbase[k * vstride + j];
    w[j] = ww;
    nrm += ww * ww;
  }
  if (want_norm) {
    const double s = block_sum(nrm, sh);
    if (threadIdx.x == 0) partial[blockIdx.x * kGroup] = s;
  }
}

// y = a*x on the slices (a on the host; invert: y = x/a); x == nullptr: y = 0
__global__ void __launch_bounds__(kVecBlock)
scale_slice_kernel(const VecSlice sl, const double* __restrict__ x, const double a,
                   double* __restrict__ y) {
  const long long n = 2 * sl.len;
  for (long long i = blockIdx.x * (long long)kVecBlock + threadIdx.x; i < n;
       i += (long long)gridDim.x * kVecBlock) {
    const long long j = slice_index(i, sl);
    y[j] = x ? a * x[j] : 0.0;
  }
}

static int slice_grid(const VecSlice sl) {
  long long g = (2 * sl.len + kVecBlock - 1) / kVecBlock;
  if (g < 1) g = 1;
  return (int)(g < kVecGrid ? g : kVecGrid);
}

int multi_dot_slice(VecSlice sl, const double* Vbase, long long vstride, int nvec, const double* w,
                    double* out, double* scratch, cudaStream_t st) {
  const int grid = slice_grid(sl);
  for (int k0 = 0; k0 < nvec; k0 += kGroup) {
    const int K = (nvec - k0 < kGroup) ? nvec - k0 : kGroup;
    CM_K_SWITCH(K, (dot_group_slice_kernel<KK><<<grid, kVecBlock, 0, st>>>(sl, Vbase + k0 * vstride,
                                                                          vstride, w, scratch)));
    reduce_partials_kernel<<<1, kVecBlock, 0, st>>>(grid, K, scratch, out + k0);
    g_launches += 2;
    CM_BYTES((double)sl.len * 16.0 * (K + 1));
  }
  --g_launches;
  CM_LAUNCH_CHECK();
  return CM_OK;
}

int multi_axpy_slice(VecSlice sl, const double* Vbase, long long vstride, int nvec, const double* h,
                     double sign, double* w, double* norm2_out, double* scratch, cudaStream_t st) {
  const int grid = slice_grid(sl);
  for (int k0 = 0; k0 < nvec; k0 += kGroup) {
    const int K = (nvec - k0 < kGroup) ? nvec - k0 : kGroup;
    const int last = (k0 + kGroup >= nvec) && norm2_out != nullptr;
    CM_K_SWITCH(K, (axpy_group_slice_kernel<KK><<<grid, kVecBlock, 0, st>>>(
                       sl, Vbase + k0 * vstride, vstride, h + k0, sign, w, scratch, last)));
    if (last) reduce_partials_kernel<<<1, kVecBlock, 0, st>>>(grid, 1, scratch, norm2_out);
    g_launches += last ? 2 : 1;
    CM_BYTES((double)sl.len * 16.0 * (K + 2));
  }
  --g_launches;
  CM_LAUNCH_CHECK();
  return CM_OK;
}

int vec_scale_slice(VecSlice sl, const double* x, double a, int invert, double* y, cudaStream_t st) {
  scale_slice_kernel<<<slice_grid(sl), kVecBlock, 0, st>>>(sl, x, invert ? 1.0 / a : a, y);
  CM_BYTES((double)sl.len * (x ? 32.0 : 16.0));
  CM_LAUNCH_CHECK();
  return CM_OK;
}

size_t vec_scratch_doubles() { return (size_t)kVecGrid * kGroup; }

}  // namespace cm
