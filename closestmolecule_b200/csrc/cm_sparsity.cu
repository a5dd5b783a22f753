// a9: sparsity pattern of the P1 space (DOLFIN SparsityPatternBuilder [ext], triggered by
// FunctionSpace(mesh, MixedElement([P0, P1])) at advection_diffusion_convergence.py:149-151 and the first
// assemble): node-level CSR with sorted columns = union over cells of vertex x vertex, plus the couples
// listed in `extra` (the two vertices opposite an interior facet, for the C0IP dS integrals).
// Keys row*nv + col -> radix sort -> unique (CUB device primitives) -> columns; the row pointer is a
// binary search per row in the sorted unique keys (deterministic, no atomics).
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>

#include "cm_common.cuh"
#include "cm_decl.cuh"

namespace cm {

__global__ void __launch_bounds__(256)
sparsity_keys_kernel(const long long nv, const int ncells, const int* __restrict__ cells, const int nextra,
                     const int* __restrict__ extra, unsigned long long* __restrict__ keys) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long ncell_keys = 9ll * ncells;
  if (t < ncell_keys) {
    const int c = (int)(t / 9), e = (int)(t % 9);
    const long long r = cells[3 * (size_t)c + e / 3], w = cells[3 * (size_t)c + e % 3];
    keys[t] = (unsigned long long)(r * nv + w);
  } else if (t < ncell_keys + 2ll * nextra) {
    const long long k = t - ncell_keys;
    const long long a = extra[2 * (k >> 1)], b = extra[2 * (k >> 1) + 1];
    keys[t] = (unsigned long long)((k & 1) ? b * nv + a : a * nv + b);
  }
}

__global__ void __launch_bounds__(256)
sparsity_cols_kernel(const long long nv, const long long nnz, const unsigned long long* __restrict__ keys,
                     int* __restrict__ col) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < nnz) col[k] = (int)(keys[k] % (unsigned long long)nv);
}

// rowptr[r] = first position whose key >= r*nv  (r = 0..nv)
__global__ void __launch_bounds__(256)
sparsity_rowptr_kernel(const long long nv, const long long nnz, const unsigned long long* __restrict__ keys,
                       int* __restrict__ rowptr) {
  const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > nv) return;
  const unsigned long long target = (unsigned long long)(r * nv);
  long long lo = 0, hi = nnz;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (keys[mid] < target) lo = mid + 1; else hi = mid;
  }
  rowptr[r] = (int)lo;
}

static int key_bits(long long nv) {
  int b = 1;
  while (b < 64 && (1ull << b) < (unsigned long long)nv * (unsigned long long)nv) ++b;
  return b;
}

struct SparsityLayout {
  size_t nkeys, off_a, off_b, off_cnt, off_tmp, tmp_bytes, total;
};

static SparsityLayout sparsity_layout(long long nv, int ncells, int nextra) {
  SparsityLayout L;
  L.nkeys = 9 * (size_t)ncells + 2 * (size_t)nextra;
  size_t t1 = 0, t2 = 0;
  cub::DeviceRadixSort::SortKeys(nullptr, t1, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                                 (long long)L.nkeys, 0, key_bits(nv));
  cub::DeviceSelect::Unique(nullptr, t2, (const unsigned long long*)nullptr, (unsigned long long*)nullptr,
                            (long long*)nullptr, (long long)L.nkeys);
  L.tmp_bytes = t1 > t2 ? t1 : t2;
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  size_t o = 0;
  L.off_a = o;   o += up(L.nkeys * 8);
  L.off_b = o;   o += up(L.nkeys * 8);
  L.off_cnt = o; o += 256;
  L.off_tmp = o; o += up(L.tmp_bytes);
  L.total = o;
  return L;
}

size_t p1_sparsity_workspace_bytes(int nv, int ncells, int nextra) {
  return sparsity_layout(nv, ncells, nextra).total;
}

int build_p1_sparsity(int nv, int ncells, const int* cells, int nextra, const int* extra, int* rowptr, int* col,
                      long long col_capacity, long long* nnz_host, void* ws, size_t ws_bytes, cudaStream_t st) {
  const SparsityLayout L = sparsity_layout(nv, ncells, nextra);
  if (ws_bytes < L.total) {
    set_error("cm_build_p1_sparsity: workspace too small (cm_p1_sparsity_workspace_bytes)");
    return CM_E_INVALID;
  }
  char* w = (char*)ws;
  unsigned long long* ka = (unsigned long long*)(w + L.off_a);
  unsigned long long* kb = (unsigned long long*)(w + L.off_b);
  long long* cnt = (long long*)(w + L.off_cnt);
  void* tmp = w + L.off_tmp;
  size_t tb = L.tmp_bytes;
  const long long nk = (long long)L.nkeys;
  sparsity_keys_kernel<<<(unsigned)((nk + 255) / 256), 256, 0, st>>>(nv, ncells, cells, nextra, extra, ka);
  CM_LAUNCH_CHECK();
  CM_CUDA(cub::DeviceRadixSort::SortKeys(tmp, tb, ka, kb, nk, 0, key_bits(nv), st));
  tb = L.tmp_bytes;
  CM_CUDA(cub::DeviceSelect::Unique(tmp, tb, kb, ka, cnt, nk, st));
  long long nnz = 0;
  CM_CUDA(cudaMemcpyAsync(&nnz, cnt, sizeof(long long), cudaMemcpyDeviceToHost, st));
  CM_CUDA(cudaStreamSynchronize(st));
  if (nnz_host) *nnz_host = nnz;
  if (nnz > col_capacity) {
    set_error("cm_build_p1_sparsity: %lld entries exceed the column capacity %lld", nnz, col_capacity);
    return CM_E_INVALID;
  }
  if (nnz >= (1ll << 31)) {
    set_error("cm_build_p1_sparsity: %lld entries exceed int32 (DOLFIN's la_index is 32 bit too)", nnz);
    return CM_E_UNSUPPORTED;
  }
  sparsity_cols_kernel<<<(unsigned)((nnz + 255) / 256), 256, 0, st>>>(nv, nnz, ka, col);
  CM_LAUNCH_CHECK();
  sparsity_rowptr_kernel<<<(unsigned)((nv + 1 + 255) / 256), 256, 0, st>>>(nv, nnz, ka, rowptr);
  CM_LAUNCH_CHECK();
  return CM_OK;
}

}  // namespace cm
