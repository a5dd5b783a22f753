// Multi-GPU plumbing of the strip-partitioned Newton solve: one process per GPU, every rank holds
// GLOBALLY indexed arrays inside one IPC-shared workspace with the same layout on all ranks and
// owns a contiguous range of vertex rows.  Halo exchange, all-gather and the Krylov all-reduces are
// hand-written peer-to-peer kernels over NVLink: the producer writes its boundary rows straight
// into the neighbour's copy of the same array (same offset), fences, then bumps a sequence flag in
// the neighbour's memory; the consumer stream runs a one-warp wait kernel on its own flags.
// Sequence numbers only grow, and the same ghost row is never re-written without a completed
// two-way exchange in between, so no extra acknowledgement is needed.
#include "cm_common.cuh"
#include "cm_decl.cuh"

namespace cm {

constexpr int kMaxPeers = 8;
constexpr int kCollMax = 64;  // doubles per all-reduce

// layout of the control block at the start of every rank's workspace (doubles / u64 units)
//   [0, 8)        halo flags, one per source rank
//   [8, 16)       collective flags, one per source rank
//   [16, 16 + 2*8*64)  all-reduce slots [parity][src][kCollMax]
constexpr size_t kCtlHaloFlags = 0;
constexpr size_t kCtlCollFlags = 8;
constexpr size_t kCtlCollBuf = 16;
constexpr size_t kCtlDoubles = 16 + 2 * kMaxPeers * kCollMax + 16;

size_t dist_ctl_doubles() { return kCtlDoubles; }

struct PushSegs {
  // up to 8 contiguous segments (offsets in doubles from the workspace base, identical on all ranks)
  unsigned long long off[8];
  int n[8];
  int nseg;
};

struct Peers {
  double* base[kMaxPeers];
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// grid = number of destination peers; CTA k copies every segment from my workspace to peer
// dst[k]'s workspace at the same offsets, then publishes `seq` in that peer's flag slot [my_rank].
__global__ void __launch_bounds__(1024)
push_kernel(const double* __restrict__ my_base, const Peers peers, const int my_rank,
            const int dst0, const int dst1, const int dst_all, const int world, const PushSegs segs,
            const size_t flag_off, const unsigned long long seq) {
  int dst;
  if (dst_all) {
    dst = blockIdx.x + (blockIdx.x >= my_rank ? 1 : 0);
    if (dst >= world) return;
  } else {
    dst = blockIdx.x == 0 ? dst0 : dst1;
    if (dst < 0) {  // only one neighbour
      dst = dst1;
      if (blockIdx.x == 1 || dst < 0) return;
    }
  }
  double* __restrict__ pb = peers.base[dst];
  for (int s = 0; s < segs.nseg; ++s) {
    const double* src = my_base + segs.off[s];
    double* d = pb + segs.off[s];
    for (int i = threadIdx.x; i < segs.n[s]; i += blockDim.x) d[i] = src[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0)
    st_release_sys(reinterpret_cast<unsigned long long*>(pb) + flag_off + my_rank, seq);
}

// halo exchange in ONE launch: CTA 0 serves the lower neighbour, CTA 1 the upper one.  Each CTA
// pushes its segments into the neighbour's workspace, publishes `seq` there, then waits until the
// same neighbour has published `seq` here.
__global__ void __launch_bounds__(1024)
exchange_kernel(const double* __restrict__ my_base, const Peers peers, const int my_rank, const int dn,
                const int up, const PushSegs segs_dn, const PushSegs segs_up, const size_t flag_off,
                const unsigned long long seq) {
  const int dst = blockIdx.x == 0 ? dn : up;
  if (dst < 0) return;
  const PushSegs& segs = blockIdx.x == 0 ? segs_dn : segs_up;
  double* __restrict__ pb = peers.base[dst];
  for (int s = 0; s < segs.nseg; ++s) {
    const double* src = my_base + segs.off[s];
    double* d = pb + segs.off[s];
    for (int i = threadIdx.x; i < segs.n[s]; i += blockDim.x) d[i] = src[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    st_release_sys(reinterpret_cast<unsigned long long*>(pb) + flag_off + my_rank, seq);
    const unsigned long long* f = reinterpret_cast<const unsigned long long*>(my_base) + flag_off + dst;
    while (ld_acquire_sys(f) < seq) __nanosleep(40);
  }
}

// one warp: lane s waits until flag[s] >= seq for every source in the mask
__global__ void wait_kernel(const double* my_base, const size_t flag_off, const unsigned mask,
                            const unsigned long long seq) {
  const int s = threadIdx.x;
  if (s < kMaxPeers && (mask & (1u << s))) {
    const unsigned long long* f = reinterpret_cast<const unsigned long long*>(my_base) + flag_off + s;
    while (ld_acquire_sys(f) < seq) __nanosleep(100);
  }
}

// all-reduce (sum) of n <= 64 doubles in ONE launch: CTA d publishes my values in rank d's slot
// (CTA my_rank writes the local slot); CTA 0 then waits for every rank's flag and sums in rank order.
__global__ void __launch_bounds__(64)
allreduce_kernel(double* __restrict__ my_base, const Peers peers, const int my_rank, const int world,
                 double* __restrict__ vals, const int n, const int parity, const unsigned long long seq) {
  __shared__ double mine[kCollMax];
  const int dst = blockIdx.x;
  if (threadIdx.x < n) mine[threadIdx.x] = vals[threadIdx.x];
  __syncthreads();
  // CTA 0 overwrites vals with the sum: it may do so only after every CTA of this launch has read it
  unsigned long long* readers = reinterpret_cast<unsigned long long*>(my_base) + (kCtlDoubles - 8);
  if (threadIdx.x == 0) atomicAdd(readers, 1ull);
  double* pb = (dst == my_rank) ? my_base : peers.base[dst];
  double* slot = pb + kCtlCollBuf + ((size_t)parity * kMaxPeers + my_rank) * kCollMax;
  if (threadIdx.x < n) slot[threadIdx.x] = mine[threadIdx.x];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0)
    st_release_sys(reinterpret_cast<unsigned long long*>(pb) + kCtlCollFlags + my_rank, seq);
  if (dst != 0) return;
  if (threadIdx.x < world) {
    const unsigned long long* f =
        reinterpret_cast<const unsigned long long*>(my_base) + kCtlCollFlags + threadIdx.x;
    while (ld_acquire_sys(f) < seq) __nanosleep(40);
  }
  if (threadIdx.x == 0)
    while (atomicAdd(readers, 0ull) < seq * (unsigned long long)world) __nanosleep(20);
  __syncthreads();
  if (threadIdx.x < n) {
    double s = 0.0;
    for (int r = 0; r < world; ++r)
      s += my_base[kCtlCollBuf + ((size_t)parity * kMaxPeers + r) * kCollMax + threadIdx.x];
    vals[threadIdx.x] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static Peers make_peers(const cm_dist* D) {
  Peers p;
  for (int i = 0; i < kMaxPeers; ++i) p.base[i] = (i < D->world) ? (double*)D->peer_base[i] : nullptr;
  return p;
}

// copy the given segments to the row neighbours (rank-1 and rank+1) and wait for theirs
int dist_halo(cm_dist* D, const unsigned long long* offs, const int* ns, int nseg, cudaStream_t st) {
  if (D == nullptr || D->world <= 1) return CM_OK;
  PushSegs segs_dn{}, segs_up{};
  (void)segs_up;
  if (nseg > 8) return CM_E_INVALID;
  segs_dn.nseg = nseg;
  for (int i = 0; i < nseg; ++i) { segs_dn.off[i] = offs[i]; segs_dn.n[i] = ns[i]; }
  const unsigned long long seq = ++D->halo_seq;
  const int dn = D->rank > 0 ? D->rank - 1 : -1, up = D->rank + 1 < D->world ? D->rank + 1 : -1;
  push_kernel<<<2, 1024, 0, st>>>((const double*)D->peer_base[D->rank], make_peers(D), D->rank, dn, up,
                                  0, D->world, segs_dn, kCtlHaloFlags, seq);
  unsigned mask = 0;
  if (dn >= 0) mask |= 1u << dn;
  if (up >= 0) mask |= 1u << up;
  wait_kernel<<<1, 32, 0, st>>>((const double*)D->peer_base[D->rank], kCtlHaloFlags, mask, seq);
  g_launches += 1;
  CM_LAUNCH_CHECK();
  return CM_OK;
}

// halo exchange where the segments sent downwards and upwards differ (my first / my last rows)
int dist_halo2(cm_dist* D, const unsigned long long* offs_dn, const int* ns_dn, int nseg_dn,
               const unsigned long long* offs_up, const int* ns_up, int nseg_up, cudaStream_t st) {
  if (D == nullptr || D->world <= 1) return CM_OK;
  if (nseg_dn > 8 || nseg_up > 8) return CM_E_INVALID;
  const unsigned long long seq = ++D->halo_seq;
  const int dn = D->rank > 0 ? D->rank - 1 : -1, up = D->rank + 1 < D->world ? D->rank + 1 : -1;
  PushSegs sd{}, su{};
  sd.nseg = nseg_dn;
  for (int i = 0; i < nseg_dn; ++i) { sd.off[i] = offs_dn[i]; sd.n[i] = ns_dn[i]; }
  su.nseg = nseg_up;
  for (int i = 0; i < nseg_up; ++i) { su.off[i] = offs_up[i]; su.n[i] = ns_up[i]; }
  exchange_kernel<<<2, 1024, 0, st>>>((const double*)D->peer_base[D->rank], make_peers(D), D->rank, dn, up,
                                      sd, su, kCtlHaloFlags, seq);
  CM_LAUNCH_CHECK();
  return CM_OK;
}

// every rank sends the same-offset segments to ALL peers (all-gather of owned rows)
int dist_allgather(cm_dist* D, const unsigned long long* offs, const int* ns, int nseg,
                   cudaStream_t st) {
  if (D == nullptr || D->world <= 1) return CM_OK;
  if (nseg > 8) return CM_E_INVALID;
  PushSegs s{};
  s.nseg = nseg;
  for (int i = 0; i < nseg; ++i) { s.off[i] = offs[i]; s.n[i] = ns[i]; }
  const unsigned long long seq = ++D->halo_seq;
  const double* mb = (const double*)D->peer_base[D->rank];
  push_kernel<<<D->world - 1, 1024, 0, st>>>(mb, make_peers(D), D->rank, -1, -1, 1, D->world, s,
                                             kCtlHaloFlags, seq);
  const unsigned mask = ((1u << D->world) - 1u) & ~(1u << D->rank);
  wait_kernel<<<1, 32, 0, st>>>(mb, kCtlHaloFlags, mask, seq);
  ++g_launches;
  CM_LAUNCH_CHECK();
  return CM_OK;
}

// in-place sum over ranks of n <= 64 device doubles (deterministic: fixed rank order)
int dist_allreduce(cm_dist* D, double* vals, int n, cudaStream_t st) {
  if (D == nullptr || D->world <= 1) return CM_OK;
  if (n > kCollMax) return CM_E_INVALID;
  const unsigned long long seq = ++D->coll_seq;
  const int parity = (int)(seq & 1ull);
  double* mb = (double*)D->peer_base[D->rank];
  allreduce_kernel<<<D->world, 64, 0, st>>>(mb, make_peers(D), D->rank, D->world, vals, n, parity, seq);
  CM_LAUNCH_CHECK();
  return CM_OK;
}

}  // namespace cm
