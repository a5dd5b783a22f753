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

// layout of the control block at the start of every rank's workspace (u64 / double units).  All
// sequence counters live on the DEVICE (each kernel bumps its own), so launches carry no changing
// arguments and whole preconditioner applications can be replayed as CUDA graphs.
//   [0]  flag written by the lower neighbour      [1]  flag written by the upper neighbour
//   [2]  my exchange counter towards below        [3]  towards above
//   [4]  all-gather counter  [5] all-reduce counter  [6] all-reduce local reader count
//   [8, 16)   all-gather flags, one per source rank
//   [16, 24)  all-reduce flags, one per source rank
//   [32, 32 + 2*8*64)  all-reduce slots [parity][src][kCollMax]
constexpr size_t kFlagFromBelow = 0, kFlagFromAbove = 1, kCntDn = 2, kCntUp = 3, kCntAg = 4, kCntAr = 5,
                 kArReaders = 6;
constexpr size_t kCtlAgFlags = 8;
constexpr size_t kCtlCollFlags = 16;
constexpr size_t kCtlCollBuf = 32;
constexpr size_t kCtlDoubles = 32 + 2 * kMaxPeers * kCollMax + 16;

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
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// wait until *flag >= seq; a peer that never arrives (crashed rank) must not hang the GPU for ever:
// after 60 s the kernel traps, which surfaces as a CUDA error in the host call chain
__device__ __forceinline__ void spin_until(const unsigned long long* flag, const unsigned long long seq) {
  const unsigned long long t0 = global_ns();
  while (ld_acquire_sys(flag) < seq) {
    __nanosleep(40);
    if (global_ns() - t0 > 60000000000ull) asm volatile("trap;");
  }
}

// halo exchange in ONE launch: CTA 0 serves the lower neighbour, CTA 1 the upper one.  Each CTA takes
// the next sequence number of its direction, pushes its segments into the neighbour's workspace,
// publishes the number there, then waits until the same neighbour has published it here.
__global__ void __launch_bounds__(1024)
exchange_kernel(double* __restrict__ my_base, const Peers peers, const int dn, const int up,
                const PushSegs segs_dn, const PushSegs segs_up) {
  __shared__ unsigned long long s_seq;
  const bool down = blockIdx.x == 0;
  const int dst = down ? dn : up;
  if (dst < 0) return;
  unsigned long long* ctl = reinterpret_cast<unsigned long long*>(my_base);
  if (threadIdx.x == 0) s_seq = ++ctl[down ? kCntDn : kCntUp];
  const PushSegs& segs = down ? segs_dn : segs_up;
  double* __restrict__ pb = peers.base[dst];
  for (int s = 0; s < segs.nseg; ++s) {
    const double* src = my_base + segs.off[s];
    double* d = pb + segs.off[s];
    for (int i = threadIdx.x; i < segs.n[s]; i += blockDim.x) d[i] = src[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long seq = s_seq;
    // I am the neighbour's upper (resp. lower) side
    st_release_sys(reinterpret_cast<unsigned long long*>(pb) + (down ? kFlagFromAbove : kFlagFromBelow), seq);
    const unsigned long long* f = ctl + (down ? kFlagFromBelow : kFlagFromAbove);
    spin_until(f, seq);
  }
}

// all-gather of same-offset segments: CTA k pushes to the k-th other rank; CTA 0 then waits for all
__global__ void __launch_bounds__(1024)
allgather_kernel(double* __restrict__ my_base, const Peers peers, const int my_rank, const int world,
                 const PushSegs segs) {
  __shared__ unsigned long long s_seq;
  unsigned long long* ctl = reinterpret_cast<unsigned long long*>(my_base);
  const int dst = blockIdx.x + (blockIdx.x >= my_rank ? 1 : 0);
  if (dst >= world) return;
  // every CTA needs the same number: CTA 0 bumps the counter, the others read counter + 1 ... avoid the
  // race by deriving it from a value only CTA 0 modifies AFTER all pushes: use a per-launch snapshot
  if (threadIdx.x == 0) s_seq = ld_acquire_sys(ctl + kCntAg) + 1ull;
  __syncthreads();
  const unsigned long long seq = s_seq;
  double* __restrict__ pb = peers.base[dst];
  for (int s = 0; s < segs.nseg; ++s) {
    const double* src = my_base + segs.off[s];
    double* d = pb + segs.off[s];
    for (int i = threadIdx.x; i < segs.n[s]; i += blockDim.x) d[i] = src[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0)
    st_release_sys(reinterpret_cast<unsigned long long*>(pb) + kCtlAgFlags + my_rank, seq);
  if (blockIdx.x != 0) return;
  if (threadIdx.x < world && threadIdx.x != my_rank) {
    const unsigned long long* f = ctl + kCtlAgFlags + threadIdx.x;
    spin_until(f, seq);
  }
}
// the counter is advanced by a trailing one-thread kernel (stream order: after every CTA of the
// all-gather launch has taken its snapshot and finished)
__global__ void bump_kernel(double* my_base, const size_t idx) {
  ++reinterpret_cast<unsigned long long*>(my_base)[idx];
}

// all-reduce (sum) of n <= 64 doubles in ONE launch: CTA d publishes my values in rank d's slot
// (CTA my_rank writes the local slot); CTA 0 then waits for every rank's flag and sums in rank order.
__global__ void __launch_bounds__(64)
allreduce_kernel(double* __restrict__ my_base, const Peers peers, const int my_rank, const int world,
                 double* __restrict__ vals, const int n) {
  __shared__ double mine[kCollMax];
  __shared__ unsigned long long s_seq;
  unsigned long long* ctl = reinterpret_cast<unsigned long long*>(my_base);
  const int dst = blockIdx.x;
  if (threadIdx.x == 0) s_seq = ld_acquire_sys(ctl + kCntAr) + 1ull;
  if (threadIdx.x < n) mine[threadIdx.x] = vals[threadIdx.x];
  __syncthreads();
  const unsigned long long seq = s_seq;
  const int parity = (int)(seq & 1ull);
  // CTA 0 overwrites vals with the sum: only after every CTA of this launch has read it
  if (threadIdx.x == 0) atomicAdd(ctl + kArReaders, 1ull);
  double* pb = (dst == my_rank) ? my_base : peers.base[dst];
  double* slot = pb + kCtlCollBuf + ((size_t)parity * kMaxPeers + my_rank) * kCollMax;
  if (threadIdx.x < n) slot[threadIdx.x] = mine[threadIdx.x];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0)
    st_release_sys(reinterpret_cast<unsigned long long*>(pb) + kCtlCollFlags + my_rank, seq);
  if (dst != 0) return;
  if (threadIdx.x < world) {
    const unsigned long long* f = ctl + kCtlCollFlags + threadIdx.x;
    spin_until(f, seq);
  }
  if (threadIdx.x == 0)
    while (atomicAdd(ctl + kArReaders, 0ull) < seq * (unsigned long long)world) __nanosleep(20);
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

// halo exchange: segments sent downwards / upwards (my first / my last rows), wait for the neighbours'
int dist_halo2(cm_dist* D, const unsigned long long* offs_dn, const int* ns_dn, int nseg_dn,
               const unsigned long long* offs_up, const int* ns_up, int nseg_up, cudaStream_t st) {
  if (D == nullptr || D->world <= 1) return CM_OK;
  if (nseg_dn > 8 || nseg_up > 8) return CM_E_INVALID;
  const int dn = D->rank > 0 ? D->rank - 1 : -1, up = D->rank + 1 < D->world ? D->rank + 1 : -1;
  PushSegs sd{}, su{};
  sd.nseg = nseg_dn;
  for (int i = 0; i < nseg_dn; ++i) { sd.off[i] = offs_dn[i]; sd.n[i] = ns_dn[i]; }
  su.nseg = nseg_up;
  for (int i = 0; i < nseg_up; ++i) { su.off[i] = offs_up[i]; su.n[i] = ns_up[i]; }
  exchange_kernel<<<2, 1024, 0, st>>>((double*)D->peer_base[D->rank], make_peers(D), dn, up, sd, su);
  CM_LAUNCH_CHECK();
  return CM_OK;
}

int dist_halo(cm_dist* D, const unsigned long long* offs, const int* ns, int nseg, cudaStream_t st) {
  return dist_halo2(D, offs, ns, nseg, offs, ns, nseg, st);
}

// every rank sends the same-offset segments to ALL peers (all-gather of owned rows)
int dist_allgather(cm_dist* D, const unsigned long long* offs, const int* ns, int nseg,
                   cudaStream_t st) {
  if (D == nullptr || D->world <= 1) return CM_OK;
  if (nseg > 8) return CM_E_INVALID;
  PushSegs s{};
  s.nseg = nseg;
  for (int i = 0; i < nseg; ++i) { s.off[i] = offs[i]; s.n[i] = ns[i]; }
  double* mb = (double*)D->peer_base[D->rank];
  allgather_kernel<<<D->world - 1, 1024, 0, st>>>(mb, make_peers(D), D->rank, D->world, s);
  bump_kernel<<<1, 1, 0, st>>>(mb, kCntAg);
  ++g_launches;
  CM_LAUNCH_CHECK();
  return CM_OK;
}

// in-place sum over ranks of n <= 64 device doubles (deterministic: fixed rank order)
int dist_allreduce(cm_dist* D, double* vals, int n, cudaStream_t st) {
  if (D == nullptr || D->world <= 1) return CM_OK;
  if (n > kCollMax) return CM_E_INVALID;
  double* mb = (double*)D->peer_base[D->rank];
  allreduce_kernel<<<D->world, 64, 0, st>>>(mb, make_peers(D), D->rank, D->world, vals, n);
  bump_kernel<<<1, 1, 0, st>>>(mb, kCntAr);
  ++g_launches;
  CM_LAUNCH_CHECK();
  return CM_OK;
}

}  // namespace cm
