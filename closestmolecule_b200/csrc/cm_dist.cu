// Multi-GPU plumbing of the strip-partitioned Newton solve: one process per GPU, every rank holds
// GLOBALLY indexed arrays inside one IPC-shared workspace with the same layout on all ranks and
// owns a contiguous range of vertex rows.  Halo exchange, all-gather and the Krylov all-reduces are
// hand-written peer-to-peer kernels over NVLink: the producer writes its boundary rows straight
// into the neighbour's copy of the same array (same offset), fences, then bumps a sequence flag in
// the neighbour's memory; the consumer stream runs a one-warp wait kernel on its own flags.
// Sequence numbers only grow.  There is no consumer acknowledgement; write-after-read safety rests on two
// invariants that the host side ENFORCES (cm_pqs_create rejects / clamps anything else):
//   * halo rows: strips start on even rows of every distributed level and keep >= 4 rows down to the first
//     replicated level, so (zebra colouring, cm_halo.cuh) every re-push of a ghost row is preceded in the
//     pusher's program order by a wait on a push the reader issued after its last read of the old value;
//   * all-gather targets / all-reduce slots: between two all-gathers of the same array every rank passes at
//     least one all-reduce (each GMRES iteration has two; the setup all-gather is followed by the solve), and an
//     all-reduce completes on a rank only after EVERY rank has contributed, i.e. has finished reading what the
//     previous all-gather delivered; the all-reduce slots themselves alternate by sequence parity.
#include "cm_common.cuh"
#include "cm_decl.cuh"
#include "cm_halo.cuh"

namespace cm {

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

// explicit halo exchange in ONE launch (used where no compute kernel produces the rows: Newton state,
// operator diagonals at setup, restart vectors): CTA 0 pushes the `down` segments to the lower neighbour
// and waits for the matching down-push of the upper neighbour; CTA 1 pushes up and waits for the lower
// neighbour's up-push.  Same counters / flags as the inline pushes of the compute kernels (cm_halo.cuh).
__global__ void __launch_bounds__(1024)
exchange_kernel(const Halo h, const PushSegs segs_dn, const PushSegs segs_up) {
  const bool down = blockIdx.x == 0;
  double* __restrict__ pb = down ? h.dn_base : h.up_base;
  const PushSegs& segs = down ? segs_dn : segs_up;
  if (pb != nullptr) {
    for (int s = 0; s < segs.nseg; ++s) {
      const double* src = h.my_base + segs.off[s];
      double* d = pb + segs.off[s];
      for (int i = threadIdx.x; i < segs.n[s]; i += blockDim.x) d[i] = src[i];
    }
  }
  halo_publish(h, down);   // fence, barrier, bump my counter, flag in the neighbour's block
  if (threadIdx.x == 0) {
    if (down) halo_wait_above(h);   // reads cnt_dn, which only this CTA bumps
    else halo_wait_below(h);
  }
}

// all-gather of same-offset segments: kAgSplit CTAs share the push to each other rank (a single CTA moves
// ~20 GB/s over NVLink: the 2 MB level-2 right-hand side of a 2-GPU run took 116 us that way, in-situ timeline
// profiles/r2_trace_n2.csv); the last of them to finish publishes the flag in the destination's control block.
// CTA 0 then waits for the flags of all ranks.  Every CTA takes the launch's sequence number from a snapshot
// of the counter; the counter itself is advanced by the last CTA of the launch to finish.
constexpr size_t kAgDone = 19;
constexpr size_t kAgPart = 40;    // [40, 48): CTAs finished per destination (cumulative)
constexpr int kAgSplit = 16;
__global__ void __launch_bounds__(1024)
allgather_kernel(double* __restrict__ my_base, const Peers peers, const int my_rank, const int world,
                 const PushSegs segs) {
  __shared__ unsigned long long s_seq;
  unsigned long long* ctl = reinterpret_cast<unsigned long long*>(my_base);
  const int di = blockIdx.x / kAgSplit, part = blockIdx.x - di * kAgSplit;
  const int dst = di + (di >= my_rank ? 1 : 0);
  if (threadIdx.x == 0) s_seq = ld_volatile_u64(ctl + kCntAg) + 1ull;
  __syncthreads();
  const unsigned long long seq = s_seq;
  double* __restrict__ pb = peers.base[dst];
  for (int s = 0; s < segs.nseg; ++s) {
    const int chunk = (segs.n[s] + kAgSplit - 1) / kAgSplit;
    const int i0 = part * chunk, i1 = min(i0 + chunk, segs.n[s]);
    const double* src = my_base + segs.off[s];
    double* d = pb + segs.off[s];
    for (int i = i0 + threadIdx.x; i < i1; i += blockDim.x) d[i] = src[i];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long fin = atomicAdd(ctl + kAgPart + di, 1ull) + 1ull;
    if (fin == seq * (unsigned long long)kAgSplit) {   // every part of this destination is stored and fenced
      __threadfence_system();
      st_release_sys(reinterpret_cast<unsigned long long*>(pb) + kCtlAgFlags + my_rank, seq);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < world && threadIdx.x != my_rank) {
    const unsigned long long* f = ctl + kCtlAgFlags + threadIdx.x;
    spin_until(f, seq);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned long long done = atomicAdd(ctl + kAgDone, 1ull) + 1ull;
    if (done == seq * (unsigned long long)gridDim.x) ctl[kCntAg] = seq;   // last CTA of this launch
  }
}

// all-reduce (sum) of n <= 64 doubles in ONE launch: CTA d publishes my values in rank d's slot
// (CTA my_rank writes the local slot); CTA 0 then waits for every rank's flag, sums in rank order and
// advances the sequence counter (every CTA of the launch has taken its snapshot by then: kArReaders).
__global__ void __launch_bounds__(64)
allreduce_kernel(double* __restrict__ my_base, const Peers peers, const int my_rank, const int world,
                 double* __restrict__ vals, const int n) {
  __shared__ double mine[kCollMax];
  __shared__ unsigned long long s_seq;
  unsigned long long* ctl = reinterpret_cast<unsigned long long*>(my_base);
  const int dst = blockIdx.x;
  if (threadIdx.x == 0) s_seq = ld_volatile_u64(ctl + kCntAr) + 1ull;
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
  if (threadIdx.x == 0) {
    while (atomicAdd(ctl + kArReaders, 0ull) < seq * (unsigned long long)world) __nanosleep(20);
    ctl[kCntAr] = seq;
  }
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

// kernel argument describing this rank's strip neighbours (channel `chan` of the control block)
Halo dist_halo_args(const cm_dist* D, int chan) {
  Halo h{};
  if (D == nullptr || D->world <= 1) return h;
  h.my_base = (double*)D->peer_base[D->rank];
  h.dn_base = D->rank > 0 ? (double*)D->peer_base[D->rank - 1] : nullptr;
  h.up_base = D->rank + 1 < D->world ? (double*)D->peer_base[D->rank + 1] : nullptr;
  h.chan = chan;
  h.active = 1;
  return h;
}

// halo exchange: segments sent downwards / upwards (my first / my last rows), wait for the neighbours'
int dist_halo2(cm_dist* D, const unsigned long long* offs_dn, const int* ns_dn, int nseg_dn,
               const unsigned long long* offs_up, const int* ns_up, int nseg_up, cudaStream_t st) {
  if (D == nullptr || D->world <= 1) return CM_OK;
  if (nseg_dn > 8 || nseg_up > 8) return CM_E_INVALID;
  PushSegs sd{}, su{};
  sd.nseg = nseg_dn;
  for (int i = 0; i < nseg_dn; ++i) { sd.off[i] = offs_dn[i]; sd.n[i] = ns_dn[i]; }
  su.nseg = nseg_up;
  for (int i = 0; i < nseg_up; ++i) { su.off[i] = offs_up[i]; su.n[i] = ns_up[i]; }
  exchange_kernel<<<2, 1024, 0, st>>>(dist_halo_args(D, 0), sd, su);
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
  allgather_kernel<<<(D->world - 1) * kAgSplit, 1024, 0, st>>>(mb, make_peers(D), D->rank, D->world, s);
  CM_LAUNCH_CHECK();
  return CM_OK;
}

// in-place sum over ranks of n <= 64 device doubles (deterministic: fixed rank order)
int dist_allreduce(cm_dist* D, double* vals, int n, cudaStream_t st) {
  if (D == nullptr || D->world <= 1) return CM_OK;
  if (n > kCollMax) return CM_E_INVALID;
  double* mb = (double*)D->peer_base[D->rank];
  allreduce_kernel<<<D->world, 64, 0, st>>>(mb, make_peers(D), D->rank, D->world, vals, n);
  CM_LAUNCH_CHECK();
  return CM_OK;
}

}  // namespace cm
