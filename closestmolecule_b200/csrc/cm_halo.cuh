// Device side of the strip-partitioned multi-GPU protocol: layout of the control block at the
// start of every rank's IPC-shared workspace and the push / wait helpers that the compute kernels call
// INLINE (no separate exchange launch): the CTA that produces a strip-boundary row stores it straight into
// the neighbour's copy of the same array over NVLink and publishes a sequence number there; the CTA that
// consumes a ghost row first waits for the matching number.
//
// Protocol.  Every rank runs the same sequence of kernels.  Per channel there are two one-way lanes per
// neighbour pair: "down" (a rank's row rb goes to the rank below) and "up" (row re-1 goes to the rank
// above).  cnt_dn / cnt_up count the down- / up-pushes this rank has issued (bumped also by ranks without
// that neighbour, so counts agree at equal program points); the flag a push leaves in the neighbour's
// block is the pusher's new count.  A consumer that needs the row its lower neighbour pushed up waits
// for flag_from_below >= own cnt_up (the neighbour has issued as many up-pushes as this rank at the same
// program point), and symmetrically flag_from_above >= own cnt_dn.  A wait never shares a kernel with a
// bump of the counter it reads.
//
// Write-after-read safety rests on the zebra colouring: strip boundaries are even rows on every
// distributed level, so a rank pushes down only from colour-0 half sweeps and up only from colour-1 half
// sweeps / the prolongation of odd rows, and every re-push of a ghost row is preceded (in the pusher's
// program order) by a wait on a push that the reader issued after its last read of the old value.  The
// host side rejects partitions that are not aligned to 2^Ld rows (cm_pqsolver.cu).
#pragma once
#include "cm_common.cuh"

namespace cm {

constexpr int kMaxPeers = 8;
constexpr int kCollMax = 64;     // doubles per all-reduce
constexpr int kHaloChannels = 2;

// control block, u64 / double units
//   channel c at [8c, 8c+8): [0] flag from below  [1] flag from above  [2] cnt_dn  [3] cnt_up
//   [16] all-gather counter  [17] all-reduce counter  [18] all-reduce local reader count
//   [24,32) all-gather flags per source rank   [32,40) all-reduce flags per source rank
//   [40,48) all-gather: CTAs finished per destination (cumulative, cm_dist.cu)
//   [48, 48 + 2*8*64) all-reduce slots [parity][src][kCollMax]
constexpr size_t kChFlagBelow = 0, kChFlagAbove = 1, kChCntDn = 2, kChCntUp = 3, kChStride = 8;
constexpr size_t kCntAg = 16, kCntAr = 17, kArReaders = 18;
constexpr size_t kCtlAgFlags = 24;
constexpr size_t kCtlCollFlags = 32;
constexpr size_t kCtlCollBuf = 48;
constexpr size_t kCtlDoubles = kCtlCollBuf + 2 * kMaxPeers * kCollMax + 16;

// kernel argument: the strip neighbours of this rank.  active == 0 (single GPU): nothing is touched.
struct Halo {
  double* my_base;   // start of my workspace = control block
  double* dn_base;   // lower neighbour's workspace as mapped here (nullptr: none)
  double* up_base;   // upper neighbour's
  int chan;
  int active;
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
  if (ld_acquire_sys(flag) >= seq) return;
  const unsigned long long t0 = global_ns();
  while (ld_acquire_sys(flag) < seq) {
    __nanosleep(32);
    if (global_ns() - t0 > 60000000000ull) asm volatile("trap;");
  }
}

__device__ __forceinline__ unsigned long long* halo_ctl(const Halo& h) {
  return reinterpret_cast<unsigned long long*>(h.my_base) + (size_t)h.chan * kChStride;
}
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
  return *reinterpret_cast<const volatile unsigned long long*>(p);
}

// one thread: the row the LOWER neighbour pushed up most recently (my ghost row rb-1) has arrived
__device__ __forceinline__ void halo_wait_below(const Halo& h) {
  if (!h.active || h.dn_base == nullptr) return;
  const unsigned long long* c = halo_ctl(h);
  spin_until(c + kChFlagBelow, ld_volatile_u64(c + kChCntUp));
}
// one thread: the row the UPPER neighbour pushed down most recently (my ghost row re) has arrived
__device__ __forceinline__ void halo_wait_above(const Halo& h) {
  if (!h.active || h.up_base == nullptr) return;
  const unsigned long long* c = halo_ctl(h);
  spin_until(c + kChFlagAbove, ld_volatile_u64(c + kChCntDn));
}

// my copy of an array -> the same array in a neighbour's workspace
__device__ __forceinline__ double* halo_peer_ptr(const Halo& h, double* peer_base, double* mine) {
  return peer_base + (mine - h.my_base);
}

// publish a finished down- / up-push.  Called by ALL threads of the (single) CTA that stored the row into
// the neighbour's memory; bumps the counter also when there is no neighbour on that side.
__device__ __forceinline__ void halo_publish(const Halo& h, const bool down) {
  if (!h.active) return;
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long* c = halo_ctl(h);
    const unsigned long long seq = ld_volatile_u64(c + (down ? kChCntDn : kChCntUp)) + 1ull;
    c[down ? kChCntDn : kChCntUp] = seq;
    double* pb = down ? h.dn_base : h.up_base;
    if (pb != nullptr)
      st_release_sys(reinterpret_cast<unsigned long long*>(pb) + (size_t)h.chan * kChStride +
                         (down ? kChFlagAbove : kChFlagBelow), seq);
  }
}

}  // namespace cm
