// Kernel argument blocks and launchers of cm_mg.cu.
#pragma once
#include "cm_common.cuh"
#include "cm_halo.cuh"

namespace cm {

// one zebra half sweep: for l = 0..nlines-1 solve the x-line of grid row jstart + l*jstep,
//   T x_line = b_line - (A - T) x,   T = tridiagonal part, factors lf / iu / cu (cm_stencil.cu)
// Line 0 is the strip-boundary line of a multi-GPU partition: it may have to wait for the ghost row it
// reads (wait: 1 the row below, 2 the row above) and is stored into the neighbour's copy of x as well
// (push: 1 lower neighbour, 2 upper neighbour).
struct LineJob {
  int nx, ny;
  int jstart, jstep, nlines;
  int wait, push;
  int prefetch;   // pull the next line of every stream into L2 (long-line kernel)
  int dbg;   // phase stamps for tools/time_ring.py (0: off, 1: thread 0, 2: leader of the last warp group)
  const double* A;
  const double* lf;
  const double* iu;
  const double* cu;
  const double* b;
  double* x;
  const float* F;   // fp32 pack of the operator streams (mg_pack_f32), nullptr: fp64 arrays
  int n1p;          // its row stride
};

// one half sweep of the cache-resident middle levels (cm_mgmid.cu)
struct MidJob {
  int nx, ny;
  int colour, nlines;
  int tl;       // threads per line (32 * warps)
  int groups;   // lines per CTA
  const double* A;
  const double* lf;
  const double* iu;
  const double* cu;
  const double* b;
  double* x;
  const double* ec;   // mode 3: coarse-grid correction
  double* bc;         // mode 2: coarse right-hand side
};

struct CoarseLevel {
  int nx, ny;
  const double* A;
  const double* lf;
  const double* iu;
  const double* cu;
  double* b;
  double* x;
};

constexpr int kCoarseMaxLevels = 8;
struct CoarseCycle {
  int nlev, pre, post;
  int dbg;   // phase stamps (diagnostics)
  int ainv_smem;   // shared-memory resident cycle: the dense inverse fits in shared memory too
  const double* Ainv;   // dense inverse on the last level (nullptr: 30 zebra sweeps there)
  CoarseLevel lev[kCoarseMaxLevels];
};

int mg_zebra(int nx, int ny, int colour, const double* A, const double* lf, const double* iu, const double* cu,
             const double* b, double* x, int x_zero, const Halo& H, int r0, int r1, cudaStream_t st,
             bool prof = false,   // prof: count the launch in the roofline class of bench.py (finest grid)
             const float* F = nullptr);   // fp32 pack of the operator streams (long-line kernel only)
bool mg_ring_fits(int nx);
size_t mg_pack_f32_doubles(int nx, int ny);
int mg_pack_f32(int nx, int ny, const double* A, const double* lf, const double* iu, const double* cu, float* F,
                int r0, int r1, cudaStream_t st);
int mg_resid_restrict(int nxf, int nyf, const double* A, const double* b, const double* x, double* rc,
                      const Halo& H, int Y0, int Y1, cudaStream_t st);
int mg_prolong_odd(int nxc, int nyc, const double* ec, double* xf, const Halo& H, int wait_coarse, int r0,
                   int r1, cudaStream_t st);
bool mg_mid_fits(int nx, int ny);
bool mg_mid_fused_restrict_fits(int nx, int ny);
int mg_mid_init_attributes();
int mg_zebra_mid(int mode, int nx, int ny, int colour, const double* A, const double* lf, const double* iu,
                 const double* cu, const double* b, double* x, const double* ec, double* bc, cudaStream_t st);
int mg_coarse_cycle(const CoarseCycle& C, int ncta, cudaStream_t st);
size_t mg_coarse_smem_bytes(const CoarseCycle& C, bool dense_last);
int mg_coarse_smem_init_attributes();
int mg_smem_debug_times(unsigned long long* host_out, int n);   // returns the number of stamps (< 0: error)
int mg_coarse_smem_cycle(const CoarseCycle& C, cudaStream_t st);
int mg_coarse_max_n1();
int mg_init_attributes();
int mg_ring_debug_times(unsigned long long* host_out, int n);

Halo dist_halo_args(const cm_dist* D, int chan);

}  // namespace cm
