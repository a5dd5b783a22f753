// Newton linear solve J dx = F on structured meshes: GMRES on the row-equilibrated system,
// right-preconditioned by the field split diag(App, Aqq) (optionally block-lower-triangular
// [[App, 0], [Aqp, Aqq]], CM_PRECOND_LOWER=1):
//   p-block: one geometric-multigrid V(1,2) cycle with zebra x-line relaxation (Galerkin coarse
//            operators in stencil form, dense inverse on the coarsest grid);
//   q-block: two zebra x-line relaxation sweeps (the q-equation is transport along r2).
// All of it is caller-provided workspace + kernels; nothing is kept between calls except the
// workspace content written by cm_pqs_setup.  Replaces the sparse LU the reference requests with
// linear_solver='mumps'|'umfpack' (advection_diffusion_convergence.py:212,
// …MixedPrimal_convergence_SUPG.py:137).  The generic Jacobi-GMRES for CSR matrices is here too.
#include <stdlib.h>

#include <vector>

#include "cm_common.cuh"
#include "cm_decl.cuh"
#include "cm_halo.cuh"
#include "cm_mg.cuh"

namespace cm {

constexpr int kMaxLevels = 24;
constexpr int kCoarseDenseMax = 256;

struct PQSLevel {
  int nx, ny;
  size_t n;
  double *A, *lf, *iu, *cu, *x, *b, *r;
  float* F;   // fp32 pack of the smoother's operator streams (levels of the long-line kernel only)
};

struct PQSolver {
  int nx, ny, nv, nlev, restart;
  int q_sweeps = 2;  // zebra sweeps on the q-block per application
  int pre_smooth = 1, post_smooth = 2;  // zebra sweeps before / after the coarse-grid correction
  int dist_ld = 2;      // multi-GPU: first replicated level
  // block-diagonal instead of block-lower-triangular field split: the q-sweeps take r_q directly
  // (no Aqp z_p correction).  Same GMRES iteration counts on the SUPG problems (the triangular
  // coupling costs GMRES nothing here), one residual pass and one halo exchange less per iteration.
  bool diag_split = true;
  bool upper_split = false;  // block-upper-triangular [[App, Apq], [0, Aqq]]: q first, then p with r_p - Apq z_q
  bool dense_coarse;
  PQSLevel lev[kMaxLevels];
  PQSLevel qlev[kMaxLevels];  // hierarchy of the 7-point stand-in of the C0IP q-block (same grids as lev)
  double* qAinv;              // its coarsest-grid inverse
  bool q_mg = false;          // C0IP: one V-cycle on the q-block instead of plain line sweeps (CM_Q_MG=1)
  int q_sweeps_wide = 6;      // C0IP: line sweeps on the 7-point stand-in of the q-block
  double *PP, *PQ, *QP, *QQ, *sp, *sq;
  double* QQx;      // 6 extra q-q diagonals of the C0IP interior-facet pattern
  double* QQl;      // 7-point stand-in of the 13-point q-block for the line relaxation (st_lump_extras)
  int* has_extra;   // device flag: any of them non-zero
  double *qlf, *qiu, *qcu;
  float* qF;        // fp32 pack of the q-block smoother streams (finest grid)
  double *Ainv, *Mwork;
  double *bs, *xs, *tq, *ru, *ustate;
  double* gm;
  int* err;
  size_t total_doubles;
};

static inline size_t even(size_t n) { return (n + 1) & ~(size_t)1; }

// Tuning knobs, read from the environment by the caller that creates a solver (once per context, never
// cached in function statics: the C ABI stays re-entrant).  Defaults measured at 16 Mi cells
// (profiles/README.md): V(1,1) 13 its / 33.9 ms, V(1,2) 11 / 32.9, V(2,2) 10 / 33.9; q-sweeps 1: +1 it.
struct PQKnobs {
  int precond_lower = 0, precond_upper = 0;
  int q_sweeps = 2, pre = 1, post = 2;
  int dist_ld = 2;          // multi-GPU: levels >= ld are replicated
  int q_mg = 0, q_sweeps_wide = 6;
  int graph = 1;            // replay fixed kernel sequences as CUDA graphs
  int v1 = 0;               // CM_SOLVER_V1=1: first-generation kernels (two-kernel zebra sweeps, host-driven GMRES)
  int coarse_ncta = 8;      // CTAs of the cluster that runs the fused coarse cycle (1 or 8)
  int coarse_n1 = -1;       // the global-memory fused coarse cycle starts at the first level with nx + 1 <= coarse_n1 (0: off)
  int coarse_smem = 1;      // shared-memory resident cycle (one CTA) from the first level whose arrays fit
  int q_concurrent = 1;     // q-block sweeps on a second stream, concurrently with the p-block V-cycle
  int post_coarse = -1;     // post-smoothing sweeps on the levels >= coarse_from (-1: same as post)
  int post_smem = 1;        // post-smoothing sweeps inside the shared-memory resident cycle (65 x 33 and coarser at 16 Mi
                            // cells): 1 instead of 2 keeps the 11 GMRES iterations and shortens the cycle by ~10 us; on the
                            // levels above it costs an iteration (measured, profiles/README.md)
  int coarse_from = 2;
  int gmres_host = 0;       // diagnostics: second-generation preconditioner under the host-driven GMRES (1 GPU)
  int precond_v1 = 0;       // diagnostics: first-generation preconditioner under the device-resident GMRES (1 GPU)
  int mid_max_n1 = 1280;    // cache-resident middle levels (cm_mgmid.cu): lines of up to this many unknowns (0: off)
  int mid_fuse = 1;         // ... with the restricted residual / the prolongation folded into the half sweeps
  int fuse_unscale = 1;     // the GMRES update kernel writes the next preconditioner input (no separate unscaling pass)
  int f32_op = 1;           // long-line smoother: operator streams from the fp32 pack (x, b and all arithmetic stay fp64)
  int graph_setup = 1;      // replay the solver setup behind a block-writing assembly as one CUDA graph
  int graph_while = 1;      // a whole GMRES cycle as ONE graph launch (WHILE node around the iteration, condition set on the device)
  int q_low_priority = 1;   // the q-block stream gets the lowest priority: its sweeps fill the SMs / the HBM time the p-block
                            // V-cycle leaves idle on its latency-bound coarse levels instead of competing with its fine sweeps
};
static int env_int(const char* name, int dflt, int minval) {
  const char* e = getenv(name);
  if (!e) return dflt;
  const int v = atoi(e);
  return v >= minval ? v : dflt;
}
static PQKnobs read_knobs() {
  PQKnobs K;
  K.precond_lower = env_int("CM_PRECOND_LOWER", 0, 0);
  K.precond_upper = env_int("CM_PRECOND_UPPER", 0, 0);
  K.q_sweeps = env_int("CM_Q_SWEEPS", 2, 1);
  K.pre = env_int("CM_MG_PRE", 1, 1);
  K.post = env_int("CM_MG_POST", 2, 0);
  K.dist_ld = env_int("CM_DIST_LD", 2, 0);
  K.q_mg = env_int("CM_Q_MG", 0, 0);
  K.q_sweeps_wide = env_int("CM_Q_SWEEPS_WIDE", 6, 1);
  K.graph = env_int("CM_GRAPH", 1, 0);
  K.v1 = env_int("CM_SOLVER_V1", 0, 0);
  K.coarse_ncta = env_int("CM_COARSE_NCTA", 8, 1) >= 8 ? 8 : 1;
  // measured at 16 Mi cells (profiles/README.md): the cluster kernel with global-memory phases (3-4 us each) is no
  // faster than separate kernels inside a CUDA graph -> off by default; the shared-memory resident cycle is on
  K.coarse_n1 = env_int("CM_COARSE_N1", 0, 0);
  K.coarse_smem = env_int("CM_COARSE_SMEM", 1, 0);
  K.q_concurrent = env_int("CM_Q_CONCURRENT", 1, 0);
  K.post_coarse = env_int("CM_MG_POST_COARSE", -1, 1);
  K.post_smem = env_int("CM_MG_POST_SMEM", 1, 1);
  K.coarse_from = env_int("CM_MG_COARSE_FROM", 2, 0);
  K.gmres_host = env_int("CM_GMRES_HOST", 0, 0);
  K.precond_v1 = env_int("CM_PRECOND_V1", 0, 0);
  K.mid_max_n1 = env_int("CM_MID_MAX_N1", 1280, 0);
  K.mid_fuse = env_int("CM_MID_FUSE", 1, 0);
  K.fuse_unscale = env_int("CM_FUSE_UNSCALE", 1, 0);
  K.f32_op = env_int("CM_F32_OP", 1, 0);
  K.graph_setup = env_int("CM_GRAPH_SETUP", 1, 0);
  K.graph_while = env_int("CM_GRAPH_WHILE", 1, 0);
  K.q_low_priority = env_int("CM_Q_LOW_PRIORITY", 1, 0);
  return K;
}

static int pqs_layout(PQSolver& S, int nx, int ny, int restart, double* base, const PQKnobs& K) {
  S.diag_split = K.precond_lower == 0;
  S.upper_split = K.precond_upper != 0;
  S.q_sweeps = K.q_sweeps;
  S.pre_smooth = K.pre;
  S.post_smooth = K.post;
  S.dist_ld = K.dist_ld;
  S.nx = nx;
  S.ny = ny;
  S.nv = (nx + 1) * (ny + 1);
  S.restart = restart;
  size_t off = 0;
  auto take = [&](size_t n) {
    double* p = base ? base + off : nullptr;
    off += even(n);
    return p;
  };
  const size_t V = (size_t)S.nv;
  take(dist_ctl_doubles());  // flags / all-reduce slots of the multi-GPU path (unused on one GPU)
  S.ustate = take(2 * V);     // Newton state u (interleaved) lives here so its halo rows can be pushed
  S.PP = take(7 * V);
  S.PQ = take(7 * V);
  S.QP = take(7 * V);
  S.QQ = take(7 * V);
  S.QQx = take(6 * V);
  S.QQl = take(7 * V);
  S.has_extra = (int*)take(2);
  S.sp = take(V);
  S.sq = take(V);
  S.qlf = take(V);
  S.qiu = take(V);
  S.qcu = take(V);
  S.bs = take(2 * V);
  S.xs = take(2 * V);
  S.tq = take(V);
  S.ru = take(2 * V);
  int lx = nx, ly = ny, l = 0;
  for (;;) {
    if (l >= kMaxLevels) return CM_E_UNSUPPORTED;
    PQSLevel& L = S.lev[l];
    L.nx = lx;
    L.ny = ly;
    L.n = (size_t)(lx + 1) * (ly + 1);
    L.A = (l == 0) ? S.PP : take(7 * L.n);
    L.lf = take(L.n);
    L.iu = take(L.n);
    L.cu = take(L.n);
    L.r = take(L.n);
    L.x = (l == 0) ? nullptr : take(L.n);
    L.b = (l == 0) ? nullptr : take(L.n);
    ++l;
    if ((lx & 1) || (ly & 1) || (lx < ly ? lx : ly) <= 4) break;
    lx >>= 1;
    ly >>= 1;
  }
  S.nlev = l;
  const size_t nc = S.lev[l - 1].n;
  S.dense_coarse = nc <= (size_t)kCoarseDenseMax;
  S.Ainv = take(S.dense_coarse ? nc * nc : 2);
  S.Mwork = take(S.dense_coarse ? 2 * nc * nc : 2);
  for (int k = 0; k < S.nlev; ++k) {
    PQSLevel& Q = S.qlev[k];
    const PQSLevel& L = S.lev[k];
    Q.nx = L.nx;
    Q.ny = L.ny;
    Q.n = L.n;
    Q.A = (k == 0) ? S.QQl : take(7 * Q.n);
    Q.lf = (k == 0) ? S.qlf : take(Q.n);
    Q.iu = (k == 0) ? S.qiu : take(Q.n);
    Q.cu = (k == 0) ? S.qcu : take(Q.n);
    Q.r = (k == 0) ? L.r : take(Q.n);   // level-0 scratch is shared: the two cycles run one after the other
    Q.x = (k == 0) ? nullptr : take(Q.n);
    Q.b = (k == 0) ? nullptr : take(Q.n);
  }
  S.qAinv = take(S.dense_coarse ? nc * nc : 2);
  for (int k = 0; k < S.nlev; ++k)
    S.lev[k].F = mg_ring_fits(S.lev[k].nx) ? (float*)take(mg_pack_f32_doubles(S.lev[k].nx, S.lev[k].ny)) : nullptr;
  S.qF = mg_ring_fits(nx) ? (float*)take(mg_pack_f32_doubles(nx, ny)) : nullptr;
  // 4 Mi cells (profiles/README.md): the V-cycle on the q-block helps the minus-sign C0IP variant (114 GMRES
  // iterations) but diverges on the "paper" variant (Galerkin coarse operators of the transport part): off
  // by default; line sweeps on the 7-point stand-in instead: 3 sweeps 94 / 222 iterations ("paper" /
  // minus-sign variant), 6 sweeps 89 / 123
  S.q_mg = K.q_mg != 0;
  S.q_sweeps_wide = K.q_sweeps_wide;
  S.err = (int*)take(4);
  {
    const size_t g1 = gmres_workspace_doubles(2 * (long long)V, restart);
    const size_t g2 = gmres_dev_doubles(2 * (long long)V, restart);
    S.gm = take(g1 > g2 ? g1 : g2);
  }
  S.total_doubles = off;
  return CM_OK;
}

static int check_err_flag(PQSolver& S, cudaStream_t st, const char* what) {
  int h = 0;
  CM_CUDA(cudaMemcpyAsync(&h, S.err, sizeof(int), cudaMemcpyDeviceToHost, st));
  CM_CUDA(cudaStreamSynchronize(st));
  if (h == 1) {
    set_error("%s: the Jacobian is not a 7-point stencil on the (nx,ny) grid (interior-facet "
              "pattern or unstructured mesh): the structured solver does not apply", what);
    return CM_E_UNSUPPORTED;
  }
  if (h == 2) {
    set_error("%s: zero pivot in an x-line factorisation (unstabilised Galerkin q-block?)", what);
    return CM_E_BREAKDOWN;
  }
  if (h == 3) {
    set_error("%s: singular coarsest-grid operator", what);
    return CM_E_BREAKDOWN;
  }
  return CM_OK;
}

// L.r is free during smoothing: it serves as the scratch of the pipelined sweep
static int zebra_sweep(const PQSLevel& L, const double* A, const double* lf, const double* iu,
                       const double* cu, const double* b, double* x, bool x_zero, cudaStream_t st, bool prof = false) {
  int rc = st_zebra(L.nx, L.ny, 0, A, lf, iu, cu, b, x, x_zero ? 1 : 0, st, 0, -1, L.r, nullptr, nullptr, prof);
  if (rc) return rc;
  return st_zebra(L.nx, L.ny, 1, A, lf, iu, cu, b, x, 0, st, 0, -1, L.r, nullptr, nullptr, prof);
}

static int vcycle_on(PQSolver& S, PQSLevel* lev, const double* Ainv, int l, const double* b, double* x,
                     cudaStream_t st) {
  PQSLevel& L = lev[l];
  int rc;
  if (l == S.nlev - 1) {
    if (S.dense_coarse) return st_coarse_apply((int)L.n, Ainv, b, x, st);
    rc = zebra_sweep(L, L.A, L.lf, L.iu, L.cu, b, x, true, st);
    for (int s = 1; s < 30 && !rc; ++s) rc = zebra_sweep(L, L.A, L.lf, L.iu, L.cu, b, x, false, st);
    return rc;
  }
  PQSLevel& C = lev[l + 1];
  for (int s = 0; s < S.pre_smooth; ++s)
    if ((rc = zebra_sweep(L, L.A, L.lf, L.iu, L.cu, b, x, s == 0, st, l == 0))) return rc;
  if ((rc = st_residual(L.nx, L.ny, L.A, b, x, L.r, st))) return rc;
  if ((rc = st_restrict(C.nx, C.ny, L.r, C.b, st))) return rc;
  if ((rc = vcycle_on(S, lev, Ainv, l + 1, C.b, C.x, st))) return rc;
  if ((rc = st_prolong_add(C.nx, C.ny, C.x, x, st))) return rc;
  for (int s = 0; s < S.post_smooth && !rc; ++s) rc = zebra_sweep(L, L.A, L.lf, L.iu, L.cu, b, x, false, st, l == 0);
  return rc;
}

static int vcycle(PQSolver& S, int l, const double* b, double* x, cudaStream_t st) {
  return vcycle_on(S, S.lev, S.Ainv, l, b, x, st);
}

// Replays a fixed kernel sequence (one preconditioner application) as a CUDA graph: captured on a
// private stream at the second call (the first call runs directly so that one-time function
// attributes are set outside the capture), launched on the caller's stream afterwards.  At small
// per-GPU sizes the sequence is launch-latency bound; the graph removes the CPU launch cost.
struct GraphReplayer {
  cudaGraphExec_t exec = nullptr;
  cudaStream_t cap = nullptr;
  const void* key = nullptr;
  int calls = 0;
  long long launches = 0;
  double bytes = 0.0;
  bool enabled = true;
  GraphReplayer() { enabled = env_int("CM_GRAPH", 1, 0) != 0; }
  ~GraphReplayer() {
    if (exec) cudaGraphExecDestroy(exec);
    if (cap) cudaStreamDestroy(cap);
  }
  template <class F>
  int run(const void* k, cudaStream_t st, F&& body) {
    if (!enabled || g_prof_is_on()) return body(st);
    if (exec && key == k) {
      g_launches += launches;
      g_alg_bytes += bytes;
      return cudaGraphLaunch(exec, st) == cudaSuccess ? CM_OK : (int)CM_E_CUDA;
    }
    if (calls++ < 1 || exec) return body(st);
    if (!cap && cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking) != cudaSuccess) return body(st);
    if (cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
      cudaGetLastError();
      enabled = false;
      return body(st);
    }
    const long long l0 = g_launches.load();
    const double b0 = g_alg_bytes;
    const int rc = body(cap);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(cap, &graph);
    launches = g_launches.load() - l0;
    bytes = g_alg_bytes - b0;
    if (rc != CM_OK || ce != cudaSuccess || graph == nullptr) {
      cudaGetLastError();
      if (graph) cudaGraphDestroy(graph);
      enabled = false;
      return rc != CM_OK ? rc : body(st);
    }
    if (cudaGraphInstantiate(&exec, graph, 0) != cudaSuccess) {
      cudaGetLastError();
      cudaGraphDestroy(graph);
      exec = nullptr;
      enabled = false;
      return body(st);
    }
    cudaGraphDestroy(graph);
    key = k;
    return cudaGraphLaunch(exec, st) == cudaSuccess ? CM_OK : (int)CM_E_CUDA;
  }
};

size_t pqs_workspace_bytes(int nx, int ny, int restart) {
  PQSolver S;
  if (pqs_layout(S, nx, ny, restart, nullptr, read_knobs())) return 0;
  return S.total_doubles * sizeof(double) + 64;
}

static double* align_ws(void* ws) { return (double*)(((uintptr_t)ws + 63) & ~(uintptr_t)63); }

int pqs_setup(int nx, int ny, int restart, const int* nrowptr, const int* ncol, const double* vals,
              void* ws, size_t ws_bytes, cudaStream_t st, bool blocks_ready, bool no_check) {
  PQSolver S;
  int rc = pqs_layout(S, nx, ny, restart, align_ws(ws), read_knobs());
  if (rc) return rc;
  if (S.total_doubles * sizeof(double) + 64 > ws_bytes) {
    set_error("pqs_setup: workspace too small");
    return CM_E_INVALID;
  }
  CM_CUDA(cudaMemsetAsync(S.err, 0, sizeof(int), st));
  CM_CUDA(cudaMemsetAsync(S.has_extra, 0, sizeof(int), st));
  int wide = 0;
  // blocks_ready: the assembly sweep wrote PP / PQ / QP / QQ / sp / sq itself (7-point pattern by construction)
  if (!blocks_ready && (rc = st_extract_blocks(nx, ny, nrowptr, ncol, vals, S.PP, S.PQ, S.QP, S.QQ, S.sp, S.sq, S.err, st, 0,
                                               -1, S.QQx, S.has_extra, &wide)))
    return rc;
  if (wide && (rc = st_lump_extras(nx, ny, S.QQ, S.QQx, S.QQl, st))) return rc;
  for (int l = 0; l < S.nlev; ++l) {
    PQSLevel& L = S.lev[l];
    if ((rc = st_line_factor(L.nx, L.ny, L.A, L.lf, L.iu, L.cu, S.err, st))) return rc;
    if (l + 1 < S.nlev && (rc = st_rap(S.lev[l + 1].nx, S.lev[l + 1].ny, L.A, S.lev[l + 1].A, st)))
      return rc;
  }
  PQSLevel& C = S.lev[S.nlev - 1];
  if (S.dense_coarse && (rc = st_coarse_invert(C.nx, C.ny, C.A, S.Mwork, S.Ainv, S.err, st))) return rc;
  if ((rc = st_line_factor(nx, ny, wide ? S.QQl : S.QQ, S.qlf, S.qiu, S.qcu, S.err, st))) return rc;
  if (wide && S.q_mg) {  // Galerkin hierarchy of the 7-point stand-in of the C0IP q-block
    for (int l = 0; l < S.nlev; ++l) {
      PQSLevel& Q = S.qlev[l];
      if (l > 0 && (rc = st_line_factor(Q.nx, Q.ny, Q.A, Q.lf, Q.iu, Q.cu, S.err, st))) return rc;
      if (l + 1 < S.nlev && (rc = st_rap(S.qlev[l + 1].nx, S.qlev[l + 1].ny, Q.A, S.qlev[l + 1].A, st)))
        return rc;
    }
    PQSLevel& QC = S.qlev[S.nlev - 1];
    if (S.dense_coarse && (rc = st_coarse_invert(QC.nx, QC.ny, QC.A, S.Mwork, S.qAinv, S.err, st))) return rc;
  }
  return no_check ? (int)CM_OK : check_err_flag(S, st, "cm_pqs_setup");
}

int pqs_solve(int nx, int ny, int restart, const double* b, double* x, double rtol, double atol,
              int maxit, void* ws, size_t ws_bytes, int* iters, double* resid, cudaStream_t st) {
  PQSolver S;
  int rc = pqs_layout(S, nx, ny, restart, align_ws(ws), read_knobs());
  if (rc) return rc;
  if (S.total_doubles * sizeof(double) + 64 > ws_bytes) {
    set_error("pqs_solve: workspace too small");
    return CM_E_INVALID;
  }
  const int V = S.nv;
  int wide = 0;  // C0IP interior-facet couplings present?
  CM_CUDA(cudaMemcpyAsync(&wide, S.has_extra, sizeof(int), cudaMemcpyDeviceToHost, st));
  CM_CUDA(cudaStreamSynchronize(st));
  if ((rc = st_split_scale(V, b, S.sp, S.sq, S.bs, st))) return rc;
  LinOp A = [&](const double* in, double* out) {
    return st_pq_spmv(nx, ny, S.PP, S.PQ, S.QP, S.QQ, S.sp, S.sq, in, out, st, 0, -1, S.QQx, S.has_extra);
  };
  // M^{-1} = [[App,0],[Aqp,Aqq]]^{-1} (approximately) on the UNSCALED operator, composed with the
  // inverse row scaling: multigrid needs the true (variational) FE operator to be mesh independent
  GraphReplayer replay;
  LinOp Minv = [&](const double* r, double* z) {
    int e = st_unscale(nx, ny, r, S.sp, S.sq, S.ru, st);
    if (e) return e;
    return replay.run(z, st, [&](cudaStream_t s_) -> int {
      // 13-point (C0IP) q-block: the relaxation runs on its 7-point stand-in (st_lump_extras); the outer
      // Krylov operator keeps the full stencil
      const double* Aq = wide ? S.QQl : S.QQ;
      auto q_sweeps = [&](const double* bq) -> int {
        if (wide && S.q_mg) return vcycle_on(S, S.qlev, S.qAinv, 0, bq, z + V, s_);
        int e3 = CM_OK;
        const int nsw = wide ? S.q_sweeps_wide : S.q_sweeps;
        for (int sw = 0; sw < nsw && !e3; ++sw) {
          e3 = st_zebra(nx, ny, 0, Aq, S.qlf, S.qiu, S.qcu, bq, z + V, sw == 0 ? 1 : 0, s_, 0, -1, S.lev[0].r, nullptr,
                        nullptr, true);
          if (!e3)
            e3 = st_zebra(nx, ny, 1, Aq, S.qlf, S.qiu, S.qcu, bq, z + V, 0, s_, 0, -1, S.lev[0].r, nullptr, nullptr, true);
        }
        return e3;
      };
      int e2;
      if (S.upper_split) {  // z_q = Mq^{-1} r_q, then z_p = Mp^{-1} (r_p - Apq z_q)
        if ((e2 = q_sweeps(S.ru + V))) return e2;
        if ((e2 = st_residual(nx, ny, S.PQ, S.ru, z + V, S.tq, s_))) return e2;
        return vcycle(S, 0, S.tq, z, s_);
      }
      if ((e2 = vcycle(S, 0, S.ru, z, s_))) return e2;
      const double* bq = S.ru + V;
      if (!S.diag_split) {
        if ((e2 = st_residual(nx, ny, S.QP, S.ru + V, z, S.tq, s_))) return e2;
        bq = S.tq;
      }
      e2 = q_sweeps(bq);
      return e2;
    });
  };
  double bnorm = 0.0;
  rc = gmres_solve(2LL * V, A, Minv, S.bs, S.xs, rtol, atol, maxit, restart, 0, S.gm, iters, resid,
                   &bnorm, st);
  if (rc < 0) return rc;
  if (resid && bnorm > 0.0) *resid /= bnorm;
  int rc2 = st_interleave(V, S.xs, x, st);
  if (rc2) return rc2;
  if (rc == 1) {
    set_error("cm_pqs_solve: GMRES did not reach rtol=%g in %d iterations (relative residual %g)",
              rtol, maxit, resid ? *resid : -1.0);
  }
  return rc;
}

// ------------------------------------------------------------------------------------------------
// strip-partitioned multi-GPU variant: levels 0..Ld-1 are distributed by vertex rows (every rank
// works on its own rows of globally indexed arrays and pushes its boundary rows to the neighbours),
// levels >= Ld are replicated after an all-gather of the level-Ld right-hand side.
// ------------------------------------------------------------------------------------------------
struct DistInfo {
  cm_dist* D;
  double* base;
  int Ld;                       // first replicated level
  int rb[kMaxLevels], re[kMaxLevels];  // owned rows per distributed level
};

static void dist_info(const PQSolver& S, cm_dist* D, double* base, DistInfo& I) {
  I.D = D;
  I.base = base;
  I.Ld = S.nlev - 1 < S.dist_ld ? S.nlev - 1 : S.dist_ld;
  for (int l = 0; l <= I.Ld; ++l) {
    I.rb[l] = D->row_begin[D->rank] >> l;
    I.re[l] = (D->rank + 1 == D->world) ? S.lev[l].ny + 1 : (D->row_begin[D->rank + 1] >> l);
  }
}

// push row rb (downwards) and row re-1 (upwards) of every listed array, wait for the neighbours'
static int halo_rows(const DistInfo& I, double* const* arrs, int narr, int n1, int rb, int re,
                     cudaStream_t st) {
  unsigned long long od[8], ou[8];
  int nd[8], nu[8];
  if (narr > 8) return CM_E_INVALID;
  for (int a = 0; a < narr; ++a) {
    od[a] = (unsigned long long)(arrs[a] - I.base) + (unsigned long long)rb * n1;
    ou[a] = (unsigned long long)(arrs[a] - I.base) + (unsigned long long)(re - 1) * n1;
    nd[a] = nu[a] = n1;
  }
  return dist_halo2(I.D, od, nd, narr, ou, nu, narr, st);
}

static int vcycle_dist(PQSolver& S, const DistInfo& I, int l, double* b, double* x, cudaStream_t st) {
  int rc;
  if (l == I.Ld) {
    // gather the owned rows of b from every rank, then the ordinary (replicated) cycle
    PQSLevel& L = S.lev[l];
    unsigned long long off = (unsigned long long)(b - I.base) + (unsigned long long)I.rb[l] * (L.nx + 1);
    int n = (I.re[l] - I.rb[l]) * (L.nx + 1);
    if ((rc = dist_allgather(I.D, &off, &n, 1, st))) return rc;
    return vcycle(S, l, b, x, st);
  }
  PQSLevel& L = S.lev[l];
  PQSLevel& C = S.lev[l + 1];
  const int n1 = L.nx + 1, rb = I.rb[l], re = I.re[l];
  double* xa[1] = {x};
  for (int s = 0; s < S.pre_smooth; ++s) {
    if ((rc = st_zebra(L.nx, L.ny, 0, L.A, L.lf, L.iu, L.cu, b, x, s == 0 ? 1 : 0, st, rb, re, L.r))) return rc;
    if ((rc = halo_rows(I, xa, 1, n1, rb, re, st))) return rc;
    if ((rc = st_zebra(L.nx, L.ny, 1, L.A, L.lf, L.iu, L.cu, b, x, 0, st, rb, re, L.r))) return rc;
    if ((rc = halo_rows(I, xa, 1, n1, rb, re, st))) return rc;
  }
  if ((rc = st_residual(L.nx, L.ny, L.A, b, x, L.r, st, rb, re))) return rc;
  double* ra[1] = {L.r};
  if ((rc = halo_rows(I, ra, 1, n1, rb, re, st))) return rc;
  if ((rc = st_restrict(C.nx, C.ny, L.r, C.b, st, I.rb[l + 1], I.re[l + 1]))) return rc;
  if ((rc = vcycle_dist(S, I, l + 1, C.b, C.x, st))) return rc;
  if (l + 1 < I.Ld) {  // replicated levels hold the full vector already
    double* ca[1] = {C.x};
    if ((rc = halo_rows(I, ca, 1, C.nx + 1, I.rb[l + 1], I.re[l + 1], st))) return rc;
  }
  if ((rc = st_prolong_add(C.nx, C.ny, C.x, x, st, rb, re))) return rc;
  for (int s = 0; s < S.post_smooth; ++s) {
    if ((rc = halo_rows(I, xa, 1, n1, rb, re, st))) return rc;
    if ((rc = st_zebra(L.nx, L.ny, 0, L.A, L.lf, L.iu, L.cu, b, x, 0, st, rb, re, L.r))) return rc;
    if ((rc = halo_rows(I, xa, 1, n1, rb, re, st))) return rc;
    if ((rc = st_zebra(L.nx, L.ny, 1, L.A, L.lf, L.iu, L.cu, b, x, 0, st, rb, re, L.r))) return rc;
  }
  return CM_OK;
}

int pqs_setup_dist(int nx, int ny, int restart, const int* nrowptr, const int* ncol,
                   const double* vals, void* ws, size_t ws_bytes, cm_dist* D, cudaStream_t st, int Ld_forced,
                   bool blocks_ready, bool no_check) {
  PQSolver S;
  double* base = align_ws(ws);
  int rc = pqs_layout(S, nx, ny, restart, base, read_knobs());
  if (rc) return rc;
  if (S.total_doubles * sizeof(double) + 64 > ws_bytes) {
    set_error("pqs_setup_dist: workspace too small");
    return CM_E_INVALID;
  }
  if (base != (double*)D->peer_base[D->rank]) {
    set_error("pqs_setup_dist: workspace must be the (64-byte aligned) IPC allocation registered in cm_dist");
    return CM_E_INVALID;
  }
  DistInfo I;
  if (Ld_forced >= 0) S.dist_ld = Ld_forced;
  dist_info(S, D, base, I);
  for (int r = 1; r < D->world; ++r)
    if (D->row_begin[r] & ((1 << I.Ld) - 1)) {
      // strips on level l use row_begin >> l for l <= Ld: anything else truncates the coarse-row ownership
      set_error("pqs_setup_dist: row_begin[%d] = %d is not a multiple of %d (first replicated level %d)", r,
                D->row_begin[r], 1 << I.Ld, I.Ld);
      return CM_E_INVALID;
    }
  CM_CUDA(cudaMemsetAsync(S.err, 0, sizeof(int), st));
  CM_CUDA(cudaMemsetAsync(S.has_extra, 0, sizeof(int), st));
  if (!blocks_ready && (rc = st_extract_blocks(nx, ny, nrowptr, ncol, vals, S.PP, S.PQ, S.QP, S.QQ, S.sp, S.sq, S.err,
                                               st, I.rb[0], I.re[0])))  // 7-point rows only on the multi-GPU path
    return rc;
  for (int l = 0; l < I.Ld; ++l) {
    PQSLevel& L = S.lev[l];
    const int n1 = L.nx + 1;
    const size_t n = L.n;
    double* diag[7];
    for (int k = 0; k < 7; ++k) diag[k] = L.A + (size_t)k * n;
    if ((rc = halo_rows(I, diag, 7, n1, I.rb[l], I.re[l], st))) return rc;  // RAP reads ghost rows of A
    if ((rc = st_line_factor(L.nx, L.ny, L.A, L.lf, L.iu, L.cu, S.err, st, I.rb[l], I.re[l]))) return rc;
    if ((rc = st_rap(S.lev[l + 1].nx, S.lev[l + 1].ny, L.A, S.lev[l + 1].A, st, I.rb[l + 1], I.re[l + 1])))
      return rc;
  }
  {  // replicate the level-Ld operator
    PQSLevel& L = S.lev[I.Ld];
    const int n1 = L.nx + 1;
    unsigned long long offs[7];
    int ns[7];
    for (int k = 0; k < 7; ++k) {
      offs[k] = (unsigned long long)(L.A + (size_t)k * L.n - base) + (unsigned long long)I.rb[I.Ld] * n1;
      ns[k] = (I.re[I.Ld] - I.rb[I.Ld]) * n1;
    }
    if ((rc = dist_allgather(D, offs, ns, 7, st))) return rc;
  }
  for (int l = I.Ld; l < S.nlev; ++l) {
    PQSLevel& L = S.lev[l];
    if ((rc = st_line_factor(L.nx, L.ny, L.A, L.lf, L.iu, L.cu, S.err, st))) return rc;
    if (l + 1 < S.nlev && (rc = st_rap(S.lev[l + 1].nx, S.lev[l + 1].ny, L.A, S.lev[l + 1].A, st)))
      return rc;
  }
  PQSLevel& C = S.lev[S.nlev - 1];
  if (S.dense_coarse && (rc = st_coarse_invert(C.nx, C.ny, C.A, S.Mwork, S.Ainv, S.err, st))) return rc;
  if ((rc = st_line_factor(nx, ny, S.QQ, S.qlf, S.qiu, S.qcu, S.err, st, I.rb[0], I.re[0]))) return rc;
  return no_check ? (int)CM_OK : check_err_flag(S, st, "cm_pqs_setup_dist");
}

int pqs_solve_dist(int nx, int ny, int restart, const double* b, double* x, double rtol, double atol,
                   int maxit, void* ws, size_t ws_bytes, cm_dist* D, int* iters, double* resid,
                   cudaStream_t st) {
  PQSolver S;
  double* base = align_ws(ws);
  int rc = pqs_layout(S, nx, ny, restart, base, read_knobs());
  if (rc) return rc;
  if (S.total_doubles * sizeof(double) + 64 > ws_bytes) {
    set_error("pqs_solve_dist: workspace too small");
    return CM_E_INVALID;
  }
  DistInfo I;
  dist_info(S, D, base, I);
  const int V = S.nv, n1 = nx + 1, rb = I.rb[0], re = I.re[0];
  const int vb = rb * n1, ve = re * n1;
  if ((rc = st_split_scale(V, b, S.sp, S.sq, S.bs, st, vb, ve))) return rc;
  LinOp A = [&](const double* in, double* out) {
    double* a2[2] = {const_cast<double*>(in), const_cast<double*>(in) + V};
    int e = halo_rows(I, a2, 2, n1, rb, re, st);
    if (e) return e;
    return st_pq_spmv(nx, ny, S.PP, S.PQ, S.QP, S.QQ, S.sp, S.sq, in, out, st, rb, re);
  };
  GraphReplayer replay;
  LinOp Minv = [&](const double* r, double* z) {
    int e = st_unscale(nx, ny, r, S.sp, S.sq, S.ru, st, rb, re);
    if (e) return e;
    return replay.run(z, st, [&](cudaStream_t s_) -> int {
      int e2 = vcycle_dist(S, I, 0, S.ru, z, s_);
      if (e2) return e2;
      const double* bq = S.ru + V;
      if (!S.diag_split) {
        double* zp[1] = {z};
        if ((e2 = halo_rows(I, zp, 1, n1, rb, re, s_))) return e2;
        if ((e2 = st_residual(nx, ny, S.QP, S.ru + V, z, S.tq, s_, rb, re))) return e2;
        bq = S.tq;
      }
      double* zq[1] = {z + V};
      for (int sw = 0; sw < S.q_sweeps; ++sw) {
        if ((e2 = st_zebra(nx, ny, 0, S.QQ, S.qlf, S.qiu, S.qcu, bq, z + V, sw == 0 ? 1 : 0, s_, rb, re,
                           S.lev[0].r)))
          return e2;
        if ((e2 = halo_rows(I, zq, 1, n1, rb, re, s_))) return e2;
        if ((e2 = st_zebra(nx, ny, 1, S.QQ, S.qlf, S.qiu, S.qcu, bq, z + V, 0, s_, rb, re, S.lev[0].r)))
          return e2;
        if (sw + 1 < S.q_sweeps && (e2 = halo_rows(I, zq, 1, n1, rb, re, s_))) return e2;
      }
      return e2;
    });
  };
  VecSlice sl{(long long)vb, (long long)V + vb, (long long)(ve - vb)};
  double bnorm = 0.0;
  rc = gmres_solve(2LL * V, A, Minv, S.bs, S.xs, rtol, atol, maxit, restart, 0, S.gm, iters, resid,
                   &bnorm, st, &sl, D);
  if (rc < 0) return rc;
  if (resid && bnorm > 0.0) *resid /= bnorm;
  int rc2 = st_interleave(V, S.xs, x, st, vb, ve);
  if (rc2) return rc2;
  if (rc == 1)
    set_error("cm_pqs_solve_dist: GMRES did not reach rtol=%g in %d iterations (relative residual %g)",
              rtol, maxit, resid ? *resid : -1.0);
  return rc;
}

// ------------------------------------------------------------------------------------------------
// Solver context (cm_pqs_create): second-generation path.  One object per workspace holds the layout, the
// knobs (read once), the strip neighbours and the captured CUDA graphs, so that repeated Newton steps replay
// the same graphs (the workspace pointers never change; setup only rewrites array contents).
//   preconditioner  V-cycle with fused half sweeps (cm_mg.cu): per level  c0 c1 | restricted residual |
//                   coarser levels | correction on odd rows | (c0 c1) x post;  from the first level with
//                   lines <= coarse_n1 downwards the whole cycle is one cluster kernel
//   multi-GPU       no exchange launches inside a preconditioner application: the half sweeps and the
//                   prolongation push the strip-boundary rows themselves, consumers wait on flags
//   GMRES           device resident (cm_gmres_dev.cu): one iteration = one graph launch + a 64-byte status
//                   read; one all-reduce per iteration
// The 13-point (C0IP) q-block and the triangular field splits run on the first-generation path.
// ------------------------------------------------------------------------------------------------
struct CachedGraph {
  cudaGraphExec_t exec = nullptr;
  long long launches = 0;
  double bytes = 0.0;
  void destroy() {
    if (exec) cudaGraphExecDestroy(exec);
    exec = nullptr;
  }
};

struct PQContext {
  PQKnobs K;
  PQSolver S;
  int nx, ny, restart;
  void* ws;
  size_t ws_bytes;
  double* base;
  bool dist = false;
  cm_dist D{};
  DistInfo I{};
  Halo H{};        // strip neighbours (inactive on one GPU), channel 0: p-block V-cycle
  Halo Hq{};       // channel 1: q-block sweeps (they run concurrently with the V-cycle on a second stream)
  Halo Hoff{};     // inactive: replicated levels
  cudaStream_t s2 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  int Lc = 0;      // first level of the global-memory fused coarse cycle (nlev: none)
  int Lsm = 0;     // first level of the shared-memory resident coarse cycle (nlev: none)
  int wide = 0;
  bool use_v1 = false;
  GmresDev G{};
  double* status_host = nullptr;   // pinned
  cudaStream_t cap = nullptr;
  CachedGraph g_iter, g_final;
  CachedGraph g_cycle;   // [condition] -> WHILE { one GMRES iteration } -> [solve y, x += Z y]
  CachedGraph g_setup;   // the whole setup behind a block-writing assembly (workspace pointers only: replayable)
  long long cycle_iter_launches = 0, cycle_fixed_launches = 0;
  double cycle_iter_bytes = 0.0, cycle_fixed_bytes = 0.0;
  bool while_ok = true;
  bool graphs_ok = true;
};

static void pqc_destroy(PQContext* C) {
  if (!C) return;
  C->g_iter.destroy();
  C->g_final.destroy();
  C->g_cycle.destroy();
  C->g_setup.destroy();
  if (C->cap) cudaStreamDestroy(C->cap);
  if (C->s2) cudaStreamDestroy(C->s2);
  if (C->ev_fork) cudaEventDestroy(C->ev_fork);
  if (C->ev_join) cudaEventDestroy(C->ev_join);
  if (C->status_host) cudaFreeHost(C->status_host);
  delete C;
}

void* pqs_create(int nx, int ny, int restart, void* ws, size_t ws_bytes, const cm_dist* D) {
  PQContext* C = new PQContext();
  C->K = read_knobs();
  C->nx = nx; C->ny = ny; C->restart = restart;
  C->ws = ws; C->ws_bytes = ws_bytes;
  C->base = align_ws(ws);
  if (pqs_layout(C->S, nx, ny, restart, C->base, C->K) != CM_OK) {
    set_error("cm_pqs_create: unsupported grid");
    delete C;
    return nullptr;
  }
  if (C->S.total_doubles * sizeof(double) + 64 > ws_bytes) {
    set_error("cm_pqs_create: workspace too small");
    delete C;
    return nullptr;
  }
  PQSolver& S = C->S;
  // fused coarse cycle: from the first level whose lines fit one warp each
  const int n1max = C->K.coarse_n1 < mg_coarse_max_n1() ? C->K.coarse_n1 : mg_coarse_max_n1();
  C->Lc = S.nlev;
  for (int l = 0; l < S.nlev; ++l)
    if (S.lev[l].nx + 1 <= n1max && S.nlev - l <= kCoarseMaxLevels && (l == S.nlev - 1 || (S.lev[l].ny & 1) == 0)) {
      C->Lc = l;
      break;
    }
  if (C->Lc == S.nlev - 1 && !S.dense_coarse) C->Lc = S.nlev;   // a lone coarsest level without dense inverse: sweeps
  C->Lsm = S.nlev;
  if (C->K.coarse_smem)
    for (int l = 0; l < S.nlev; ++l) {
      if (S.nlev - l > kCoarseMaxLevels) continue;
      CoarseCycle cc{};
      cc.nlev = S.nlev - l;
      for (int k = 0; k < cc.nlev; ++k) { cc.lev[k].nx = S.lev[l + k].nx; cc.lev[k].ny = S.lev[l + k].ny; }
      const size_t need = mg_coarse_smem_bytes(cc, S.dense_coarse);
      if (need > 0 && (l < S.nlev - 1 || S.dense_coarse)) {
        C->Lsm = l;
        break;
      }
    }
  if (D != nullptr && D->world > 1) {
    C->dist = true;
    C->D = *D;
    if (C->base != (double*)D->peer_base[D->rank]) {
      set_error("cm_pqs_create: workspace must be the (64-byte aligned) IPC allocation registered in cm_dist");
      delete C;
      return nullptr;
    }
    if (restart + 2 > 64) {
      set_error("cm_pqs_create: multi-GPU GMRES needs restart <= 62 (one all-reduce carries the Hessenberg column)");
      delete C;
      return nullptr;
    }
    // strips must start on even rows of every distributed level and divide exactly down to the first
    // replicated level (zebra colouring of the halo protocol, cm_halo.cuh): clamp Ld to the alignment present
    int Ld = S.dist_ld;
    if (Ld > C->Lc) Ld = C->Lc;
    if (Ld > C->Lsm) Ld = C->Lsm;
    if (Ld > S.nlev - 1) Ld = S.nlev - 1;
    for (int r = 1; r < D->world; ++r)
      while (Ld > 0 && (D->row_begin[r] & ((1 << Ld) - 1)) != 0) --Ld;
    for (int r = 0; r < D->world; ++r) {
      const int rows = D->row_begin[r + 1] - D->row_begin[r];
      while (Ld > 0 && (rows >> (Ld - 1)) < 4) --Ld;   // every distributed level keeps >= 4 rows per strip
      if ((r > 0 && (D->row_begin[r] & 1)) || rows < 4) {
        set_error("cm_pqs_create: strip %d starts at row %d with %d rows (need an even start and >= 4 rows)", r,
                  D->row_begin[r], rows);
        delete C;
        return nullptr;
      }
    }
    S.dist_ld = Ld;
    dist_info(S, &C->D, C->base, C->I);
    C->I.Ld = Ld;
    C->H = dist_halo_args(&C->D, 0);
    C->Hq = dist_halo_args(&C->D, 1);
  }
  if (C->K.q_concurrent) {
    int prio_least = 0, prio_greatest = 0;
    cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
    if (cudaStreamCreateWithPriority(&C->s2, cudaStreamNonBlocking, C->K.q_low_priority ? prio_least : prio_greatest) != cudaSuccess ||
        cudaEventCreateWithFlags(&C->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&C->ev_join, cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      C->K.q_concurrent = 0;
    }
  }
  if (mg_init_attributes() != CM_OK || mg_mid_init_attributes() != CM_OK || mg_coarse_smem_init_attributes() != CM_OK ||
      st_init_attributes() != CM_OK) {
    delete C;
    return nullptr;
  }
  gmres_dev_init();
  gmres_dev_layout(C->G, 2LL * S.nv, restart, S.gm);
  C->use_v1 = C->K.v1 != 0 || !S.diag_split || S.upper_split || S.pre_smooth < 1 || S.post_smooth < 1;
  if (cudaMallocHost((void**)&C->status_host, sizeof(double) * kGmresStatus) != cudaSuccess) {
    cudaGetLastError();
    set_error("cm_pqs_create: cudaMallocHost failed");
    delete C;
    return nullptr;
  }
  return C;
}

int pqs_destroy(void* ctx) {
  pqc_destroy(static_cast<PQContext*>(ctx));
  return CM_OK;
}

// fp32 copies of the operator streams the long-line half sweeps read (own rows of the distributed levels)
static int pqc_pack_f32(PQContext* C, cudaStream_t st) {
  PQSolver& S = C->S;
  if (!C->K.f32_op) return CM_OK;
  int rc;
  for (int l = 0; l < S.nlev; ++l) {
    PQSLevel& L = S.lev[l];
    if (!L.F) continue;
    const bool lev_dist = C->dist && l < C->I.Ld;
    if ((rc = mg_pack_f32(L.nx, L.ny, L.A, L.lf, L.iu, L.cu, L.F, lev_dist ? C->I.rb[l] : 0, lev_dist ? C->I.re[l] : -1, st)))
      return rc;
  }
  if (S.qF && (rc = mg_pack_f32(C->nx, C->ny, S.QQ, S.qlf, S.qiu, S.qcu, S.qF, C->dist ? C->I.rb[0] : 0,
                                C->dist ? C->I.re[0] : -1, st)))
    return rc;
  return CM_OK;
}

template <class F>
static int run_cached(PQContext* C, CachedGraph& g, cudaStream_t st, F&& body);

int pqs_ctx_setup(void* ctx, const int* nrowptr, const int* ncol, const double* vals, cudaStream_t st,
                  bool blocks_ready) {
  PQContext* C = static_cast<PQContext*>(ctx);
  PQSolver& S = C->S;
  int rc;
  if (blocks_ready && C->K.graph_setup) {
    // the assembly sweep wrote the stencil blocks: everything the setup touches lives in the workspace, so the
    // ~35 small launches (line factors, Galerkin products, packs, ghost rows of the operator, coarse inverse) are
    // one cached graph; only the error flag is read back afterwards
    rc = run_cached(C, C->g_setup, st, [&](cudaStream_t s_) -> int {
      int e = C->dist ? pqs_setup_dist(C->nx, C->ny, C->restart, nullptr, nullptr, nullptr, C->ws, C->ws_bytes, &C->D, s_,
                                       C->I.Ld, true, true)
                      : pqs_setup(C->nx, C->ny, C->restart, nullptr, nullptr, nullptr, C->ws, C->ws_bytes, s_, true, true);
      return e ? e : pqc_pack_f32(C, s_);
    });
    if (rc) return rc;
    C->wide = 0;
    return check_err_flag(S, st, "cm_pqs_ctx_setup");
  }
  if (!C->dist) {
    // same operator hierarchy as the first generation (the kernels of the cycle differ, the factors do not)
    if ((rc = pqs_setup(C->nx, C->ny, C->restart, nrowptr, ncol, vals, C->ws, C->ws_bytes, st, blocks_ready))) return rc;
    int wide = 0;
    if (!blocks_ready) {
      CM_CUDA(cudaMemcpyAsync(&wide, S.has_extra, sizeof(int), cudaMemcpyDeviceToHost, st));
      CM_CUDA(cudaStreamSynchronize(st));
    }
    C->wide = wide;
    return wide ? (int)CM_OK : pqc_pack_f32(C, st);
  }
  if ((rc = pqs_setup_dist(C->nx, C->ny, C->restart, nrowptr, ncol, vals, C->ws, C->ws_bytes, &C->D, st, C->I.Ld,
                           blocks_ready)))
    return rc;
  return pqc_pack_f32(C, st);
}

int pqs_ctx_blocks(void* ctx, cm_pq_blocks* out) {
  PQContext* C = static_cast<PQContext*>(ctx);
  const PQSolver& S = C->S;
  out->PP = S.PP; out->PQ = S.PQ; out->QP = S.QP; out->QQ = S.QQ; out->sp = S.sp; out->sq = S.sq;
  return CM_OK;
}

// one V-cycle on the p-block, level l: x = approx. A_l^{-1} b
static int vcycle2(PQContext* C, int l, const double* b, double* x, cudaStream_t st) {
  PQSolver& S = C->S;
  PQSLevel& L = S.lev[l];
  int rc;
  const bool lev_dist = C->dist && l < C->I.Ld;
  if (C->dist && l == C->I.Ld) {
    // first replicated level: every rank contributes its rows of the right-hand side
    unsigned long long off = (unsigned long long)(b - C->base) + (unsigned long long)C->I.rb[l] * (L.nx + 1);
    int n = (C->I.re[l] - C->I.rb[l]) * (L.nx + 1);
    if ((rc = dist_allgather(&C->D, &off, &n, 1, st))) return rc;
  }
  if (l >= C->Lc || l >= C->Lsm) {
    CoarseCycle cc{};
    cc.nlev = S.nlev - l;
    cc.pre = S.pre_smooth;
    cc.post = (C->K.post_coarse > 0 && l >= C->K.coarse_from) ? C->K.post_coarse
              : (l >= C->Lsm && C->Lsm >= 2 && C->K.post_smem < S.post_smooth ? C->K.post_smem : S.post_smooth);
    cc.Ainv = S.dense_coarse ? S.Ainv : nullptr;
    for (int k = 0; k < cc.nlev; ++k) {
      const PQSLevel& Lk = S.lev[l + k];
      cc.lev[k] = CoarseLevel{Lk.nx, Lk.ny, Lk.A, Lk.lf, Lk.iu, Lk.cu, k == 0 ? const_cast<double*>(b) : Lk.b,
                              k == 0 ? x : Lk.x};
    }
    if (l >= C->Lsm) return mg_coarse_smem_cycle(cc, st);
    return mg_coarse_cycle(cc, C->K.coarse_ncta, st);
  }
  const Halo& H = lev_dist ? C->H : C->Hoff;
  const int r0 = lev_dist ? C->I.rb[l] : 0, r1 = lev_dist ? C->I.re[l] : -1;
  if (l == S.nlev - 1) {
    if (S.dense_coarse) return st_coarse_apply((int)L.n, S.Ainv, b, x, st);
    for (int s = 0; s < 30; ++s) {
      if ((rc = mg_zebra(L.nx, L.ny, 0, L.A, L.lf, L.iu, L.cu, b, x, s == 0, H, r0, r1, st))) return rc;
      if ((rc = mg_zebra(L.nx, L.ny, 1, L.A, L.lf, L.iu, L.cu, b, x, 0, H, r0, r1, st))) return rc;
    }
    return CM_OK;
  }
  PQSLevel& Cl = S.lev[l + 1];
  const int npost = (C->K.post_coarse > 0 && l >= C->K.coarse_from) ? C->K.post_coarse : S.post_smooth;
  if (!lev_dist && L.nx + 1 <= C->K.mid_max_n1 && mg_mid_fits(L.nx, L.ny)) {
    // cache-resident middle level: latency-built half sweeps, transfers folded into them (cm_mgmid.cu)
    const bool fuse = C->K.mid_fuse != 0;
    const bool fuse_rr = fuse && mg_mid_fused_restrict_fits(L.nx, L.ny);
    for (int s = 0; s < S.pre_smooth; ++s) {
      if ((rc = mg_zebra_mid(s == 0 ? 0 : 1, L.nx, L.ny, 0, L.A, L.lf, L.iu, L.cu, b, x, nullptr, nullptr, st))) return rc;
      const bool last = s + 1 == S.pre_smooth;
      if ((rc = mg_zebra_mid(last && fuse_rr ? 2 : 1, L.nx, L.ny, 1, L.A, L.lf, L.iu, L.cu, b, x, nullptr, Cl.b, st)))
        return rc;
    }
    if (!fuse_rr && (rc = mg_resid_restrict(L.nx, L.ny, L.A, b, x, Cl.b, C->Hoff, 0, -1, st))) return rc;
    if ((rc = vcycle2(C, l + 1, Cl.b, Cl.x, st))) return rc;
    if (!fuse && (rc = mg_prolong_odd(Cl.nx, Cl.ny, Cl.x, x, C->Hoff, 0, 0, -1, st))) return rc;
    for (int s = 0; s < npost; ++s) {
      if ((rc = mg_zebra_mid(s == 0 && fuse ? 3 : 1, L.nx, L.ny, 0, L.A, L.lf, L.iu, L.cu, b, x, Cl.x, nullptr, st)))
        return rc;
      if ((rc = mg_zebra_mid(1, L.nx, L.ny, 1, L.A, L.lf, L.iu, L.cu, b, x, nullptr, nullptr, st))) return rc;
    }
    return CM_OK;
  }
  for (int s = 0; s < S.pre_smooth; ++s) {
    if ((rc = mg_zebra(L.nx, L.ny, 0, L.A, L.lf, L.iu, L.cu, b, x, s == 0, H, r0, r1, st, l == 0, C->K.f32_op ? L.F : nullptr))) return rc;
    if ((rc = mg_zebra(L.nx, L.ny, 1, L.A, L.lf, L.iu, L.cu, b, x, 0, H, r0, r1, st, l == 0, C->K.f32_op ? L.F : nullptr))) return rc;
  }
  if ((rc = mg_resid_restrict(L.nx, L.ny, L.A, b, x, Cl.b, H, lev_dist ? C->I.rb[l + 1] : 0,
                              lev_dist ? C->I.re[l + 1] : -1, st)))
    return rc;
  if ((rc = vcycle2(C, l + 1, Cl.b, Cl.x, st))) return rc;
  if ((rc = mg_prolong_odd(Cl.nx, Cl.ny, Cl.x, x, H, (lev_dist && l + 1 < C->I.Ld) ? 1 : 0, r0, r1, st))) return rc;
  for (int s = 0; s < npost; ++s) {
    if ((rc = mg_zebra(L.nx, L.ny, 0, L.A, L.lf, L.iu, L.cu, b, x, 0, H, r0, r1, st, l == 0, C->K.f32_op ? L.F : nullptr))) return rc;
    if ((rc = mg_zebra(L.nx, L.ny, 1, L.A, L.lf, L.iu, L.cu, b, x, 0, H, r0, r1, st, l == 0, C->K.f32_op ? L.F : nullptr))) return rc;
  }
  return CM_OK;
}

// z = diag(Mp, Mq)^{-1} ru on the unscaled blocks
static int precond2(PQContext* C, const double* ru, double* z, cudaStream_t st) {
  PQSolver& S = C->S;
  const int V = S.nv;
  if (C->K.precond_v1 && !C->dist) {
    int e = vcycle(S, 0, const_cast<double*>(ru), z, st);
    for (int sw = 0; sw < S.q_sweeps && !e; ++sw) {
      e = st_zebra(C->nx, C->ny, 0, S.QQ, S.qlf, S.qiu, S.qcu, ru + V, z + V, sw == 0 ? 1 : 0, st, 0, -1, S.lev[0].r);
      if (!e) e = st_zebra(C->nx, C->ny, 1, S.QQ, S.qlf, S.qiu, S.qcu, ru + V, z + V, 0, st, 0, -1, S.lev[0].r);
    }
    return e;
  }
  // the two blocks of the field split are independent: the q-sweeps (four launches on the finest grid) go to a
  // second stream and fill the SMs while the V-cycle works through its latency-bound coarse levels; multi-GPU:
  // they push / wait on their own halo channel
  const Halo& H = C->dist ? C->Hq : C->Hoff;
  const int r0 = C->dist ? C->I.rb[0] : 0, r1 = C->dist ? C->I.re[0] : -1;
  const bool fork = C->K.q_concurrent && !g_prof_is_on();
  cudaStream_t sq = st;
  int rc;
  if (fork) {
    CM_CUDA(cudaEventRecord(C->ev_fork, st));
    CM_CUDA(cudaStreamWaitEvent(C->s2, C->ev_fork, 0));
    sq = C->s2;
  } else if ((rc = vcycle2(C, 0, ru, z, st))) {
    return rc;
  }
  for (int sw = 0; sw < S.q_sweeps; ++sw) {
    if ((rc = mg_zebra(C->nx, C->ny, 0, S.QQ, S.qlf, S.qiu, S.qcu, ru + V, z + V, sw == 0, H, r0, r1, sq, true, C->K.f32_op ? S.qF : nullptr))) return rc;
    if ((rc = mg_zebra(C->nx, C->ny, 1, S.QQ, S.qlf, S.qiu, S.qcu, ru + V, z + V, 0, H, r0, r1, sq, true, C->K.f32_op ? S.qF : nullptr))) return rc;
  }
  if (fork) {
    CM_CUDA(cudaEventRecord(C->ev_join, C->s2));
    if ((rc = vcycle2(C, 0, ru, z, st))) return rc;
    CM_CUDA(cudaStreamWaitEvent(st, C->ev_join, 0));
  }
  return CM_OK;
}

// the capture stream: highest priority, so that the kernel nodes of the p-block V-cycle win the SMs over the
// q-block sweeps captured from the low-priority second stream (kernel nodes inherit the priority of the stream
// they were captured on)
static bool ensure_capture_stream(PQContext* C) {
  if (C->cap) return true;
  int prio_least = 0, prio_greatest = 0;
  cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
  if (cudaStreamCreateWithPriority(&C->cap, cudaStreamNonBlocking, prio_greatest) != cudaSuccess) {
    cudaGetLastError();
    C->cap = nullptr;
    return false;
  }
  return true;
}

// One GMRES cycle as ONE graph:  [condition kernel] -> WHILE { iteration } -> [child graph: final].
// `iter` and `fin` enqueue their kernels on the stream they are given; the last kernel of `iter` (the Givens
// kernel) sets the condition from the device-side status block, so no host round trip separates two iterations.
// Returns false (and leaves g_cycle empty) when the runtime refuses any step: the caller falls back to one
// graph launch + status read per iteration.
template <class FI, class FF>
static bool build_cycle_graph(PQContext* C, unsigned long long* handle_out, FI&& iter, FF&& fin) {
  if (!ensure_capture_stream(C)) return false;
  cudaGraph_t g = nullptr, gfin = nullptr;
  cudaGraphExec_t exec = nullptr;
  bool ok = false;
  do {
    if (cudaGraphCreate(&g, 0) != cudaSuccess) break;
    cudaGraphConditionalHandle h;
    if (cudaGraphConditionalHandleCreate(&h, g, 0, 0) != cudaSuccess) break;
    *handle_out = (unsigned long long)h;
    cudaGraphNode_t n_cond = nullptr, n_while = nullptr, n_fin = nullptr;
    if (gmres_dev_add_cond_node(g, &n_cond, C->G, (unsigned long long)h) != CM_OK) break;
    cudaGraphNodeParams wp{};
    wp.type = cudaGraphNodeTypeConditional;
    wp.conditional.handle = h;
    wp.conditional.type = cudaGraphCondTypeWhile;
    wp.conditional.size = 1;
    if (cudaGraphAddNode(&n_while, g, &n_cond, 1, &wp) != cudaSuccess) break;
    cudaGraph_t body = wp.conditional.phGraph_out[0];
    // the iteration, captured straight into the body graph
    if (cudaStreamBeginCaptureToGraph(C->cap, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed) != cudaSuccess) break;
    long long l0 = g_launches.load();
    double b0 = g_alg_bytes;
    int rc = iter(C->cap);
    cudaGraph_t dummy = nullptr;
    cudaError_t ce = cudaStreamEndCapture(C->cap, &dummy);
    C->cycle_iter_launches = g_launches.load() - l0;
    C->cycle_iter_bytes = g_alg_bytes - b0;
    g_launches -= C->cycle_iter_launches;   // nothing has run yet
    g_alg_bytes -= C->cycle_iter_bytes;
    if (rc != CM_OK || ce != cudaSuccess) break;
    // the final part (y = R^-1 g, x += Z y) as a child graph behind the loop
    if (cudaStreamBeginCapture(C->cap, cudaStreamCaptureModeRelaxed) != cudaSuccess) break;
    l0 = g_launches.load();
    b0 = g_alg_bytes;
    rc = fin(C->cap);
    ce = cudaStreamEndCapture(C->cap, &gfin);
    C->cycle_fixed_launches = g_launches.load() - l0 + 1;   // + the condition kernel
    C->cycle_fixed_bytes = g_alg_bytes - b0;
    g_launches -= C->cycle_fixed_launches - 1;
    g_alg_bytes -= C->cycle_fixed_bytes;
    if (rc != CM_OK || ce != cudaSuccess || gfin == nullptr) break;
    if (cudaGraphAddChildGraphNode(&n_fin, g, &n_while, 1, gfin) != cudaSuccess) break;
    if (cudaGraphInstantiate(&exec, g, 0) != cudaSuccess) break;
    ok = true;
  } while (false);
  cudaGetLastError();
  if (gfin) cudaGraphDestroy(gfin);
  if (g) cudaGraphDestroy(g);
  if (!ok) {
    if (exec) cudaGraphExecDestroy(exec);
    // a failed capture may leave the stream capturing: end it
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(C->cap, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone) {
      cudaGraph_t junk = nullptr;
      cudaStreamEndCapture(C->cap, &junk);
      if (junk) cudaGraphDestroy(junk);
    }
    cudaGetLastError();
    return false;
  }
  C->g_cycle.exec = exec;
  return true;
}

// run `body` on the stream: replayed from a cached CUDA graph when allowed
template <class F>
static int run_cached(PQContext* C, CachedGraph& g, cudaStream_t st, F&& body) {
  if (!C->K.graph || !C->graphs_ok || g_prof_is_on()) return body(st);
  if (!g.exec) {
    if (!ensure_capture_stream(C)) {
      C->graphs_ok = false;
      return body(st);
    }
    if (cudaStreamBeginCapture(C->cap, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
      cudaGetLastError();
      C->graphs_ok = false;
      return body(st);
    }
    const long long l0 = g_launches.load();
    const double b0 = g_alg_bytes;
    const int rc = body(C->cap);
    cudaGraph_t graph = nullptr;
    const cudaError_t ce = cudaStreamEndCapture(C->cap, &graph);
    g.launches = g_launches.load() - l0;
    g.bytes = g_alg_bytes - b0;
    g_launches -= g.launches;   // nothing has run yet
    g_alg_bytes -= g.bytes;
    if (rc != CM_OK || ce != cudaSuccess || graph == nullptr) {
      cudaGetLastError();
      if (graph) cudaGraphDestroy(graph);
      C->graphs_ok = false;
      return rc != CM_OK ? rc : body(st);
    }
    if (cudaGraphInstantiate(&g.exec, graph, 0) != cudaSuccess) {
      cudaGetLastError();
      cudaGraphDestroy(graph);
      g.exec = nullptr;
      C->graphs_ok = false;
      return body(st);
    }
    cudaGraphDestroy(graph);
  }
  g_launches += g.launches;
  g_alg_bytes += g.bytes;
  if (cudaGraphLaunch(g.exec, st) != cudaSuccess) {
    set_error("cm_pqs_solve: cudaGraphLaunch failed: %s", cudaGetErrorString(cudaGetLastError()));
    return CM_E_CUDA;
  }
  return CM_OK;
}

int pqs_ctx_solve(void* ctx, const double* b, double* x, double rtol, double atol, int maxit, int* iters,
                  double* resid, cudaStream_t st) {
  PQContext* C = static_cast<PQContext*>(ctx);
  PQSolver& S = C->S;
  if (C->use_v1 || C->wide) {
    if (C->dist)
      return pqs_solve_dist(C->nx, C->ny, C->restart, b, x, rtol, atol, maxit, C->ws, C->ws_bytes, &C->D, iters,
                            resid, st);
    return pqs_solve(C->nx, C->ny, C->restart, b, x, rtol, atol, maxit, C->ws, C->ws_bytes, iters, resid, st);
  }
  const int nx = C->nx, ny = C->ny, V = S.nv, n1 = nx + 1, m = C->restart;
  const int rb = C->dist ? C->I.rb[0] : 0, re = C->dist ? C->I.re[0] : ny + 1;
  const int vb = rb * n1, ve = re * n1;
  VecSlice slv{(long long)vb, (long long)V + vb, (long long)(ve - vb)};
  const VecSlice* sl = C->dist ? &slv : nullptr;
  const GmresDev& G = C->G;
  const Halo* Hp = C->dist ? &C->H : nullptr;
  cm_dist* Dp = C->dist ? &C->D : nullptr;
  int rc;
  double* hs = C->status_host;
  auto status = [&]() -> int {   // device status block -> pinned host copy, wait
    CM_CUDA(cudaMemcpyAsync(hs, G.st, sizeof(double) * kGmresStatus, cudaMemcpyDeviceToHost, st));
    CM_CUDA(cudaStreamSynchronize(st));
    return CM_OK;
  };
  // out = A in; keep_z: also store `in` as the preconditioned basis vector Z_j (flexible GMRES)
  const bool fuse_unscale = C->K.fuse_unscale != 0;
  auto apply_A = [&](const double* in, double* out, cudaStream_t s_, bool keep_z = false) -> int {
    return st_pq_spmv(nx, ny, S.PP, S.PQ, S.QP, S.QQ, S.sp, S.sq, in, out, s_, rb, C->dist ? re : -1, nullptr,
                      nullptr, Hp, keep_z ? G.Z : nullptr, G.ne, G.st + kStJ, 2,
                      (keep_z && fuse_unscale) ? G.vscale : nullptr);
  };
  // Gram-Schmidt against the j+1 basis vectors, new (unnormalised) basis vector, its norm, scalar update
  auto orthogonalise = [&](int jhost, cudaStream_t s_, unsigned long long cond = 0ull) -> int {
    int e = gmres_dev_dot(G, sl, jhost, s_);
    if (!e && jhost >= 0) e = dist_allreduce(Dp, G.dots, m + 1, s_);
    if (!e) e = fuse_unscale ? gmres_dev_update(G, sl, jhost, s_, S.sp, S.sq, S.ru) : gmres_dev_update(G, sl, jhost, s_);
    if (!e) e = dist_allreduce(Dp, G.nrm, 1, s_);
    if (!e) e = gmres_dev_givens(G, rtol, atol, s_, cond);
    return e;
  };
  // one GMRES iteration / the end of a cycle, as enqueued on a stream (directly, or into a graph)
  auto iteration = [&](cudaStream_t s_, unsigned long long cond) -> int {
    // the preconditioner input ru = D^-1 V_j comes out of the previous update kernel (unnormalised: apply_A scales)
    int e = fuse_unscale ? (int)CM_OK
                         : st_unscale_dev(nx, ny, G.V, G.ne, G.st + kStJ, G.vscale, S.sp, S.sq, S.ru, s_, rb, C->dist ? re : -1);
    if (!e) e = precond2(C, S.ru, G.z, s_);
    if (!e) e = apply_A(G.z, G.w, s_, true);
    if (!e) e = orthogonalise(4, s_, cond);     // byte accounting: a typical mid-cycle iteration, corrected by the caller
    return e;
  };
  auto cycle_end = [&](cudaStream_t s_) -> int {   // x += Z y with Z_k = M^{-1} v_k kept from the iterations
    int e = gmres_dev_solve_y(G, s_);
    if (!e) e = gmres_dev_combine(G, sl, S.xs, 8, s_);
    return e;
  };
  auto normalise = [&](int jhost) -> int { return orthogonalise(jhost, st); };
  if ((rc = st_split_scale(V, b, S.sp, S.sq, S.bs, st, vb, ve))) return rc;
  if (C->K.gmres_host && !C->dist) {
    LinOp A = [&](const double* in, double* out) { return apply_A(in, out, st, false); };
    LinOp Minv = [&](const double* r, double* z) {
      int e = st_unscale(nx, ny, r, S.sp, S.sq, S.ru, st);
      return e ? e : precond2(C, S.ru, z, st);
    };
    double bn = 0.0;
    rc = gmres_solve(2LL * V, A, Minv, S.bs, S.xs, rtol, atol, maxit, m, 0, S.gm, iters, resid, &bn, st);
    if (rc < 0) return rc;
    if (resid && bn > 0.0) *resid /= bn;
    int rc2 = st_interleave(V, S.xs, x, st);
    return rc2 ? rc2 : rc;
  }
  if ((rc = gmres_dev_axpby(G, sl, 0.0, nullptr, 0.0, S.xs, st))) return rc;      // x0 = 0
  if ((rc = gmres_dev_axpby(G, sl, 1.0, S.bs, 0.0, G.w, st))) return rc;          // r0 = b
  if ((rc = gmres_dev_begin(G, 1, st, maxit))) return rc;
  if ((rc = normalise(-1))) return rc;
  if ((rc = status())) return rc;
  if (hs[kStFlag] == 2.0) {
    set_error("cm_pqs_solve: right-hand side norm is not finite");
    return CM_E_BREAKDOWN;
  }
  const double bnorm = hs[kStBnorm];
  bool converged = hs[kStConv] != 0.0;
  int its = 0;
  double res = hs[kStResid];
  while (!converged && its < maxit) {
    // one restart cycle
    int j = 0;
    bool stop = false;
    const bool try_while = C->K.graph && C->K.graph_while && C->graphs_ok && C->while_ok && !g_prof_is_on();
    if (try_while && !C->g_cycle.exec) {
      unsigned long long h = 0ull;
      auto it_body = [&](cudaStream_t s_) -> int { return iteration(s_, h); };
      if (!build_cycle_graph(C, &h, it_body, cycle_end)) C->while_ok = false;
    }
    if (try_while && C->g_cycle.exec) {
      // the whole cycle is one launch: the WHILE node repeats the iteration until the device-side status says stop
      const int its0 = its;
      if (cudaGraphLaunch(C->g_cycle.exec, st) != cudaSuccess) {
        set_error("cm_pqs_solve: cudaGraphLaunch (GMRES cycle) failed: %s", cudaGetErrorString(cudaGetLastError()));
        return CM_E_CUDA;
      }
      if ((rc = status())) return rc;
      if (hs[kStFlag] == 2.0) {
        set_error("cm_pqs_solve: NaN/Inf in the Arnoldi process at iteration %d", (int)hs[kStIts]);
        return CM_E_BREAKDOWN;
      }
      its = (int)hs[kStIts];
      j = (int)hs[kStJ];
      res = hs[kStResid];
      converged = hs[kStConv] != 0.0;
      const int nit = its - its0;
      g_launches += C->cycle_fixed_launches + (long long)nit * C->cycle_iter_launches;
      g_alg_bytes += C->cycle_fixed_bytes + (double)nit * C->cycle_iter_bytes;
      // the captured iteration was accounted with j = 4, the final combine with 8 vectors: correct both
      const double vb8 = (double)(sl ? 2 * sl->len : G.n) * 8.0;
      for (int jj = 0; jj < nit; ++jj) CM_BYTES(vb8 * 2.0 * ((double)jj - 4.0));
      CM_BYTES(vb8 * ((double)j - 8.0));
    } else {
      while (!stop) {
        rc = run_cached(C, C->g_iter, st, [&](cudaStream_t s_) -> int { return iteration(s_, 0ull); });
        if (rc) return rc;
        if ((rc = status())) return rc;
        if (hs[kStFlag] == 2.0) {
          set_error("cm_pqs_solve: NaN/Inf in the Arnoldi process at iteration %d", its);
          return CM_E_BREAKDOWN;
        }
        its = (int)hs[kStIts];
        j = (int)hs[kStJ];
        // the captured sequence was accounted with j = 4: correct the two j-dependent kernels (dot: j+2, update: j+3
        // vector passes for the iteration index j-1 that just ran)
        CM_BYTES((double)(sl ? 2 * sl->len : G.n) * 8.0 * 2.0 * ((double)(j - 1) - 4.0));
        res = hs[kStResid];
        converged = hs[kStConv] != 0.0;
        stop = converged || its >= maxit || j >= m || hs[kStFlag] == 1.0;
      }
      // x += Z y with Z_k = M^{-1} v_k kept from the iterations: no further preconditioner application
      rc = run_cached(C, C->g_final, st, cycle_end);
      if (rc) return rc;
    }
    if (converged || its >= maxit) break;
    // restart: w = b - A x, normalised into v_0
    if (C->dist) {
      double* a2[2] = {S.xs, S.xs + V};
      if ((rc = halo_rows(C->I, a2, 2, n1, rb, re, st))) return rc;
    }
    if ((rc = apply_A(S.xs, G.z, st))) return rc;
    if ((rc = gmres_dev_axpby(G, sl, 1.0, S.bs, 0.0, G.w, st))) return rc;
    if ((rc = gmres_dev_axpby(G, sl, -1.0, G.z, 1.0, G.w, st))) return rc;
    if ((rc = gmres_dev_begin(G, 0, st))) return rc;
    if ((rc = normalise(-1))) return rc;
    if ((rc = status())) return rc;
    res = hs[kStResid];
    converged = hs[kStConv] != 0.0;
  }
  if ((rc = st_interleave(V, S.xs, x, st, vb, ve))) return rc;
  CM_CUDA(cudaStreamSynchronize(st));
  if (iters) *iters = its;
  if (resid) *resid = bnorm > 0.0 ? res / bnorm : res;
  if (!converged) {
    set_error("cm_pqs_solve: GMRES did not reach rtol=%g in %d iterations (relative residual %g)", rtol, maxit,
              bnorm > 0.0 ? res / bnorm : res);
    return 1;
  }
  return CM_OK;
}

int pqs_ctx_info(void* ctx, int* out8) {
  PQContext* C = static_cast<PQContext*>(ctx);
  out8[0] = C->S.nlev;
  out8[1] = C->Lsm < C->Lc ? C->Lsm : C->Lc;
  out8[2] = C->dist ? C->I.Ld : -1;
  out8[3] = (C->use_v1 || C->wide) ? 1 : 2;
  out8[4] = C->g_iter.exec ? (int)C->g_iter.launches : (C->g_cycle.exec ? (int)C->cycle_iter_launches : 0);
  out8[5] = C->K.coarse_ncta;
  out8[6] = C->graphs_ok ? 1 : 0;
  out8[7] = C->wide;
  return CM_OK;
}

size_t pqs_state_offset_bytes(int nx, int ny, int restart) {
  PQSolver S;
  if (pqs_layout(S, nx, ny, restart, (double*)64, read_knobs())) return 0;
  return (size_t)((char*)S.ustate - (char*)64);
}

// ------------------------------------------------------------------------------------------------
// generic path: Jacobi-preconditioned GMRES on the node-level CSR layouts (bs = 1 or 2)
// ------------------------------------------------------------------------------------------------
__global__ void jacobi_setup_kernel(const int bs, const int nv, const int* __restrict__ nrowptr,
                                    const int* __restrict__ diag_slot, const double* __restrict__ vals,
                                    double* __restrict__ dinv) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  const int rp = nrowptr[v], deg = nrowptr[v + 1] - rp, k = diag_slot[v];
  if (bs == 1) {
    const double d = vals[rp + k];
    dinv[v] = d != 0.0 ? 1.0 / d : 1.0;
  } else {
    // 2x2 node block [[pp,pq],[qp,qq]] inverted (identity when singular)
    const size_t base = 4 * (size_t)rp;
    const double a = vals[base + 2 * k], b = vals[base + 2 * k + 1];
    const double c = vals[base + 2 * deg + 2 * k], d = vals[base + 2 * deg + 2 * k + 1];
    const double det = a * d - b * c;
    const double scale = fmax(fmax(fabs(a), fabs(b)), fmax(fabs(c), fabs(d)));
    if (fabs(det) > 1e-14 * scale * scale && scale > 0.0) {
      const double id = 1.0 / det;
      dinv[4 * (size_t)v] = d * id; dinv[4 * (size_t)v + 1] = -b * id;
      dinv[4 * (size_t)v + 2] = -c * id; dinv[4 * (size_t)v + 3] = a * id;
    } else {
      dinv[4 * (size_t)v] = 1.0; dinv[4 * (size_t)v + 1] = 0.0;
      dinv[4 * (size_t)v + 2] = 0.0; dinv[4 * (size_t)v + 3] = 1.0;
    }
  }
}

__global__ void jacobi_apply_kernel(const int bs, const int nv, const double* __restrict__ dinv,
                                    const double* __restrict__ r, double* __restrict__ z) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nv) return;
  if (bs == 1) {
    z[v] = dinv[v] * r[v];
  } else {
    const double r0 = r[2 * (size_t)v], r1 = r[2 * (size_t)v + 1];
    const double* m = dinv + 4 * (size_t)v;
    z[2 * (size_t)v] = m[0] * r0 + m[1] * r1;
    z[2 * (size_t)v + 1] = m[2] * r0 + m[3] * r1;
  }
}

size_t krylov_workspace_bytes(int bs, int nv, int restart) {
  const long long n = (long long)bs * nv;
  return (gmres_workspace_doubles(n, restart) + (size_t)bs * bs * nv + 8) * sizeof(double) + 64;
}

int krylov_solve(int bs, int nv, const int* nrowptr, const int* ncol, const int* diag_slot,
                 const double* vals, const double* b, double* x, double rtol, double atol, int maxit,
                 int restart, void* ws, size_t ws_bytes, int* iters, double* resid,
                 cudaStream_t st) {
  if (krylov_workspace_bytes(bs, nv, restart) > ws_bytes) {
    set_error("krylov_solve: workspace too small");
    return CM_E_INVALID;
  }
  double* base = align_ws(ws);
  double* dinv = base;
  double* gm = base + even((size_t)bs * bs * nv);
  const int grid = (nv + 255) / 256;
  jacobi_setup_kernel<<<grid, 256, 0, st>>>(bs, nv, nrowptr, diag_slot, vals, dinv);
  CM_LAUNCH_CHECK();
  LinOp A = [&](const double* in, double* out) {
    return spmv(bs, nv, nrowptr, ncol, vals, in, out, 1.0, nullptr, st);
  };
  LinOp Minv = [&](const double* r, double* z) {
    jacobi_apply_kernel<<<grid, 256, 0, st>>>(bs, nv, dinv, r, z);
    ++g_launches;
    return cudaGetLastError() == cudaSuccess ? CM_OK : CM_E_CUDA;
  };
  double bnorm = 0.0;
  int rc = gmres_solve((long long)bs * nv, A, Minv, b, x, rtol, atol, maxit, restart, 0, gm, iters,
                       resid, &bnorm, st);
  if (resid && bnorm > 0.0) *resid /= bnorm;
  if (rc == 1)
    set_error("cm_krylov_solve: GMRES did not reach rtol=%g in %d iterations (relative residual %g)",
              rtol, maxit, resid ? *resid : -1.0);
  return rc;
}

size_t bicgstab_workspace_bytes(int bs, int nv) {
  const long long n = (long long)bs * nv;
  return (bicgstab_workspace_doubles(n) + (size_t)bs * bs * nv + 8) * sizeof(double) + 64;
}

// BiCGStab + (block-)Jacobi on the node CSR layouts: same operator / preconditioner as krylov_solve.
int bicgstab_csr_solve(int bs, int nv, const int* nrowptr, const int* ncol, const int* diag_slot,
                       const double* vals, const double* b, double* x, double rtol, double atol, int maxit,
                       void* ws, size_t ws_bytes, int* iters, double* resid, cudaStream_t st) {
  if (bicgstab_workspace_bytes(bs, nv) > ws_bytes) {
    set_error("bicgstab_solve: workspace too small");
    return CM_E_INVALID;
  }
  double* base = align_ws(ws);
  double* dinv = base;
  double* wk = base + even((size_t)bs * bs * nv);
  const int grid = (nv + 255) / 256;
  jacobi_setup_kernel<<<grid, 256, 0, st>>>(bs, nv, nrowptr, diag_slot, vals, dinv);
  CM_LAUNCH_CHECK();
  LinOp A = [&](const double* in, double* out) {
    return spmv(bs, nv, nrowptr, ncol, vals, in, out, 1.0, nullptr, st);
  };
  LinOp Minv = [&](const double* r, double* z) {
    jacobi_apply_kernel<<<grid, 256, 0, st>>>(bs, nv, dinv, r, z);
    ++g_launches;
    return cudaGetLastError() == cudaSuccess ? CM_OK : CM_E_CUDA;
  };
  double bnorm = 0.0;
  int rc = bicgstab_solve((long long)bs * nv, A, Minv, b, x, rtol, atol, maxit, wk, iters, resid, &bnorm, st);
  if (resid && bnorm > 0.0) *resid /= bnorm;
  if (rc == 1)
    set_error("cm_bicgstab_solve: BiCGStab did not reach rtol=%g in %d iterations (relative residual %g)",
              rtol, maxit, resid ? *resid : -1.0);
  return rc;
}

}  // namespace cm
