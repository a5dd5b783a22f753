/*
 * closestmol_b200.h -- C ABI of the B200-native closestMolecule FEM hot path.
 *
 * The reference (ruizbaier/closestMolecule) has no FFI layer of its own: its scripts program
 * against the FEniCS 2019.1 Python API, and underneath that against the UFC/DOLFIN C++ interfaces
 * (tabulate_tensor / GenericTensor::add_local / DirichletBC::apply / GenericLinearSolver::solve).
 * Every entry point below names the reference call site (file:line under the reference checkout)
 * whose FEniCS-side work it replaces.  All functions are batched over the whole mesh.
 *
 * Conventions
 *   - every pointer is a CUDA DEVICE pointer owned by the caller unless the name ends in _host;
 *     the library never allocates or frees caller memory; workspaces are caller-provided after a
 *     *_workspace_bytes query;
 *   - calls are asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant across
 *     streams, no global mutable state;
 *   - return value: 0 = ok, <0 = error (CM_E_*), message via cm_last_error() (thread-local);
 *   - indices int32, values fp64.  Dof layout of the CG1xCG1 space: dof = 2*vertex + component
 *     (DOLFIN with reorder_dofs_serial=False).  The P-Q Jacobian value array is laid out exactly
 *     as the PETSc AIJ CSR DOLFIN produces: for node row v with sorted neighbours w_0<...<w_{d-1}
 *     (node-level rowptr/col), scalar row 2v+c starts at 4*rowptr[v] + c*2*d and holds
 *     (w_0,p),(w_0,q),(w_1,p),(w_1,q),...  -- so only the node-level index is ever stored.
 */
#ifndef CLOSESTMOL_B200_H
#define CLOSESTMOL_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CM_OK 0
#define CM_E_INVALID (-1)   /* bad argument */
#define CM_E_CUDA (-2)      /* CUDA runtime error */
#define CM_E_UNSUPPORTED (-3)
#define CM_E_BREAKDOWN (-4) /* Krylov breakdown / NaN */

/* nonlinearity kinds (cm_pq_params.gkind) */
#define CM_G_NONE 0
#define CM_G_POWER 1  /* gcoef*x^2*p^2/(q+geps): MixedPrimal_convergence.py:43-44, ..._C0IP.py:46-47 */
#define CM_G_GSTAR 2  /* conditional Gstar: advection_diffusion_convergence.py:21-27,
                         advection_diffusion_mixed.py:17-23 */

/* exterior-facet kind bits (per facet) */
#define CM_EF_ADV 1     /* + (r2vec.n) q w W ds            ..._C0IP.py:146 */
#define CM_EF_DATA 2    /* + (r2vec.n) q_ex w W ds         test_for_paper_convergence.py:152 */
#define CM_EF_PENALTY 4 /* + stab2 (hK neg(r2vec.n)^2 + 8/hK)(q-q_ex) w ds   ...:145,153 */

/* Parameters of the P-Q primal form family (host struct, passed by pointer).
 * p-equation: (p-p_old)*inv_dt*v1*W + D2*(dp/dx + G)*dv1/dx*W + D1*dp/dy*dv1/dy*W
 *             (advection_diffusion_convergence.py:196-204)
 * q-equation: [q_react*q + q_adv*dq/dx + x^2*(q_x2p*p + q_x2q*q) - q_divr*(2/x)*q]*(w + tau*dw/dx)*W
 *             - q_ibp*q*dw/dx*W,   tau = supg_stab*hK/2
 *             (Galerkin: ...:203; SUPG: ..._SUPG.py:118-129; C0IP cell terms: ..._C0IP.py:144-145)
 * W = (4*pi*x*y)^2 (advection_diffusion_convergence.py:135), x = r2 = coordinate 0. */
typedef struct cm_pq_params {
  double D1, D2;
  double inv_dt;
  double gcoef, geps;
  double gs_thr, gs_cond_scale, gs_lin;
  double q_react, q_adv, q_x2p, q_x2q, q_divr, q_ibp;
  double supg_stab;
  double c0ip_gamma; /* sign*stab/(k_q)^3.5*beta, ..._C0IP.py:147 */
  double pen_stab2;  /* test_for_paper_convergence.py:145 */
  double r2x, r2y;   /* r2vec */
  int32_t gkind;
  int32_t degree;    /* quadrature degree: 2, 4 (6-point FIAT rule) or 6 (12-point) */
  int32_t c0ip_h;    /* 0: avg(CellDiameter), 1: avg(FacetArea) */
  int32_t reserved;
} cm_pq_params;

/* Parameters of the scalar P1 forms.
 * F = u*v + D2*du/dx*dv/dx + D1*du/dy*dv/dy - (2*D2/x)*du/dx*v - (2*D1/y)*du/dy*v
 *     + g(u)*dv/dx - (2/x)*g(u)*v,   g = gcoef*x^2*u^2
 * (SphericalAdvectionDiffusionReaction_convergence.py:92-100; gcoef=0, D1=D2 gives
 *  SphericalLaplace_convergence.py:128). */
typedef struct cm_scalar_params {
  double D1, D2, gcoef;
  int32_t degree;
  int32_t reserved;
} cm_scalar_params;

/* Strip-partitioned multi-GPU context (host struct owned by the caller, one process per GPU).
 * Every rank allocates the same IPC-shared workspace (cm_ipc_alloc) and opens its peers'
 * (cm_ipc_open); arrays are GLOBALLY indexed, rank r owns vertex rows [row_begin[r], row_begin[r+1])
 * of the fine grid (boundaries multiples of 8 so that three multigrid levels stay aligned).
 * halo_seq / coll_seq are sequence counters the library advances; all ranks must make the same
 * sequence of library calls. */
struct cm_dist;
typedef struct cm_dist {
  int32_t rank, world;
  void* peer_base[8];      /* workspace base of every rank as mapped in THIS process (own at [rank]) */
  uint64_t halo_seq, coll_seq;
  int32_t row_begin[9];
  int32_t reserved;
} cm_dist;

int cm_version(void);
const char* cm_last_error(void);
/* number of CUDA kernels this library has launched so far in this process (diagnostics) */
int64_t cm_launch_count(void);
/* algorithmic (compulsory) bytes of all kernels launched so far: every array counted once per pass,
 * index/gather tables and cache re-reads not counted */
double cm_algorithmic_bytes(void);
/* Diagnostics: CUDA-event timing of selected kernel classes on the launching stream
 * (0: zebra_line_kernel on the finest grid, 1: pq_stencil_spmv_kernel, 2: Jacobian assembly).
 * Process-wide state, not thread safe; off by default. */
int cm_profile_enable(int32_t on);
int cm_profile_report(int32_t cls, int64_t* launches, double* total_ms, double* total_bytes);
int cm_profile_reset(void);
/* 1 if a CUDA device is usable, 0 otherwise (never launches work) */
int cm_device_available(void);

/* ---- a3/a5/a6(i): fused Jacobian+residual of the P-Q cell integrals, unstructured meshes --------
 * Replaces DOLFIN Assembler::assemble_cells + FFC tabulate_tensor for FF and Tang, triggered by
 * solver.solve() at advection_diffusion_convergence.py:207-219 (…MixedPrimal_convergence.py:132-140).
 * Row-owner gather: one thread per vertex accumulates its incident cells in ascending cell order
 * (DOLFIN's serial summation order), every CSR slot is written exactly once, no atomics.
 *   v2c_ptr[nv+1], v2c_ent[3*ncells]: vertex -> incident cells, entry = 4*cell + local index
 *   v2c_slot[3*ncells]: bytes (k0,k1,k2) = slots of (v, next, next-next cell vertex) in row v
 *   nrowptr[nv+1]: node-level row pointer (cell-only or cell+interior-facet pattern)
 *   ncol[nnz_node] node-level columns (only read when Kq is given), maxdeg = max row degree
 *   u[2*nv] state, p_old[nv] or NULL, load[2*nv] or NULL (state-independent forcing, subtracted)
 *   Kq[nnz_node], cq[nv] or NULL: constant facet operator of the C0IP q-equation (see below):
 *   J_qq += Kq, F_q += Kq*q + cq, fused into the same pass
 *   vals[4*nnz_node] or NULL (residual only), b[2*nv]
 *   accumulate: 0 = overwrite vals/b, 1 = add to existing content (facet terms assembled before) */
int cm_assemble_pq_cells(const cm_pq_params* prm_host, int32_t nv, int32_t ncells,
                         const int32_t* cells, const double* coords, const int32_t* v2c_ptr,
                         const int32_t* v2c_ent, const int32_t* v2c_slot, const int32_t* nrowptr,
                         const int32_t* ncol, int32_t maxdeg, const double* u, const double* p_old,
                         const double* load, const double* Kq, const double* cq, double* vals,
                         double* b, int32_t accumulate, void* stream);

/* ---- a3 on structured meshes (RectangleMesh numbering, create_flat_boundary.py:65; arbitrary
 * vertex coordinates): same output, bit-identical layout, no connectivity or map is read: cells,
 * neighbours and CSR slots are derived from (ix,iy).  Each cell's quadrature sums are computed
 * once; element rows meet in shared-memory accumulators in DOLFIN's summation order (no atomics);
 * every CSR slot is written exactly once.  nrowptr must be the cell-only P1 pattern of the (nx,ny) mesh. */
int cm_assemble_pq_cells_structured(const cm_pq_params* prm_host, int32_t nx, int32_t ny,
                                    const double* coords, const int32_t* nrowptr, const double* u,
                                    const double* p_old, const double* load, double* vals,
                                    double* b, void* stream);

/* ---- a4: scalar P1 forms (F-A/F-B), Jacobian+residual, row-owner gather -------------------------
 * Replaces assemble of FF/Tang at SphericalAdvectionDiffusionReaction_convergence.py:92-110 and of
 * auv at SphericalLaplace_convergence.py:128-135.  vals[nnz_node] in node CSR order. */
int cm_assemble_scalar_cells(const cm_scalar_params* prm_host, int32_t nv, int32_t ncells,
                             const int32_t* cells, const double* coords, const int32_t* v2c_ptr,
                             const int32_t* v2c_ent, const int32_t* v2c_slot,
                             const int32_t* nrowptr, int32_t maxdeg, const double* u,
                             const double* load, double* vals, double* b, void* stream);

/* ---- a1/a12: state-independent load vector  load[i] = sum_cells int f*phi_i (*W) -----------------
 * fq[ncells*nq*ncomp] holds the forcing at the quadrature points (cell-major, then point, then
 * component).  ncomp=2 (P-Q: f2 against v1; f3 against w + tau*dw/dx), ncomp=1 scalar.
 * Replaces the f2_ex/f3_ex terms of …MixedPrimal_convergence.py:126-128 (assembled once per mesh
 * instead of once per residual evaluation). */
int cm_assemble_load(int32_t ncomp, int32_t weighted, double supg_stab, int32_t degree, int32_t nv,
                     int32_t ncells, const int32_t* cells, const double* coords,
                     const int32_t* v2c_ptr, const int32_t* v2c_ent, const double* fq, double* load,
                     void* stream);

/* ---- a6(ii,iii): facet integrals of the C0IP q-equation -----------------------------------------
 * State independent (the q-equation is linear): assembled into a constant operator Kq on the
 * node-level (cell+interior-facet) pattern and a constant vector cq, so that
 *   J_qq += Kq,  F_q += Kq*q + cq.
 * Interior: c0ip_gamma*avg(hK)^2*(jump(grad q).n+)(jump(grad w).n+)*W dS (..._C0IP.py:147).
 * Exterior: per-facet kind bits CM_EF_*; qdata[nfe*nqf] = q_ex at the facet Gauss points.
 *   v2f_ptr[nv+1], v2f_ent[...]: vertex -> facets whose macro element contains the vertex
 *   (ascending facet index); facet tables: fverts[nf*2], fcell[nf*2] (cell1=-1 on the boundary).
 *   exterior-facet tables: efverts[nfe*2], efopp[nfe] (vertex of the attached cell opposite the
 *   facet, orients the normal), efkind[nfe]; v2e_ptr/v2e_ent: vertex -> exterior facets (index
 *   into those arrays) it is an end point of.  Kq and cq must be zeroed by the caller. */
int cm_assemble_c0ip_facets(const cm_pq_params* prm_host, int32_t nv, int32_t nf,
                            const int32_t* cells, const double* coords, const int32_t* fverts,
                            const int32_t* fcell, const int32_t* v2f_ptr, const int32_t* v2f_ent,
                            const int32_t* nrowptr, const int32_t* ncol, double* Kq, void* stream);
int cm_assemble_exterior_facets(const cm_pq_params* prm_host, int32_t nv, int32_t nfe,
                                const double* coords, const int32_t* efverts,
                                const int32_t* efopp, const int32_t* efkind, const double* qdata,
                                const int32_t* v2e_ptr, const int32_t* v2e_ent,
                                const int32_t* nrowptr, const int32_t* ncol, double* Kq, double* cq,
                                void* stream);

/* ---- a8: DirichletBC::apply(A), apply(b,x) inside NewtonSolver [ext], triggered by bcs=bc at
 * advection_diffusion_convergence.py:208.  bs = 1 (scalar) or 2 (P-Q).  rows are dof numbers
 * (unique), g the prescribed values; diag_slot[nrows] = slot of the node in its own node row.
 * A[row,:]=0, A[row,row]=1, b[row]=x[row]-g  (vals may be NULL: residual only). */
int cm_apply_dirichlet(int32_t bs, int32_t nrows, const int32_t* rows, const int32_t* diag_slot,
                       const double* g, const double* x, const int32_t* nrowptr, double* vals,
                       double* b, void* stream);

/* ---- a11: linear algebra for the Newton linear solve (replaces PETScLUSolver('mumps'|'umfpack'),
 * advection_diffusion_convergence.py:212, …MixedPrimal_convergence.py:136) ---------------------- */
/* y = A x for the interleaved P-Q layout (bs=2) or scalar node CSR (bs=1) */
int cm_spmv(int32_t bs, int32_t nv, const int32_t* nrowptr, const int32_t* ncol, const double* vals,
            const double* x, double* y, void* stream);

/* Newton linear solve on structured (RectangleMesh-numbered) meshes: right-preconditioned
 * GMRES(restart) on the row-equilibrated Jacobian with the block-lower-triangular field split
 * [[App,0],[Aqp,Aqq]]: p-block = one geometric-multigrid V(1,2) cycle (zebra x-line relaxation,
 * Galerkin coarse operators, dense coarsest solve), q-block = one zebra x-line sweep.
 *   cm_pqs_setup : once per Jacobian (vals = assembled J with Dirichlet rows applied): extracts and
 *                  scales the four stencil blocks, builds the hierarchy and the line factors.
 *   cm_pqs_solve : J x = b (b, x interleaved dof order); returns 0 when the relative residual of
 *                  the row-scaled system reached max(rtol*|b|, atol), 1 when maxit was hit (x holds
 *                  the last iterate), <0 on error.  iters/resid are host outputs.
 * workspace: device memory of cm_pqs_workspace_bytes(nx,ny,restart) bytes owned by the caller;
 * its content carries the setup between the two calls. */
int64_t cm_pqs_workspace_bytes(int32_t nx, int32_t ny, int32_t restart);
int cm_pqs_setup(int32_t nx, int32_t ny, int32_t restart, const int32_t* nrowptr,
                 const int32_t* ncol, const double* vals, void* workspace, int64_t workspace_bytes,
                 void* stream);
int cm_pqs_solve(int32_t nx, int32_t ny, int32_t restart, const double* b, double* x, double rtol,
                 double atol, int32_t maxit, void* workspace, int64_t workspace_bytes,
                 int32_t* iters_host, double* resid_host, void* stream);

/* Solver context (preferred entry points; cm_pqs_setup / cm_pqs_solve above remain and run the
 * first-generation kernels).  cm_pqs_create binds a workspace (and, multi-GPU, a cm_dist describing the strip
 * partition; NULL on one GPU), reads the tuning knobs from the environment ONCE and caches the CUDA graphs
 * of one GMRES iteration, so repeated Newton steps replay them.  Per Newton step: cm_pqs_ctx_setup (new
 * Jacobian values), cm_pqs_ctx_solve.  The context owns no device memory; destroy it before the workspace.
 *   Kernels: fused zebra half sweeps (couplings + tridiagonal solve + multi-GPU boundary push in one launch,
 *   inputs streamed by cp.async.bulk), restricted residual on even rows, correction on odd rows, one cluster
 *   kernel for the coarse levels, device-resident GMRES with one reduction per iteration.
 * cm_pqs_ctx_info: out8 = {levels, first level of the fused coarse cycle, first replicated level (-1: one
 *   GPU), generation of the path in use (1|2), kernels per GMRES iteration graph, cluster size of the coarse
 *   cycle, graphs usable, 13-point q-block present}. */
void* cm_pqs_create(int32_t nx, int32_t ny, int32_t restart, void* workspace, int64_t workspace_bytes,
                    const struct cm_dist* dist_or_null);
int cm_pqs_destroy(void* ctx);
int cm_pqs_ctx_setup(void* ctx, const int32_t* nrowptr, const int32_t* ncol, const double* vals, void* stream);
int cm_pqs_ctx_solve(void* ctx, const double* b, double* x, double rtol, double atol, int32_t maxit,
                     int32_t* iters_host, double* resid_host, void* stream);
int cm_pqs_ctx_info(void* ctx, int32_t* out8);

/* ---- (e) multi GPU: strip-partitioned Newton step, one process per GPU ---------------------------
 * The mesh is split into strips of whole vertex rows (RCB along r1 keeps the r2-lines of the line
 * solver intact).  Every rank holds globally indexed arrays in an IPC-shared workspace
 * (cm_ipc_alloc / cm_ipc_get_handle / cm_ipc_open) and computes only its own rows; boundary rows,
 * the coarse-grid all-gather and the Krylov all-reduces travel over NVLink by peer-to-peer stores +
 * sequence flags (no data-path collective library).  The Newton state u must live at
 * cm_pqs_state_offset() inside the workspace so that cm_dist_exchange_rows can push its halo rows.
 * cm_pqs_setup_dist / cm_pqs_solve_dist mirror cm_pqs_setup / cm_pqs_solve; vals, b and x are
 * globally indexed too (only the owned rows are read / written). */
void* cm_ipc_alloc(int64_t bytes);
int cm_ipc_free(void* p);
int cm_ipc_get_handle(void* p, void* handle64_host);
void* cm_ipc_open(const void* handle64_host);
int cm_ipc_close(void* p);
int64_t cm_pqs_state_offset(int32_t nx, int32_t ny, int32_t restart);
int cm_assemble_pq_cells_structured_rows(const cm_pq_params* prm_host, int32_t nx, int32_t ny,
                                         const double* coords, const int32_t* nrowptr,
                                         const double* u, const double* p_old, const double* load,
                                         double* vals, double* b, int32_t row_begin,
                                         int32_t row_end, void* stream);

/* Stencil form of the P-Q Jacobian inside a solver workspace (cm_pqs_ctx_blocks): four 7-point blocks stored as
 * diagonals [7][nv] (slot order of cm_stencil_*), and the row equilibration sp, sq = 1 / max |row entry|. */
typedef struct cm_pq_blocks {
  double* PP;
  double* PQ;
  double* QP;
  double* QQ;
  double* sp;
  double* sq;
} cm_pq_blocks;

/* cm_assemble_pq_cells_structured_rows that ALSO writes the rows [row_begin, row_end) of the stencil blocks and of
 * the row scalings straight from the accumulators of the sweep (the values are bit-identical to what cm_pqs_ctx_setup
 * would extract from `vals`): the 1.2 ms extraction pass over the 1.9 GB CSR values of a 16 Mi-cell mesh disappears
 * from every Newton step.  Valid when the Jacobian consists of the cell integrals only (7-point pattern); vals must
 * not be NULL.  cm_apply_dirichlet_blocks gives the Dirichlet rows of the blocks the treatment cm_apply_dirichlet
 * gives the CSR rows; cm_pqs_ctx_setup_from_blocks is cm_pqs_ctx_setup without the extraction. */
int cm_assemble_pq_cells_structured_blocks(const cm_pq_params* prm_host, int32_t nx, int32_t ny, const double* coords,
                                           const int32_t* nrowptr, const double* u, const double* p_old,
                                           const double* load, double* vals, double* b, const cm_pq_blocks* blocks,
                                           int32_t row_begin, int32_t row_end, void* stream);
int cm_apply_dirichlet_blocks(int32_t nrows, const int32_t* rows, int32_t nv, const cm_pq_blocks* blocks, void* stream);
int cm_pqs_ctx_blocks(void* ctx, cm_pq_blocks* out);
int cm_pqs_ctx_setup_from_blocks(void* ctx, void* stream);
/* push rows row_begin and row_end-1 of the array at array_offset_bytes (inside the workspace) to the
 * lower / upper neighbour and wait for theirs */
int cm_dist_exchange_rows(cm_dist* D, int64_t array_offset_bytes, int32_t row_doubles,
                          int32_t row_begin, int32_t row_end, void* stream);
/* push rows [row_begin,row_end) to every peer and wait for everybody's */
int cm_dist_allgather_rows(cm_dist* D, int64_t array_offset_bytes, int32_t row_doubles,
                           int32_t row_begin, int32_t row_end, void* stream);
/* in-place sum over ranks of n <= 64 device doubles, fixed rank order */
int cm_dist_allreduce_sum(cm_dist* D, double* vals, int32_t n, void* stream);
int cm_pqs_setup_dist(int32_t nx, int32_t ny, int32_t restart, const int32_t* nrowptr,
                      const int32_t* ncol, const double* vals, void* workspace,
                      int64_t workspace_bytes, cm_dist* D, void* stream);
int cm_pqs_solve_dist(int32_t nx, int32_t ny, int32_t restart, const double* b, double* x,
                      double rtol, double atol, int32_t maxit, void* workspace,
                      int64_t workspace_bytes, cm_dist* D, int32_t* iters_host, double* resid_host,
                      void* stream);

/* Second-generation building blocks (cm_mg.cu), single-GPU forms, exported for tests:
 *   cm_mg_zebra: one zebra half sweep (colour = line parity) in ONE launch, x_is_zero skips the couplings;
 *   cm_mg_resid_restrict: rc = R (b - A x) assuming the residual vanishes on odd fine rows (i.e. right after
 *     a sweep that ended with colour 1); nxf, nyf even;
 *   cm_mg_prolong_odd: x += P e on the odd fine rows only (the even rows are overwritten by the colour-0
 *     half sweep that follows). */
int cm_mg_zebra(int32_t nx, int32_t ny, int32_t colour, const double* A, const double* lf, const double* iu,
                const double* cu, const double* b, double* x, int32_t x_is_zero, void* stream);
/* Long-line half sweep with the seven operator streams (off-line diagonals 0, 1, 5, 6 and the line factors) read
 * from an fp32 pack instead of the fp64 arrays: cm_mg_pack_f32 writes the pack (cm_mg_pack_f32_bytes bytes, 16-byte
 * aligned, once per Jacobian), cm_mg_zebra_f32 is cm_mg_zebra on it (lines of 1024 .. 4300 unknowns).  x, b and all
 * arithmetic stay fp64; the result is the line solve of the fp32-rounded operator (a preconditioner detail). */
int64_t cm_mg_pack_f32_bytes(int32_t nx, int32_t ny);
int cm_mg_pack_f32(int32_t nx, int32_t ny, const double* A, const double* lf, const double* iu, const double* cu,
                   void* pack, void* stream);
int cm_mg_zebra_f32(int32_t nx, int32_t ny, int32_t colour, const double* A, const double* lf, const double* iu,
                    const double* cu, const double* b, double* x, int32_t x_is_zero, const void* pack, void* stream);
/* Half sweeps of the cache-resident middle levels (cm_mgmid.cu; lines of up to 1280 unknowns, whole grid):
 *   mode 0: colour half sweep from a zero guess;  mode 1: with couplings;
 *   mode 2: colour-1 half sweep fused with the restricted residual, bc[(ny/2+1)*(nx/2+1)] = R (b - A x)
 *           (lines of up to 640 unknowns, nx and ny even);
 *   mode 3: colour-0 half sweep that reads x + P ec on the odd rows (prolongation folded in; the odd rows of
 *           x itself are NOT updated: the colour-1 half sweep that must follow overwrites them). */
int cm_mg_zebra_mid(int32_t mode, int32_t nx, int32_t ny, int32_t colour, const double* A, const double* lf,
                    const double* iu, const double* cu, const double* b, double* x, const double* ec, double* bc,
                    void* stream);
/* diagnostics: %globaltimer phase stamps [cta < 64][line < 16][8] of the last cm_mg_zebra launch that ran with
 * the environment variable CM_RING_DBG set (tools/time_ring.py) */
int cm_mg_ring_debug_times(uint64_t* host_out, int32_t n);
/* diagnostics: %globaltimer phase stamps of the last shared-memory resident coarse cycle that ran with CM_COARSE_DBG
 * set (tools/time_smem_cycle.py); returns the number of stamps written (negative: error code) */
int cm_mg_smem_debug_times(uint64_t* host_out, int32_t n);
int cm_mg_resid_restrict(int32_t nxf, int32_t nyf, const double* A, const double* b, const double* x,
                         double* rc, void* stream);
int cm_mg_prolong_odd(int32_t nxc, int32_t nyc, const double* ec, double* xf, void* stream);

/* Building blocks of the structured solver, exported for tests and reuse.  A is a 7-point stencil
 * stored as diagonals A[k*n + node], n = (nx+1)*(ny+1), k: 0:(-1,-1) 1:(0,-1) 2:(-1,0) 3:(0,0)
 * 4:(+1,0) 5:(0,+1) 6:(+1,+1); entries pointing outside the grid must be zero.
 *   cm_stencil_line_factor: LU factors of the tridiagonal x-line blocks (l, 1/u, c/u per node);
 *     *err_flag (device int, zeroed by the caller) is set to 2 on a zero pivot.
 *   cm_stencil_zebra: for every grid line of one colour (line index parity) solve
 *     T x_line = b_line - (A - T) x with the current x on the neighbouring lines (x_is_zero: skip them).
 *     tmp: n doubles of scratch or NULL; with scratch the sweep runs as a streaming right-hand-side
 *     pass + persistent line-solve CTAs that prefetch the next line by cp.async under the scans.
 *   cm_stencil_residual: r = b - A x (b may be NULL). */
int cm_stencil_line_factor(int32_t nx, int32_t ny, const double* A, double* lf, double* iu,
                           double* cu, int32_t* err_flag, void* stream);
int cm_stencil_zebra(int32_t nx, int32_t ny, int32_t colour, const double* A, const double* lf,
                     const double* iu, const double* cu, const double* b, double* x,
                     int32_t x_is_zero, double* tmp, void* stream);
int cm_stencil_residual(int32_t nx, int32_t ny, const double* A, const double* b, const double* x,
                        double* r, void* stream);

/* Generic path (any mesh): GMRES(restart) with point (bs=1) or 2x2 node-block (bs=2) Jacobi on the
 * node-level CSR layouts.  Adequate for the scalar forms (SphericalLaplace_convergence.py:135,
 * SphericalAdvectionDiffusionReaction_convergence.py:104); NOT for the Galerkin P-Q system. */
int64_t cm_krylov_workspace_bytes(int32_t bs, int32_t nv, int32_t restart);
int cm_krylov_solve(int32_t bs, int32_t nv, const int32_t* nrowptr, const int32_t* ncol,
                    const int32_t* diag_slot, const double* vals, const double* b, double* x,
                    double rtol, double atol, int32_t maxit, int32_t restart, void* workspace,
                    int64_t workspace_bytes, int32_t* iters_host, double* resid_host, void* stream);

/* The same operator and preconditioner under right-preconditioned BiCGStab (DOLFIN's 'bicgstab' choice of
 * parameters['newton_solver']['linear_solver']); x0 = 0; one iteration = two SpMVs.  Returns 0 converged,
 * 1 not converged within maxit, < 0 error (breakdown included). */
int64_t cm_bicgstab_workspace_bytes(int32_t bs, int32_t nv);
int cm_bicgstab_solve(int32_t bs, int32_t nv, const int32_t* nrowptr, const int32_t* ncol,
                      const int32_t* diag_slot, const double* vals, const double* b, double* x,
                      double rtol, double atol, int32_t maxit, void* workspace, int64_t workspace_bytes,
                      int32_t* iters_host, double* resid_host, void* stream);

/* ---- a11, direct path: banded LU with partial pivoting (block-tridiagonal organisation) ----------
 * Replaces PETScLUSolver('mumps'|'umfpack'|default LU): advection_diffusion_convergence.py:212,
 * advection_diffusion_mixed.py:223, ...MixedPrimal_convergence.py:136, SphericalLaplace_convergence.py:135
 * -- the exact (to rounding) solve the reference relies on for the unstabilised Galerkin q-equation
 * (zero q-q diagonal, advection_diffusion_convergence.py:203) and on unstructured meshes, where the
 * multigrid-preconditioned Krylov path does not apply.
 *   perm[nv]  : new position of every node; iperm its inverse.  Coupled nodes must end up at most nbn
 *               positions apart (RectangleMesh numbering: nbn = nx+2, or ny+2 after transposing).
 *   bs        : 1 scalar node CSR, 2 the interleaved P-Q value layout described at the top.
 *   workspace : cm_blu_workspace_bytes(bs,nv,nbn) bytes = 32*bs^2*nbn^2 per block of nbn nodes
 *               (L, U with fill and the pivot permutations); carries the factors to cm_blu_solve.
 *   ctx       : host object owning the cuSOLVER/cuBLAS handles and the getrf scratch (dense panel LU,
 *               triangular solves and Schur updates are library calls; scatter / pivot / gather /
 *               permutation kernels are this library's).
 * cm_blu_factor synchronises the stream once at the end and reports info_host = 0, or 1 + the
 * renumbered dof of an exactly zero pivot (returns CM_E_BREAKDOWN). */
int64_t cm_blu_workspace_bytes(int32_t bs, int32_t nv, int32_t nbn);
void* cm_blu_create(void);
int cm_blu_destroy(void* ctx);
int cm_blu_factor(void* ctx, int32_t bs, int32_t nv, int32_t nbn, const int32_t* perm,
                  const int32_t* iperm, const int32_t* nrowptr, const int32_t* ncol,
                  const double* vals, void* workspace, int64_t workspace_bytes, int32_t* info_host,
                  void* stream);
int cm_blu_solve(void* ctx, int32_t bs, int32_t nv, int32_t nbn, const int32_t* perm, const double* b,
                 double* x, void* workspace, int64_t workspace_bytes, void* stream);

/* ---- a12: scalar functionals and the flux post-processing of the drivers --------------------------
 * rule_host[3*nq] = (xi, eta, weight) of the triangle rule (weights sum to 1/2; FIAT default rules).
 * All reductions use a fixed grid and summation order (bit-reproducible).
 *   cm_functional_error_norms: work[0] = sum a (u_ex - u_h)^2, work[1] = sum a |grad u_ex - grad u_h|^2
 *     (x derivative only when x_derivative_only), a = w |det| [* (4 pi x y)^2 when weighted]; take the
 *     square roots on the host.  Replaces assemble((u_ex-u_h)**2*dx) etc. at
 *     SphericalAdvectionDiffusionReaction_convergence.py:117-118 and ...MixedPrimal_convergence.py:148-150.
 *     uh: nodal values with stride uh_stride (2 for a component of the interleaved P-Q vector);
 *     exact_q[ncells*nq], grad_q[ncells*nq*2]: exact data at the quadrature points;
 *     work: cm_functional_workspace_doubles() device doubles.
 *   cm_project_flux_dg0: flux[ncells*2] = cell means of (D2*(dp/dx + x^2 p^2/q), D1*dp/dy) = the
 *     project(dMatrix*grad(p_h) + D2*r2*r2*p_h/q_h*p_h*r2_vec, VectorFunctionSpace(mesh,"DG",0)) of
 *     convergence/advection_diffusion_convergence.py:222-223.
 *   cm_functional_boundary_flux: out3 = (r2 part, r1 part, total) of
 *     assemble(dot(flux,n)*4*pi*r1**2*ds(marker)) (advection_diffusion_mixed.py:261-263) over the listed
 *     exterior facets: efverts[nfe*2], efopp[nfe] (vertex of the attached cell opposite the facet),
 *     efcell[nfe] (the attached cell). */
int64_t cm_functional_workspace_doubles(void);
int cm_functional_error_norms(int32_t nq, const double* rule_host, int32_t weighted,
                              int32_t x_derivative_only, int32_t ncells, const int32_t* cells,
                              const double* coords, const double* uh, int32_t uh_stride,
                              const double* exact_q, const double* grad_q, double* work, void* stream);
int cm_project_flux_dg0(int32_t nq, const double* rule_host, double D1, double D2, int32_t ncells,
                        const int32_t* cells, const double* coords, const double* u, double* flux,
                        void* stream);
int cm_functional_boundary_flux(int32_t nfe, const int32_t* efverts, const int32_t* efopp,
                                const int32_t* efcell, const double* coords, const double* flux,
                                double* out3, void* stream);

/* ---- section 8f row 1: lowest-order mixed-primal system RT1 x DG0 x CG1 (deg = 0) ------------------
 * convergence/test_for_paper_convergence.py:68,108-167; its recorded k=0 table (:200-207) is the
 * known-answer test.  Global dofs [edges | cells | vertices]; per cell 7 local dofs (3 edge fluxes on the
 * facets opposite local vertices 0,1,2, the cell value of p, 3 vertex values of q).
 *   cm_mixed_m0_cells : element residual Fe[7][ncells] and Jacobian Je[49][ncells] (entry-major; Je may be
 *     NULL) of the cell integrals (:137-141,150-151).  cell_edges[ncells*3], sign[ncells*3] (+-1: outward
 *     normal of the cell vs the global edge normal), f2q/f3q[ncells*nq] forcing at the quadrature points.
 *   cm_csr_gather     : out[s] (+)= sum of src[idx[k]], k in [ptr[s],ptr[s+1]) in list order -- the
 *     deterministic replacement of GenericTensor::add_local [ext] through a precomputed cell->nnz map
 *     (lists sorted by cell = DOLFIN's serial insertion order); works for matrices and vectors.
 *   cm_csr_add_mapped : dst[map[k]] += src[k] (injective map; adds the q-q facet operator of
 *     cm_assemble_c0ip_facets / cm_assemble_exterior_facets into the mixed CSR).
 *   cm_mixed_m0_errors: work[0..2] = squares of e_div(s), e_0(p), e_1(q) (:175-177);
 *     exact_q[ncells*nq*6] = (s_x, s_y, div_rad s, p, q, dq/dx) at the quadrature points;
 *     work: cm_mixed_workspace_doubles() device doubles. */
int cm_mixed_m0_cells(int32_t nq, const double* rule_host, double D1, double D2, double geps,
                      int32_t ncells, const int32_t* cells, const double* coords,
                      const int32_t* cell_edges, const double* sign, int32_t nedges, const double* u,
                      const double* f2q, const double* f3q, double* Je, double* Fe, void* stream);
int cm_csr_gather(int32_t nslots, const int32_t* ptr, const int32_t* idx, const double* src, double* out,
                  int32_t accumulate, void* stream);
int cm_csr_add_mapped(int32_t n, const int32_t* map, const double* src, double* dst, void* stream);
int64_t cm_mixed_workspace_doubles(void);
int cm_mixed_m0_errors(int32_t nq, const double* rule_host, int32_t ncells, const int32_t* cells,
                       const double* coords, const int32_t* cell_edges, const double* sign,
                       int32_t nedges, const double* u, const double* exact_q, double* work, void* stream);

/* k = 1 (RT2 x DG1 x CG2, recorded table test_for_paper_convergence.py:211-217): 17 local dofs per cell
 * [8 fluxes: (lambda_j N_i, lambda_k N_i) for the facets i = 0,1,2 with end points j < k, then lambda_1 N_1,
 * lambda_2 N_2 | 3 DG1 values | 6 CG2 values: vertices, then the facets' mid points]; global dofs
 * [2 per edge | 2 per cell | 3 per cell | vertices | edges].  Je[289][ncells], Fe[17][ncells] entry-major;
 * assembled by cm_csr_gather like the k = 0 tensors.
 * G = gcoef x^2 p^2/(q + geps).  qmode 0: the q-equation integrated by parts, its facet terms coming from
 * cm_cg2_*_facets: ((q_react - div_rad(r2vec)) q + q_x2p r2^2 p - f3) w W - (r2vec.grad w) q W with r2vec = (1,0):
 * (0,1) = test_for_paper_convergence.py:140-141 and ...MixedPrimal_convergence_C0IP.py:144-145, (1,0) = the q-q block
 * of C0IP_pure_advection_spherical_convergence.py:109-110; qmode 2: the same with r2vec = (-1,0)
 * (..._C0IP_otherBCs.py:58,131-132); qmode 1: the q-equation
 * (q_react q + r2vec.grad q + r2^2 (q_x2p p + q_x2q q) - f3)(w + supg_stab hK/2 r2vec.grad w) W:
 * (1,1,0,0) = ...MixedPrimal_convergence.py:123 (BASELINE config 2's own script, Galerkin),
 * (1,0,1,1) = ...MixedPrimal_convergence_SUPG.py:118-129 (hK = CellDiameter).
 * cm_mixed_m1_errors: exact_q[ncells*nq*7] = (s_x, s_y, div_rad s, p, q, dq/dx, dq/dy); full_grad selects the q
 * error with the whole gradient (...MixedPrimal_convergence.py:148) instead of d/dx only (test_for_paper:175). */
int cm_mixed_m1_cells(int32_t nq, const double* rule_host, double D1, double D2, double geps, double gcoef,
                      int32_t qmode, double q_react, double q_x2p, double q_x2q, double supg_stab,
                      int32_t ncells, int32_t nedges, int32_t nverts, const int32_t* cells, const double* coords,
                      const int32_t* cell_edges, const double* sign, const double* u, const double* f2q,
                      const double* f3q, double* Je, double* Fe, void* stream);
int cm_mixed_m1_errors(int32_t nq, const double* rule_host, int32_t full_grad, int32_t ncells, int32_t nedges,
                       int32_t nverts, const int32_t* cells, const double* coords, const int32_t* cell_edges,
                       const double* sign, const double* u, const double* exact_q, double* work, void* stream);

/* CG2 facet operators of the k = 1 q-equation (state independent, once per mesh; :142-145,148-149,152-153).
 * rule_host[2*nq] = (s, w) Gauss-Legendre on [0,1].  fverts[nf*2] ascending vertex ids.
 *   exterior: fcell[nfe] attached cell, kind bits 1: (r2vec.n) q w W ds, 2: (r2vec.n) q_ex w W ds,
 *     4: natural p data on the facet's two flux dofs, 8: r2vec = (-1,0) instead of (1,0); the boundary penalty stab2 (hK neg(r2vec.n)^2 + 8/hK)
 *     (q - q_ex) w ds applies to every listed facet.  qex / pex[nfe*nq]: q_ex, p_ex at the facet points.
 *     Outputs entry-major: Jx[36][nfe] (6 x 6 on the attached cell's CG2 dofs), cx[6][nfe], nat[2][nfe].
 *   interior: fcell[nfi*2] = (cell0 = '+', cell1); Ji[144][nfi], 12 x 12 on (cell0's 6, cell1's 6) CG2 dofs:
 *     gamma avg(hK)^2 (jump(grad q).n+)(jump(grad w).n+) W dS; avg(hK) = hbar[nfi] (e.g. the mean CellDiameter of
 *     the two cells, ...MixedPrimal_convergence_C0IP.py:98,147) or, with hbar = NULL, the facet length (FacetArea,
 *     test_for_paper_convergence.py:97). */
int cm_cg2_exterior_facets(int32_t nq, const double* rule_host, double stab2, int32_t nfe,
                           const int32_t* fverts, const int32_t* fcell, const int32_t* kind,
                           const int32_t* cells, const double* coords, const double* qex, const double* pex,
                           double* Jx, double* cx, double* nat, void* stream);
int cm_cg2_interior_facets(int32_t nq, const double* rule_host, double gamma, int32_t nfi,
                           const int32_t* fverts, const int32_t* fcell, const int32_t* cells,
                           const double* coords, const double* hbar, double* Ji, void* stream);

/* Production mixed formulation DG2 x RT3 x CG3 (advection_diffusion_mixed.py:147-150,196-216; deg = 2): element
 * residual Fe[31][ncells] and Jacobian Je[961][ncells] (entry-major; Je may be NULL) of the cell integrals
 *   [(p - p_old)/dt] v W + div_rad(D s) v W + (s + Gstar(p,q) e_x).tau W - p div_rad(tau) W + (dq/dx + x^2 p) w W
 * with the production Gstar conditional (:17-23).  Local dofs [6 DG2 | 15 RT3: per facet i (end points j < k)
 * lambda_j^2 N_i, lambda_j lambda_k N_i, lambda_k^2 N_i, then lambda_i lambda_m N_i for i = 1,2, m = 0,1,2 |
 * 10 CG3: vertices, two nodes per facet, bubble]; global dofs [6 per cell | 3 per edge, 6 per cell | vertices,
 * 2 per edge, 1 per cell].  p_old[6*ncells] (DG2 coefficients) is read when inv_dt != 0.  rule_host: up to 40
 * points.  Assembled by cm_csr_gather, solved by cm_blu_*. */
int cm_mixed_p2_cells(int32_t nq, const double* rule_host, double D1, double D2, double concentration,
                      double inv_dt, int32_t ncells, int32_t nedges, int32_t nverts, const int32_t* cells,
                      const double* coords, const int32_t* cell_edges, const double* sign, const double* u,
                      const double* p_old, double* Je, double* Fe, void* stream);

/* ---- a9: sparsity pattern of the P1 space (DOLFIN SparsityPatternBuilder [ext], triggered by
 * FunctionSpace(mesh, MixedElement([P0,P1])) at advection_diffusion_convergence.py:149-151 and the first assemble).
 * Node-level CSR with sorted columns: union over cells of vertex x vertex, plus the nextra vertex couples of
 * extra_pairs[nextra*2] inserted symmetrically (the two vertices opposite every interior facet: the C0IP dS
 * pattern).  col needs capacity for 9*ncells + 2*nextra entries at most; the number of entries is returned in
 * nnz_host (the stream is synchronised once).  The blocked P-Q pattern follows from the node pattern (see the
 * value layout at the top of this header). */
int64_t cm_p1_sparsity_workspace_bytes(int32_t nv, int32_t ncells, int32_t nextra);
int cm_build_p1_sparsity(int32_t nv, int32_t ncells, const int32_t* cells, int32_t nextra,
                         const int32_t* extra_pairs, int32_t* rowptr, int32_t* col, int64_t col_capacity,
                         int64_t* nnz_host, void* workspace, int64_t workspace_bytes, void* stream);

/* Fused Krylov vector kernels (the building blocks of the GMRES orthogonalisation, exported for reuse and tests):
 *   cm_dot_batch : out[k] = <V_k, w>, k < nvec, V_k = V + k*vstride -- every group of up to 8 basis vectors in
 *                  ONE pass over w (classical Gram-Schmidt reads w once per group, not once per vector);
 *   cm_axpy_batch: w += sign * sum_k h[k] V_k (h on the device) and, when norm2_out != NULL, *norm2_out = |w|^2
 *                  from the same pass.
 * scratch: cm_vec_scratch_doubles() device doubles.  Fixed grid and summation order: bit-reproducible. */
int64_t cm_vec_scratch_doubles(void);
int cm_dot_batch(int64_t n, const double* V, int64_t vstride, int32_t nvec, const double* w, double* out,
                 double* scratch, void* stream);
int cm_axpy_batch(int64_t n, const double* V, int64_t vstride, int32_t nvec, const double* h, double sign,
                  double* w, double* norm2_out, double* scratch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CLOSESTMOL_B200_H */
