// Internal (C++) launchers behind the C ABI.
#pragma once
#include "cm_common.cuh"

namespace cm {
struct Halo;
}

namespace cm {

int assemble_pq_rows(const cm_pq_params& P, int nv, const int* cells, const double* coords,
                     const int* v2c_ptr, const int* v2c_ent, const int* v2c_slot,
                     const int* nrowptr, const int* ncol, const double* u, const double* p_old,
                     const double* load, const double* Kq, const double* cq, double* vals,
                     double* b, int accumulate, int maxdeg, cudaStream_t st);
int assemble_scalar_rows(const cm_scalar_params& P, int nv, const int* cells, const double* coords,
                         const int* v2c_ptr, const int* v2c_ent, const int* v2c_slot,
                         const int* nrowptr, const double* u, const double* load, double* vals,
                         double* b, int maxdeg, cudaStream_t st);
int assemble_load(int ncomp, int weighted, double supg_stab, int degree, int nv, const int* cells,
                  const double* coords, const int* v2c_ptr, const int* v2c_ent, const double* fq,
                  double* load, cudaStream_t st);
int assemble_c0ip(const cm_pq_params& P, int nv, const int* cells, const double* coords,
                  const int* fverts, const int* fcell, const int* v2f_ptr, const int* v2f_ent,
                  const int* nrowptr, const int* ncol, double* Kq, cudaStream_t st);
int assemble_ext(const cm_pq_params& P, int nv, const double* coords, const int* efverts,
                 const int* efopp, const int* efkind, const double* qdata, const int* v2e_ptr,
                 const int* v2e_ent, const int* nrowptr, const int* ncol, double* Kq, double* cq,
                 cudaStream_t st);
int apply_dirichlet(int bs, int nrows, const int* rows, const int* diag_slot, const double* g,
                    const double* x, const int* nrowptr, double* vals, double* b, cudaStream_t st);

}  // namespace cm

namespace cm {
// y = A*x, or y = z + alpha*A*x when z != nullptr
int spmv(int bs, int nv, const int* nrowptr, const int* ncol, const double* vals, const double* x,
         double* y, double alpha, const double* z, cudaStream_t st);
}  // namespace cm

#include <functional>
namespace cm {
using LinOp = std::function<int(const double*, double*)>;

// cm_vec.cu
int multi_dot(long long n, const double* Vbase, long long vstride, int nvec, const double* w,
              double* out, double* scratch, cudaStream_t st);
int multi_axpy(long long n, const double* Vbase, long long vstride, int nvec, const double* h,
               double sign, double* w, double* norm2_out, double* scratch, cudaStream_t st);
size_t vec_scratch_doubles();
int vec_scale(long long n, const double* x, const double* a_dev, double a_host, int invert,
              double* y, cudaStream_t st);

struct VecSlice {
  long long off0, off1, len;  // indices [off0, off0+len) U [off1, off1+len)
};
int multi_dot_slice(VecSlice sl, const double* Vbase, long long vstride, int nvec, const double* w,
                    double* out, double* scratch, cudaStream_t st);
int multi_axpy_slice(VecSlice sl, const double* Vbase, long long vstride, int nvec, const double* h,
                     double sign, double* w, double* norm2_out, double* scratch, cudaStream_t st);
int vec_scale_slice(VecSlice sl, const double* x, double a, int invert, double* y, cudaStream_t st);

// cm_dist.cu
size_t dist_ctl_doubles();
int dist_halo(cm_dist* D, const unsigned long long* offs, const int* ns, int nseg, cudaStream_t st);
int dist_halo2(cm_dist* D, const unsigned long long* offs_dn, const int* ns_dn, int nseg_dn,
               const unsigned long long* offs_up, const int* ns_up, int nseg_up, cudaStream_t st);
int dist_allgather(cm_dist* D, const unsigned long long* offs, const int* ns, int nseg,
                   cudaStream_t st);
int dist_allreduce(cm_dist* D, double* vals, int n, cudaStream_t st);

// cm_gmres_dev.cu: device-resident GMRES state (all pointers into the caller's workspace)
constexpr int kGmresStatus = 16;
enum GmresStatusSlot { kStJ = 0, kStIts = 1, kStConv = 2, kStResid = 3, kStTarget = 4, kStBnorm = 5,
                       kStFlag = 6, kStK = 7, kStMaxit = 8 };
struct GmresDev {
  long long n, ne;
  int m;
  double *V, *w, *z, *Z;
  double *dots, *coef, *vscale, *nrm, *H, *cs, *sn, *g, *y, *st;
  double* partial;
  unsigned* tickets;
};
size_t gmres_dev_doubles(long long n, int m);
void gmres_dev_layout(GmresDev& G, long long n, int m, double* work);
void gmres_dev_init();   // occupancy-derived launch geometry, once per process (not inside a stream capture)
int gmres_dev_begin(const GmresDev& G, int first, cudaStream_t st, int maxit = 1 << 30);
int gmres_dev_dot(const GmresDev& G, const VecSlice* sl, int jhost, cudaStream_t st);
// ru != nullptr: also writes ru = diag(sp | sq)^-1 V_{j+1} (unnormalised), the next preconditioner input
int gmres_dev_update(const GmresDev& G, const VecSlice* sl, int jhost, cudaStream_t st, const double* sp = nullptr,
                     const double* sq = nullptr, double* ru = nullptr);
// cond_handle != 0: a cudaGraphConditionalHandle; the kernel sets it to "another iteration is needed" (the GMRES
// cycle runs as a WHILE node of a CUDA graph, cm_pqsolver.cu)
int gmres_dev_givens(const GmresDev& G, double rtol, double atol, cudaStream_t st, unsigned long long cond_handle = 0);
// adds a one-thread kernel node to `graph` that evaluates the same condition from the current status block
int gmres_dev_add_cond_node(cudaGraph_t graph, cudaGraphNode_t* node, const GmresDev& G, unsigned long long cond_handle);
int gmres_dev_solve_y(const GmresDev& G, cudaStream_t st);
int gmres_dev_combine(const GmresDev& G, const VecSlice* sl, double* out, int khost, cudaStream_t st);
int gmres_dev_axpby(const GmresDev& G, const VecSlice* sl, double alpha, const double* x, double beta, double* y,
                    cudaStream_t st);

// cm_gmres.cu
size_t gmres_workspace_doubles(long long n, int restart);
int gmres_solve(long long n, const LinOp& A, const LinOp& Minv, const double* b, double* x,
                double rtol, double atol, int maxit, int restart, int reorth, double* work,
                int* iters_out, double* resid_out, double* bnorm_out, cudaStream_t st,
                const VecSlice* sl = nullptr, cm_dist* D = nullptr);

size_t bicgstab_workspace_doubles(long long n);
int bicgstab_solve(long long n, const LinOp& A, const LinOp& Minv, const double* b, double* x, double rtol,
                   double atol, int maxit, double* work, int* iters_out, double* resid_out,
                   double* bnorm_out, cudaStream_t st);

// cm_stencil.cu
int st_extract_blocks(int nx, int ny, const int* nrowptr, const int* ncol, const double* vals,
                      double* PP, double* PQ, double* QP, double* QQ, double* sp, double* sq,
                      int* err, cudaStream_t st, int r0 = 0, int r1 = -1, double* QQx = nullptr,
                      int* has_extra = nullptr, int* wide_host = nullptr);
int st_lump_extras(int nx, int ny, const double* QQ, const double* QQx, double* QQl, cudaStream_t st);
int st_split_scale(int nv, const double* b, const double* sp, const double* sq, double* out,
                   cudaStream_t st, int vb = 0, int ve = -1);
int st_interleave(int nv, const double* xs, double* x, cudaStream_t st, int vb = 0, int ve = -1);
int st_pq_spmv(int nx, int ny, const double* PP, const double* PQ, const double* QP,
               const double* QQ, const double* sp, const double* sq, const double* x, double* y,
               cudaStream_t st, int r0 = 0, int r1 = -1, const double* QQx = nullptr,
               const int* has_extra = nullptr, const struct Halo* H = nullptr, double* zcopy = nullptr,
               long long zstride = 0, const double* jptr = nullptr, int halo_channels = 1,
               const double* xscale = nullptr);   // xscale: y and zcopy are multiplied by xscale[*jptr]
int st_unscale_dev(int nx, int ny, const double* rbase, long long stride, const double* jptr, const double* vscale,
                   const double* sp, const double* sq, double* out, cudaStream_t st, int r0 = 0, int r1 = -1);
int st_unscale(int nx, int ny, const double* r, const double* sp, const double* sq, double* out,
               cudaStream_t st, int r0 = 0, int r1 = -1);
int st_residual(int nx, int ny, const double* A, const double* b, const double* x, double* r,
                cudaStream_t st, int r0 = 0, int r1 = -1);
int st_restrict(int nxc, int nyc, const double* rf, double* rc, cudaStream_t st, int R0 = 0, int R1 = -1);
int st_prolong_add(int nxc, int nyc, const double* ec, double* xf, cudaStream_t st, int r0 = 0, int r1 = -1);
int st_rap(int nxc, int nyc, const double* Af, double* Ac, cudaStream_t st, int R0 = 0, int R1 = -1);
int st_init_attributes();
int st_line_factor(int nx, int ny, const double* A, double* lf, double* iu, double* cu, int* err,
                   cudaStream_t st, int r0 = 0, int r1 = -1);
int st_zebra(int nx, int ny, int colour, const double* A, const double* lf, const double* iu,
             const double* cu, const double* b, double* x, int x_zero, cudaStream_t st, int r0 = 0,
             int r1 = -1, double* tmp = nullptr, const double* Ax = nullptr,
             const int* has_extra = nullptr, bool prof = false);
int st_coarse_invert(int nx, int ny, const double* A, double* M, double* Ainv, int* err,
                     cudaStream_t st);
int st_coarse_apply(int n, const double* Ainv, const double* b, double* x, cudaStream_t st);

// cm_sweep.cu
int assemble_pq_sweep(const cm_pq_params& P, int nx, int ny, const double* coords,
                      const int* nrowptr, const double* u, const double* p_old, const double* load,
                      double* vals, double* b, cudaStream_t st, int row_begin = 0, int row_end = -1,
                      const cm_pq_blocks* blocks = nullptr);   // blocks: also write the solver's stencil form

// cm_pqsolver.cu
size_t pqs_workspace_bytes(int nx, int ny, int restart);
int pqs_setup(int nx, int ny, int restart, const int* nrowptr, const int* ncol, const double* vals,
              void* ws, size_t ws_bytes, cudaStream_t st, bool blocks_ready = false, bool no_check = false);
int pqs_solve(int nx, int ny, int restart, const double* b, double* x, double rtol, double atol,
              int maxit, void* ws, size_t ws_bytes, int* iters, double* resid, cudaStream_t st);
int pqs_setup_dist(int nx, int ny, int restart, const int* nrowptr, const int* ncol,
                   const double* vals, void* ws, size_t ws_bytes, cm_dist* D, cudaStream_t st,
                   int Ld_forced = -1, bool blocks_ready = false, bool no_check = false);
void* pqs_create(int nx, int ny, int restart, void* ws, size_t ws_bytes, const cm_dist* D);
int pqs_destroy(void* ctx);
int pqs_ctx_setup(void* ctx, const int* nrowptr, const int* ncol, const double* vals, cudaStream_t st,
                  bool blocks_ready = false);   // blocks_ready: the stencil blocks were written by the assembly sweep
int pqs_ctx_blocks(void* ctx, cm_pq_blocks* out);
int apply_dirichlet_blocks(int nrows, const int* rows, int nv, const cm_pq_blocks& blk, cudaStream_t st);
int pqs_ctx_solve(void* ctx, const double* b, double* x, double rtol, double atol, int maxit, int* iters,
                  double* resid, cudaStream_t st);
int pqs_ctx_info(void* ctx, int* out8);
int pqs_solve_dist(int nx, int ny, int restart, const double* b, double* x, double rtol, double atol,
                   int maxit, void* ws, size_t ws_bytes, cm_dist* D, int* iters, double* resid,
                   cudaStream_t st);
size_t pqs_state_offset_bytes(int nx, int ny, int restart);
size_t bicgstab_workspace_bytes(int bs, int nv);
int bicgstab_csr_solve(int bs, int nv, const int* nrowptr, const int* ncol, const int* diag_slot,
                       const double* vals, const double* b, double* x, double rtol, double atol, int maxit,
                       void* ws, size_t ws_bytes, int* iters, double* resid, cudaStream_t st);
size_t krylov_workspace_bytes(int bs, int nv, int restart);
int krylov_solve(int bs, int nv, const int* nrowptr, const int* ncol, const int* diag_slot,
                 const double* vals, const double* b, double* x, double rtol, double atol, int maxit,
                 int restart, void* ws, size_t ws_bytes, int* iters, double* resid, cudaStream_t st);

// cm_blu.cu
size_t blu_workspace_bytes(int bs, int nv, int nbn);
void* blu_create();
int blu_destroy(void* ctx);
int blu_factor(void* ctx, int bs, int nv, int nbn, const int* perm, const int* iperm,
               const int* nrowptr, const int* ncol, const double* vals, void* ws, size_t ws_bytes,
               int* info_host, cudaStream_t st);
int blu_solve(void* ctx, int bs, int nv, int nbn, const int* perm, const double* b, double* x, void* ws,
              size_t ws_bytes, cudaStream_t st);

// cm_functional.cu
size_t functional_workspace_doubles();
int functional_error_norms(int nq, const double* rule_host, int weighted, int xonly, int ncells,
                           const int* cells, const double* coords, const double* uh, int stride,
                           const double* exq, const double* grq, double* work, cudaStream_t st);
int project_flux_dg0(int nq, const double* rule_host, double D1, double D2, int ncells, const int* cells,
                     const double* coords, const double* u, double* flux, cudaStream_t st);
int functional_boundary_flux(int nfe, const int* efverts, const int* efopp, const int* efcell,
                             const double* coords, const double* flux, double* out3, cudaStream_t st);

// cm_mixed.cu
int mixed_m0_cells(int nq, const double* rule_host, double D1, double D2, double geps, int ncells,
                   const int* cells, const double* coords, const int* cell_edges, const double* sign,
                   int nE, const double* u, const double* f2q, const double* f3q, double* Je, double* Fe,
                   cudaStream_t st);
int csr_gather(int nslots, const int* ptr, const int* idx, const double* src, double* out, int accumulate,
               cudaStream_t st);
int add_mapped(int n, const int* map, const double* src, double* dst, cudaStream_t st);
size_t mixed_workspace_doubles();
int mixed_m0_errors(int nq, const double* rule_host, int ncells, const int* cells, const double* coords,
                    const int* cell_edges, const double* sign, int nE, const double* u, const double* exq,
                    double* work, cudaStream_t st);
int mixed_m1_cells(int nq, const double* rule_host, double D1, double D2, double geps, double gcoef, int qmode,
                   double q_react, double q_x2p, double q_x2q, double supg_stab, int ncells, int nE,
                   int nV, const int* cells, const double* coords, const int* cell_edges, const double* sign,
                   const double* u, const double* f2q, const double* f3q, double* Je, double* Fe,
                   cudaStream_t st);
int mixed_m1_errors(int nq, const double* rule_host, int full_grad, int ncells, int nE, int nV, const int* cells,
                    const double* coords, const int* cell_edges, const double* sign, const double* u,
                    const double* exq, double* work, cudaStream_t st);
int cg2_exterior_facets(int nq, const double* rule_host, double stab2, int nfe, const int* fverts,
                        const int* fcell, const int* kind, const int* cells, const double* coords,
                        const double* qex, const double* pex, double* Jx, double* cx, double* nat,
                        cudaStream_t st);
int cg2_interior_facets(int nq, const double* rule_host, double gamma, int nfi, const int* fverts,
                        const int* fcell, const int* cells, const double* coords, const double* hbar, double* Ji,
                        cudaStream_t st);
int mixed_p2_cells(int nq, const double* rule_host, double D1, double D2, double conc, double inv_dt, int ncells,
                   int nE, int nV, const int* cells, const double* coords, const int* cell_edges,
                   const double* sign, const double* u, const double* p_old, double* Je, double* Fe,
                   cudaStream_t st);
// cm_sparsity.cu
size_t p1_sparsity_workspace_bytes(int nv, int ncells, int nextra);
int build_p1_sparsity(int nv, int ncells, const int* cells, int nextra, const int* extra, int* rowptr, int* col,
                      long long col_capacity, long long* nnz_host, void* ws, size_t ws_bytes, cudaStream_t st);
}  // namespace cm
