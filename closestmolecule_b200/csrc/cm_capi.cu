// extern "C" entry points declared in include/closestmol_b200.h
#include <stdarg.h>

#include "cm_common.cuh"
#include "cm_decl.cuh"
#include "cm_halo.cuh"
#include "cm_mg.cuh"

namespace cm {
static thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};
double g_alg_bytes = 0.0;
const char* last_error_text() { return g_err; }
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace cm

#include <vector>
namespace cm {
struct ProfRec {
  cudaEvent_t a, b;
  int cls;
  double bytes;
};
static bool g_prof_on = false;
bool g_prof_is_on() { return g_prof_on; }
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_prof_pool;
static cudaEvent_t prof_event() {
  cudaEvent_t e;
  if (!g_prof_pool.empty()) {
    e = g_prof_pool.back();
    g_prof_pool.pop_back();
    return e;
  }
  cudaEventCreate(&e);
  return e;
}
void prof_begin(int cls, double bytes, cudaStream_t st) {
  if (!g_prof_on) return;
  ProfRec r{prof_event(), prof_event(), cls, bytes};
  cudaEventRecord(r.a, st);
  g_prof.push_back(r);
}
void prof_end(int cls, cudaStream_t st) {
  if (!g_prof_on || g_prof.empty() || g_prof.back().cls != cls) return;
  cudaEventRecord(g_prof.back().b, st);
}
}  // namespace cm

using namespace cm;
namespace cm { int st_zebra_debug_times(unsigned long long*, int); }
int cm_zebra_debug_times_impl(unsigned long long* o, int n) { return cm::st_zebra_debug_times(o, n); }

extern "C" {

int cm_version(void) { return 100; }

const char* cm_last_error(void) { return g_err; }

int64_t cm_launch_count(void) { return (int64_t)g_launches.load(); }

double cm_algorithmic_bytes(void) { return g_alg_bytes; }

int cm_device_available(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n > 0 ? 1 : 0;
}

int cm_assemble_pq_cells(const cm_pq_params* prm, int32_t nv, int32_t ncells, const int32_t* cells,
                         const double* coords, const int32_t* v2c_ptr, const int32_t* v2c_ent,
                         const int32_t* v2c_slot, const int32_t* nrowptr, const int32_t* ncol,
                         int32_t maxdeg, const double* u, const double* p_old, const double* load,
                         const double* Kq, const double* cq, double* vals, double* b,
                         int32_t accumulate, void* stream) {
  CM_CHECK_ARG(prm && cells && coords && v2c_ptr && v2c_ent && v2c_slot && nrowptr && u && b,
               "null pointer");
  CM_CHECK_ARG(nv > 0 && ncells > 0, "empty mesh");
  CM_CHECK_ARG(maxdeg > 0 && maxdeg <= 48, "row degree outside 1..48");
  CM_CHECK_ARG(prm->inv_dt == 0.0 || p_old != nullptr, "transient form needs p_old");
  CM_CHECK_ARG(Kq == nullptr || ncol != nullptr, "facet operator needs ncol");
  return assemble_pq_rows(*prm, nv, cells, coords, v2c_ptr, v2c_ent, v2c_slot, nrowptr, ncol, u,
                          p_old, load, Kq, cq, vals, b, accumulate, maxdeg, (cudaStream_t)stream);
}

int cm_assemble_pq_cells_structured(const cm_pq_params* prm, int32_t nx, int32_t ny,
                                    const double* coords, const int32_t* nrowptr, const double* u,
                                    const double* p_old, const double* load, double* vals,
                                    double* b, void* stream) {
  CM_CHECK_ARG(prm && coords && nrowptr && u && b, "null pointer");
  CM_CHECK_ARG(nx >= 1 && ny >= 1, "empty mesh");
  CM_CHECK_ARG(prm->inv_dt == 0.0 || p_old != nullptr, "transient form needs p_old");
  return assemble_pq_sweep(*prm, nx, ny, coords, nrowptr, u, p_old, load, vals, b,
                           (cudaStream_t)stream);
}

int cm_assemble_scalar_cells(const cm_scalar_params* prm, int32_t nv, int32_t ncells,
                             const int32_t* cells, const double* coords, const int32_t* v2c_ptr,
                             const int32_t* v2c_ent, const int32_t* v2c_slot,
                             const int32_t* nrowptr, int32_t maxdeg, const double* u,
                             const double* load, double* vals, double* b, void* stream) {
  CM_CHECK_ARG(prm && cells && coords && v2c_ptr && v2c_ent && v2c_slot && nrowptr && u && b,
               "null pointer");
  CM_CHECK_ARG(nv > 0 && ncells > 0, "empty mesh");
  CM_CHECK_ARG(maxdeg > 0 && maxdeg <= 192, "row degree outside 1..192");
  return assemble_scalar_rows(*prm, nv, cells, coords, v2c_ptr, v2c_ent, v2c_slot, nrowptr, u, load,
                              vals, b, maxdeg, (cudaStream_t)stream);
}

int cm_assemble_load(int32_t ncomp, int32_t weighted, double supg_stab, int32_t degree, int32_t nv,
                     int32_t ncells, const int32_t* cells, const double* coords,
                     const int32_t* v2c_ptr, const int32_t* v2c_ent, const double* fq, double* load,
                     void* stream) {
  CM_CHECK_ARG(cells && coords && v2c_ptr && v2c_ent && fq && load, "null pointer");
  CM_CHECK_ARG(ncomp == 1 || ncomp == 2, "ncomp must be 1 or 2");
  CM_CHECK_ARG(nv > 0 && ncells > 0, "empty mesh");
  return assemble_load(ncomp, weighted, supg_stab, degree, nv, cells, coords, v2c_ptr, v2c_ent, fq,
                       load, (cudaStream_t)stream);
}

int cm_assemble_c0ip_facets(const cm_pq_params* prm, int32_t nv, int32_t nf, const int32_t* cells,
                            const double* coords, const int32_t* fverts, const int32_t* fcell,
                            const int32_t* v2f_ptr, const int32_t* v2f_ent, const int32_t* nrowptr,
                            const int32_t* ncol, double* Kq, void* stream) {
  CM_CHECK_ARG(prm && cells && coords && fverts && fcell && v2f_ptr && v2f_ent && nrowptr && ncol &&
                   Kq, "null pointer");
  (void)nf;
  return assemble_c0ip(*prm, nv, cells, coords, fverts, fcell, v2f_ptr, v2f_ent, nrowptr, ncol, Kq,
                       (cudaStream_t)stream);
}

int cm_assemble_exterior_facets(const cm_pq_params* prm, int32_t nv, int32_t nfe,
                                const double* coords, const int32_t* efverts, const int32_t* efopp,
                                const int32_t* efkind, const double* qdata, const int32_t* v2e_ptr,
                                const int32_t* v2e_ent, const int32_t* nrowptr, const int32_t* ncol,
                                double* Kq, double* cq, void* stream) {
  CM_CHECK_ARG(prm && coords && efverts && efopp && efkind && v2e_ptr && v2e_ent && nrowptr &&
                   ncol && Kq && cq, "null pointer");
  (void)nfe;
  return assemble_ext(*prm, nv, coords, efverts, efopp, efkind, qdata, v2e_ptr, v2e_ent, nrowptr,
                      ncol, Kq, cq, (cudaStream_t)stream);
}

int cm_apply_dirichlet(int32_t bs, int32_t nrows, const int32_t* rows, const int32_t* diag_slot,
                       const double* g, const double* x, const int32_t* nrowptr, double* vals,
                       double* b, void* stream) {
  CM_CHECK_ARG(bs == 1 || bs == 2, "bs must be 1 or 2");
  CM_CHECK_ARG(nrows >= 0, "negative row count");
  if (nrows == 0) return CM_OK;
  CM_CHECK_ARG(rows && diag_slot && g && x && nrowptr && b, "null pointer");
  return apply_dirichlet(bs, nrows, rows, diag_slot, g, x, nrowptr, vals, b, (cudaStream_t)stream);
}

int cm_spmv(int32_t bs, int32_t nv, const int32_t* nrowptr, const int32_t* ncol, const double* vals,
            const double* x, double* y, void* stream) {
  CM_CHECK_ARG(bs == 1 || bs == 2, "bs must be 1 or 2");
  CM_CHECK_ARG(nrowptr && ncol && vals && x && y, "null pointer");
  CM_CHECK_ARG(x != y, "x and y must not alias");
  return spmv(bs, nv, nrowptr, ncol, vals, x, y, 1.0, nullptr, (cudaStream_t)stream);
}

int64_t cm_pqs_workspace_bytes(int32_t nx, int32_t ny, int32_t restart) {
  if (nx < 2 || ny < 2 || restart < 1) return 0;
  return (int64_t)pqs_workspace_bytes(nx, ny, restart);
}

int cm_pqs_setup(int32_t nx, int32_t ny, int32_t restart, const int32_t* nrowptr,
                 const int32_t* ncol, const double* vals, void* workspace, int64_t workspace_bytes,
                 void* stream) {
  CM_CHECK_ARG(nx >= 2 && ny >= 2 && restart >= 1, "bad grid or restart");
  CM_CHECK_ARG(nrowptr && ncol && vals && workspace, "null pointer");
  return pqs_setup(nx, ny, restart, nrowptr, ncol, vals, workspace, (size_t)workspace_bytes,
                   (cudaStream_t)stream);
}

int cm_pqs_solve(int32_t nx, int32_t ny, int32_t restart, const double* b, double* x, double rtol,
                 double atol, int32_t maxit, void* workspace, int64_t workspace_bytes,
                 int32_t* iters_host, double* resid_host, void* stream) {
  CM_CHECK_ARG(nx >= 2 && ny >= 2 && restart >= 1 && maxit >= 1, "bad grid, restart or maxit");
  CM_CHECK_ARG(b && x && workspace, "null pointer");
  int it = 0;
  double rr = 0.0;
  const int rc = pqs_solve(nx, ny, restart, b, x, rtol, atol, maxit, workspace,
                           (size_t)workspace_bytes, &it, &rr, (cudaStream_t)stream);
  if (iters_host) *iters_host = it;
  if (resid_host) *resid_host = rr;
  return rc;
}

int64_t cm_krylov_workspace_bytes(int32_t bs, int32_t nv, int32_t restart) {
  if ((bs != 1 && bs != 2) || nv < 1 || restart < 1) return 0;
  return (int64_t)krylov_workspace_bytes(bs, nv, restart);
}

int cm_krylov_solve(int32_t bs, int32_t nv, const int32_t* nrowptr, const int32_t* ncol,
                    const int32_t* diag_slot, const double* vals, const double* b, double* x,
                    double rtol, double atol, int32_t maxit, int32_t restart, void* workspace,
                    int64_t workspace_bytes, int32_t* iters_host, double* resid_host, void* stream) {
  CM_CHECK_ARG(bs == 1 || bs == 2, "bs must be 1 or 2");
  CM_CHECK_ARG(nv >= 1 && restart >= 1 && maxit >= 1, "bad size, restart or maxit");
  CM_CHECK_ARG(nrowptr && ncol && diag_slot && vals && b && x && workspace, "null pointer");
  int it = 0;
  double rr = 0.0;
  const int rc = krylov_solve(bs, nv, nrowptr, ncol, diag_slot, vals, b, x, rtol, atol, maxit,
                              restart, workspace, (size_t)workspace_bytes, &it, &rr,
                              (cudaStream_t)stream);
  if (iters_host) *iters_host = it;
  if (resid_host) *resid_host = rr;
  return rc;
}

int64_t cm_bicgstab_workspace_bytes(int32_t bs, int32_t nv) {
  if ((bs != 1 && bs != 2) || nv < 1) return -1;
  return (int64_t)bicgstab_workspace_bytes(bs, nv);
}

int cm_bicgstab_solve(int32_t bs, int32_t nv, const int32_t* nrowptr, const int32_t* ncol,
                      const int32_t* diag_slot, const double* vals, const double* b, double* x, double rtol,
                      double atol, int32_t maxit, void* workspace, int64_t workspace_bytes,
                      int32_t* iters_host, double* resid_host, void* stream) {
  CM_CHECK_ARG(bs == 1 || bs == 2, "bs must be 1 or 2");
  CM_CHECK_ARG(nv >= 1 && maxit >= 1, "bad size or maxit");
  CM_CHECK_ARG(nrowptr && ncol && diag_slot && vals && b && x && workspace, "null pointer");
  int it = 0;
  double rr = 0.0;
  const int rc = bicgstab_csr_solve(bs, nv, nrowptr, ncol, diag_slot, vals, b, x, rtol, atol, maxit, workspace,
                                    (size_t)workspace_bytes, &it, &rr, (cudaStream_t)stream);
  if (iters_host) *iters_host = it;
  if (resid_host) *resid_host = rr;
  return rc;
}

int cm_stencil_line_factor(int32_t nx, int32_t ny, const double* A, double* lf, double* iu,
                           double* cu, int32_t* err_flag, void* stream) {
  CM_CHECK_ARG(nx >= 1 && ny >= 1 && A && lf && iu && cu && err_flag, "bad argument");
  return st_line_factor(nx, ny, A, lf, iu, cu, err_flag, (cudaStream_t)stream);
}

int cm_stencil_zebra(int32_t nx, int32_t ny, int32_t colour, const double* A, const double* lf,
                     const double* iu, const double* cu, const double* b, double* x,
                     int32_t x_is_zero, double* tmp, void* stream) {
  CM_CHECK_ARG(nx >= 1 && ny >= 1 && (colour == 0 || colour == 1), "bad argument");
  CM_CHECK_ARG(A && lf && iu && cu && b && x, "null pointer");
  return st_zebra(nx, ny, colour, A, lf, iu, cu, b, x, x_is_zero, (cudaStream_t)stream, 0, -1, tmp);
}

int cm_stencil_residual(int32_t nx, int32_t ny, const double* A, const double* b, const double* x,
                        double* r, void* stream) {
  CM_CHECK_ARG(nx >= 1 && ny >= 1 && A && x && r, "bad argument");
  return st_residual(nx, ny, A, b, x, r, (cudaStream_t)stream);
}

int cm_dev_zebra_times(unsigned long long* host_out, int32_t n) {
  extern int cm_zebra_debug_times_impl(unsigned long long*, int);
  return cm_zebra_debug_times_impl(host_out, n);
}

// ---- multi-GPU (strip-partitioned) entry points ---------------------------------------------------
void* cm_ipc_alloc(int64_t bytes) {
  void* p = nullptr;
  if (cudaMalloc(&p, (size_t)bytes) != cudaSuccess) {
    set_error("cm_ipc_alloc: cudaMalloc(%lld) failed", (long long)bytes);
    cudaGetLastError();
    return nullptr;
  }
  cudaMemset(p, 0, (size_t)bytes);
  return p;
}

int cm_ipc_free(void* p) {
  CM_CUDA(cudaFree(p));
  return CM_OK;
}

int cm_ipc_get_handle(void* p, void* handle64_host) {
  CM_CHECK_ARG(p && handle64_host, "null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  CM_CUDA(cudaIpcGetMemHandle((cudaIpcMemHandle_t*)handle64_host, p));
  return CM_OK;
}

void* cm_ipc_open(const void* handle64_host) {
  void* p = nullptr;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64_host, sizeof(h));
  if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
    set_error("cm_ipc_open: %s", cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
  return p;
}

int cm_ipc_close(void* p) {
  CM_CUDA(cudaIpcCloseMemHandle(p));
  return CM_OK;
}

void* cm_pqs_create(int32_t nx, int32_t ny, int32_t restart, void* workspace, int64_t workspace_bytes,
                    const cm_dist* D) {
  if (nx < 2 || ny < 2 || restart < 1 || workspace == nullptr) {
    set_error("cm_pqs_create: bad grid, restart or workspace");
    return nullptr;
  }
  if (D != nullptr && !(D->world >= 1 && D->world <= 8 && D->rank >= 0 && D->rank < D->world)) {
    set_error("cm_pqs_create: bad rank/world");
    return nullptr;
  }
  return pqs_create(nx, ny, restart, workspace, (size_t)workspace_bytes, D);
}

int cm_pqs_destroy(void* ctx) { return ctx ? pqs_destroy(ctx) : CM_OK; }

int cm_pqs_ctx_setup(void* ctx, const int32_t* nrowptr, const int32_t* ncol, const double* vals, void* stream) {
  CM_CHECK_ARG(ctx && nrowptr && ncol && vals, "null pointer");
  return pqs_ctx_setup(ctx, nrowptr, ncol, vals, (cudaStream_t)stream);
}

int cm_pqs_ctx_solve(void* ctx, const double* b, double* x, double rtol, double atol, int32_t maxit,
                     int32_t* iters_host, double* resid_host, void* stream) {
  CM_CHECK_ARG(ctx && b && x && maxit >= 1, "null pointer or bad maxit");
  int it = 0;
  double rr = 0.0;
  const int rc = pqs_ctx_solve(ctx, b, x, rtol, atol, maxit, &it, &rr, (cudaStream_t)stream);
  if (iters_host) *iters_host = it;
  if (resid_host) *resid_host = rr;
  return rc;
}

int cm_pqs_ctx_info(void* ctx, int32_t* out8) {
  CM_CHECK_ARG(ctx && out8, "null pointer");
  return pqs_ctx_info(ctx, out8);
}

int cm_mg_zebra(int32_t nx, int32_t ny, int32_t colour, const double* A, const double* lf, const double* iu,
                const double* cu, const double* b, double* x, int32_t x_is_zero, void* stream) {
  CM_CHECK_ARG(nx >= 1 && ny >= 1 && (colour == 0 || colour == 1), "bad grid or colour");
  CM_CHECK_ARG(A && lf && iu && cu && b && x, "null pointer");
  return mg_zebra(nx, ny, colour, A, lf, iu, cu, b, x, x_is_zero, Halo{}, 0, -1, (cudaStream_t)stream);
}

int64_t cm_mg_pack_f32_bytes(int32_t nx, int32_t ny) {
  if (nx < 1 || ny < 1) return 0;
  return (int64_t)(mg_pack_f32_doubles(nx, ny) * sizeof(double));
}

int cm_mg_pack_f32(int32_t nx, int32_t ny, const double* A, const double* lf, const double* iu, const double* cu,
                   void* pack, void* stream) {
  CM_CHECK_ARG(nx >= 1 && ny >= 1 && A && lf && iu && cu && pack, "bad argument");
  CM_CHECK_ARG(((uintptr_t)pack & 15) == 0, "the pack must be 16-byte aligned");
  return mg_pack_f32(nx, ny, A, lf, iu, cu, (float*)pack, 0, -1, (cudaStream_t)stream);
}

int cm_mg_zebra_f32(int32_t nx, int32_t ny, int32_t colour, const double* A, const double* lf, const double* iu,
                    const double* cu, const double* b, double* x, int32_t x_is_zero, const void* pack, void* stream) {
  CM_CHECK_ARG(nx >= 1 && ny >= 1 && (colour == 0 || colour == 1), "bad grid or colour");
  CM_CHECK_ARG(A && lf && iu && cu && b && x && pack, "null pointer");
  CM_CHECK_ARG(mg_ring_fits(nx), "lines too short (or too long) for the long-line kernel");
  return mg_zebra(nx, ny, colour, A, lf, iu, cu, b, x, x_is_zero, Halo{}, 0, -1, (cudaStream_t)stream, false,
                  (const float*)pack);
}

int cm_mg_zebra_mid(int32_t mode, int32_t nx, int32_t ny, int32_t colour, const double* A, const double* lf,
                    const double* iu, const double* cu, const double* b, double* x, const double* ec, double* bc,
                    void* stream) {
  CM_CHECK_ARG(nx >= 1 && ny >= 1 && (colour == 0 || colour == 1) && mode >= 0 && mode <= 3, "bad grid, colour or mode");
  CM_CHECK_ARG(A && lf && iu && cu && b && x, "null pointer");
  CM_CHECK_ARG(mode != 2 || bc, "mode 2 writes the coarse right-hand side: bc is null");
  CM_CHECK_ARG(mode != 3 || ec, "mode 3 reads the coarse correction: ec is null");
  return mg_zebra_mid(mode, nx, ny, colour, A, lf, iu, cu, b, x, ec, bc, (cudaStream_t)stream);
}

int cm_mg_ring_debug_times(uint64_t* host_out, int32_t n) {
  CM_CHECK_ARG(host_out && n > 0 && n <= 64 * 16 * 8, "bad argument");
  return mg_ring_debug_times((unsigned long long*)host_out, n);
}

int cm_mg_smem_debug_times(uint64_t* host_out, int32_t n) {
  CM_CHECK_ARG(host_out && n > 0 && n <= 64, "bad argument");
  return mg_smem_debug_times((unsigned long long*)host_out, n);
}

int cm_mg_resid_restrict(int32_t nxf, int32_t nyf, const double* A, const double* b, const double* x,
                         double* rc, void* stream) {
  CM_CHECK_ARG(nxf >= 2 && nyf >= 2 && !(nxf & 1) && !(nyf & 1), "fine grid must have even nx, ny");
  CM_CHECK_ARG(A && b && x && rc, "null pointer");
  return mg_resid_restrict(nxf, nyf, A, b, x, rc, Halo{}, 0, -1, (cudaStream_t)stream);
}

int cm_mg_prolong_odd(int32_t nxc, int32_t nyc, const double* ec, double* xf, void* stream) {
  CM_CHECK_ARG(nxc >= 1 && nyc >= 1 && ec && xf, "bad argument");
  return mg_prolong_odd(nxc, nyc, ec, xf, Halo{}, 0, 0, -1, (cudaStream_t)stream);
}

int64_t cm_pqs_state_offset(int32_t nx, int32_t ny, int32_t restart) {
  return (int64_t)pqs_state_offset_bytes(nx, ny, restart);
}

int cm_assemble_pq_cells_structured_rows(const cm_pq_params* prm, int32_t nx, int32_t ny,
                                         const double* coords, const int32_t* nrowptr,
                                         const double* u, const double* p_old, const double* load,
                                         double* vals, double* b, int32_t row_begin,
                                         int32_t row_end, void* stream) {
  CM_CHECK_ARG(prm && coords && nrowptr && u && b, "null pointer");
  CM_CHECK_ARG(nx >= 1 && ny >= 1 && row_begin >= 0 && row_end <= ny + 1, "bad row range");
  return assemble_pq_sweep(*prm, nx, ny, coords, nrowptr, u, p_old, load, vals, b,
                           (cudaStream_t)stream, row_begin, row_end);
}

int cm_assemble_pq_cells_structured_blocks(const cm_pq_params* prm, int32_t nx, int32_t ny, const double* coords,
                                           const int32_t* nrowptr, const double* u, const double* p_old,
                                           const double* load, double* vals, double* b, const cm_pq_blocks* blocks,
                                           int32_t row_begin, int32_t row_end, void* stream) {
  CM_CHECK_ARG(prm && coords && nrowptr && u && b && vals && blocks, "null pointer");
  CM_CHECK_ARG(blocks->PP && blocks->PQ && blocks->QP && blocks->QQ && blocks->sp && blocks->sq, "null block pointer");
  if (row_end < 0) { row_begin = 0; row_end = ny + 1; }
  CM_CHECK_ARG(nx >= 1 && ny >= 1 && row_begin >= 0 && row_end <= ny + 1, "bad row range");
  return assemble_pq_sweep(*prm, nx, ny, coords, nrowptr, u, p_old, load, vals, b, (cudaStream_t)stream, row_begin,
                           row_end, blocks);
}

int cm_apply_dirichlet_blocks(int32_t nrows, const int32_t* rows, int32_t nv, const cm_pq_blocks* blocks, void* stream) {
  CM_CHECK_ARG(nrows >= 0 && nv > 0 && blocks && (nrows == 0 || rows), "bad argument");
  CM_CHECK_ARG(blocks->PP && blocks->PQ && blocks->QP && blocks->QQ && blocks->sp && blocks->sq, "null block pointer");
  return apply_dirichlet_blocks(nrows, rows, nv, *blocks, (cudaStream_t)stream);
}

int cm_pqs_ctx_blocks(void* ctx, cm_pq_blocks* out) {
  CM_CHECK_ARG(ctx && out, "null pointer");
  return pqs_ctx_blocks(ctx, out);
}

int cm_pqs_ctx_setup_from_blocks(void* ctx, void* stream) {
  CM_CHECK_ARG(ctx, "null pointer");
  return pqs_ctx_setup(ctx, nullptr, nullptr, nullptr, (cudaStream_t)stream, true);
}

int cm_dist_exchange_rows(cm_dist* D, int64_t array_offset_bytes, int32_t row_doubles,
                          int32_t row_begin, int32_t row_end, void* stream) {
  CM_CHECK_ARG(D && (array_offset_bytes % 8) == 0 && row_doubles > 0 && row_end > row_begin, "bad argument");
  unsigned long long od = (unsigned long long)(array_offset_bytes / 8) + (unsigned long long)row_begin * row_doubles;
  unsigned long long ou = (unsigned long long)(array_offset_bytes / 8) + (unsigned long long)(row_end - 1) * row_doubles;
  int n = row_doubles;
  return dist_halo2(D, &od, &n, 1, &ou, &n, 1, (cudaStream_t)stream);
}

int cm_dist_allgather_rows(cm_dist* D, int64_t array_offset_bytes, int32_t row_doubles,
                           int32_t row_begin, int32_t row_end, void* stream) {
  CM_CHECK_ARG(D && (array_offset_bytes % 8) == 0 && row_doubles > 0 && row_end > row_begin, "bad argument");
  unsigned long long off = (unsigned long long)(array_offset_bytes / 8) + (unsigned long long)row_begin * row_doubles;
  int n = (row_end - row_begin) * row_doubles;
  return dist_allgather(D, &off, &n, 1, (cudaStream_t)stream);
}

int cm_dist_allreduce_sum(cm_dist* D, double* vals, int32_t n, void* stream) {
  CM_CHECK_ARG(D && vals && n > 0 && n <= 64, "bad argument");
  return dist_allreduce(D, vals, n, (cudaStream_t)stream);
}

int cm_pqs_setup_dist(int32_t nx, int32_t ny, int32_t restart, const int32_t* nrowptr,
                      const int32_t* ncol, const double* vals, void* workspace,
                      int64_t workspace_bytes, cm_dist* D, void* stream) {
  CM_CHECK_ARG(nx >= 2 && ny >= 2 && restart >= 1, "bad grid or restart");
  CM_CHECK_ARG(nrowptr && ncol && vals && workspace && D, "null pointer");
  CM_CHECK_ARG(D->world >= 1 && D->world <= 8 && D->rank >= 0 && D->rank < D->world, "bad rank/world");
  return pqs_setup_dist(nx, ny, restart, nrowptr, ncol, vals, workspace, (size_t)workspace_bytes, D,
                        (cudaStream_t)stream);
}

int cm_pqs_solve_dist(int32_t nx, int32_t ny, int32_t restart, const double* b, double* x,
                      double rtol, double atol, int32_t maxit, void* workspace,
                      int64_t workspace_bytes, cm_dist* D, int32_t* iters_host, double* resid_host,
                      void* stream) {
  CM_CHECK_ARG(nx >= 2 && ny >= 2 && restart >= 1 && maxit >= 1, "bad grid, restart or maxit");
  CM_CHECK_ARG(b && x && workspace && D, "null pointer");
  int it = 0;
  double rr = 0.0;
  const int rc = pqs_solve_dist(nx, ny, restart, b, x, rtol, atol, maxit, workspace,
                                (size_t)workspace_bytes, D, &it, &rr, (cudaStream_t)stream);
  if (iters_host) *iters_host = it;
  if (resid_host) *resid_host = rr;
  return rc;
}

/* per-kernel-class timing (diagnostics for bench.py): enable, run, then report.  Classes:
 * 0 zebra_line_kernel on the finest grid, 1 pq_stencil_spmv_kernel, 2 Jacobian assembly. */
int cm_profile_enable(int32_t on) {
  g_prof_on = on != 0;
  return CM_OK;
}

int cm_profile_report(int32_t cls, int64_t* launches, double* total_ms, double* total_bytes) {
  CM_CHECK_ARG(cls >= 0 && cls < kProfClasses && launches && total_ms && total_bytes, "bad argument");
  CM_CUDA(cudaDeviceSynchronize());
  int64_t n = 0;
  double ms = 0.0, by = 0.0;
  for (auto& r : g_prof) {
    if (r.cls != cls) continue;
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) == cudaSuccess) { ms += t; by += r.bytes; ++n; }
  }
  cudaGetLastError();
  *launches = n;
  *total_ms = ms;
  *total_bytes = by;
  return CM_OK;
}

int cm_profile_reset(void) {
  for (auto& r : g_prof) { g_prof_pool.push_back(r.a); g_prof_pool.push_back(r.b); }
  g_prof.clear();
  return CM_OK;
}

int64_t cm_blu_workspace_bytes(int32_t bs, int32_t nv, int32_t nbn) {
  if ((bs != 1 && bs != 2) || nv < 1 || nbn < 1) {
    set_error("cm_blu_workspace_bytes: bad argument");
    return CM_E_INVALID;
  }
  return (int64_t)blu_workspace_bytes(bs, nv, nbn);
}

void* cm_blu_create(void) { return blu_create(); }

int cm_blu_destroy(void* ctx) { return blu_destroy(ctx); }

int cm_blu_factor(void* ctx, int32_t bs, int32_t nv, int32_t nbn, const int32_t* perm,
                  const int32_t* iperm, const int32_t* nrowptr, const int32_t* ncol,
                  const double* vals, void* workspace, int64_t workspace_bytes, int32_t* info_host,
                  void* stream) {
  CM_CHECK_ARG(ctx != nullptr, "null context (cm_blu_create)");
  CM_CHECK_ARG(bs == 1 || bs == 2, "bs must be 1 or 2");
  CM_CHECK_ARG(nv >= 1 && nbn >= 1, "bad size");
  CM_CHECK_ARG((int64_t)bs * nbn <= 16384, "block size above 16384 dofs");
  CM_CHECK_ARG(perm && iperm && nrowptr && ncol && vals && workspace, "null pointer");
  return blu_factor(ctx, bs, nv, nbn, perm, iperm, nrowptr, ncol, vals, workspace,
                    (size_t)workspace_bytes, info_host, (cudaStream_t)stream);
}

int cm_blu_solve(void* ctx, int32_t bs, int32_t nv, int32_t nbn, const int32_t* perm, const double* b,
                 double* x, void* workspace, int64_t workspace_bytes, void* stream) {
  CM_CHECK_ARG(ctx != nullptr, "null context (cm_blu_create)");
  CM_CHECK_ARG(bs == 1 || bs == 2, "bs must be 1 or 2");
  CM_CHECK_ARG(nv >= 1 && nbn >= 1, "bad size");
  CM_CHECK_ARG(perm && b && x && workspace, "null pointer");
  return blu_solve(ctx, bs, nv, nbn, perm, b, x, workspace, (size_t)workspace_bytes,
                   (cudaStream_t)stream);
}

int64_t cm_functional_workspace_doubles(void) { return (int64_t)functional_workspace_doubles(); }

int cm_functional_error_norms(int32_t nq, const double* rule_host, int32_t weighted,
                              int32_t x_derivative_only, int32_t ncells, const int32_t* cells,
                              const double* coords, const double* uh, int32_t uh_stride,
                              const double* exact_q, const double* grad_q, double* work, void* stream) {
  CM_CHECK_ARG(ncells >= 1 && uh_stride >= 1, "bad size");
  CM_CHECK_ARG(rule_host && cells && coords && uh && exact_q && grad_q && work, "null pointer");
  return functional_error_norms(nq, rule_host, weighted, x_derivative_only, ncells, cells, coords, uh,
                                uh_stride, exact_q, grad_q, work, (cudaStream_t)stream);
}

int cm_project_flux_dg0(int32_t nq, const double* rule_host, double D1, double D2, int32_t ncells,
                        const int32_t* cells, const double* coords, const double* u, double* flux,
                        void* stream) {
  CM_CHECK_ARG(ncells >= 1, "bad size");
  CM_CHECK_ARG(rule_host && cells && coords && u && flux, "null pointer");
  return project_flux_dg0(nq, rule_host, D1, D2, ncells, cells, coords, u, flux, (cudaStream_t)stream);
}

int cm_functional_boundary_flux(int32_t nfe, const int32_t* efverts, const int32_t* efopp,
                                const int32_t* efcell, const double* coords, const double* flux,
                                double* out3, void* stream) {
  CM_CHECK_ARG(nfe >= 0, "bad size");
  CM_CHECK_ARG(coords && flux && out3 && (nfe == 0 || (efverts && efopp && efcell)), "null pointer");
  return functional_boundary_flux(nfe, efverts, efopp, efcell, coords, flux, out3, (cudaStream_t)stream);
}

int cm_mixed_m0_cells(int32_t nq, const double* rule_host, double D1, double D2, double geps,
                      int32_t ncells, const int32_t* cells, const double* coords,
                      const int32_t* cell_edges, const double* sign, int32_t nedges, const double* u,
                      const double* f2q, const double* f3q, double* Je, double* Fe, void* stream) {
  CM_CHECK_ARG(ncells >= 1 && nedges >= 3, "bad size");
  CM_CHECK_ARG(rule_host && cells && coords && cell_edges && sign && u && f2q && f3q && Fe, "null pointer");
  return mixed_m0_cells(nq, rule_host, D1, D2, geps, ncells, cells, coords, cell_edges, sign, nedges, u,
                        f2q, f3q, Je, Fe, (cudaStream_t)stream);
}

int cm_csr_gather(int32_t nslots, const int32_t* ptr, const int32_t* idx, const double* src, double* out,
                  int32_t accumulate, void* stream) {
  CM_CHECK_ARG(nslots >= 0, "bad size");
  CM_CHECK_ARG(nslots == 0 || (ptr && idx && src && out), "null pointer");
  return csr_gather(nslots, ptr, idx, src, out, accumulate, (cudaStream_t)stream);
}

int cm_csr_add_mapped(int32_t n, const int32_t* map, const double* src, double* dst, void* stream) {
  CM_CHECK_ARG(n >= 0, "bad size");
  CM_CHECK_ARG(n == 0 || (map && src && dst), "null pointer");
  return add_mapped(n, map, src, dst, (cudaStream_t)stream);
}

int64_t cm_mixed_workspace_doubles(void) { return (int64_t)mixed_workspace_doubles(); }

int cm_mixed_m0_errors(int32_t nq, const double* rule_host, int32_t ncells, const int32_t* cells,
                       const double* coords, const int32_t* cell_edges, const double* sign,
                       int32_t nedges, const double* u, const double* exact_q, double* work, void* stream) {
  CM_CHECK_ARG(ncells >= 1 && nedges >= 3, "bad size");
  CM_CHECK_ARG(rule_host && cells && coords && cell_edges && sign && u && exact_q && work, "null pointer");
  return mixed_m0_errors(nq, rule_host, ncells, cells, coords, cell_edges, sign, nedges, u, exact_q, work,
                         (cudaStream_t)stream);
}

int cm_mixed_m1_cells(int32_t nq, const double* rule_host, double D1, double D2, double geps, double gcoef,
                      int32_t qmode, double q_react, double q_x2p, double q_x2q, double supg_stab,
                      int32_t ncells, int32_t nedges, int32_t nverts, const int32_t* cells,
                      const double* coords, const int32_t* cell_edges, const double* sign, const double* u,
                      const double* f2q, const double* f3q, double* Je, double* Fe, void* stream) {
  CM_CHECK_ARG(ncells >= 1 && nedges >= 3 && nverts >= 3, "bad size");
  CM_CHECK_ARG(rule_host && cells && coords && cell_edges && sign && u && f2q && f3q && Fe, "null pointer");
  CM_CHECK_ARG(qmode >= 0 && qmode <= 2, "qmode must be 0 (integrated by parts), 1 (Galerkin / SUPG) or 2 (0 with r2vec = (-1,0))");
  return mixed_m1_cells(nq, rule_host, D1, D2, geps, gcoef, qmode, q_react, q_x2p, q_x2q, supg_stab, ncells, nedges,
                        nverts, cells, coords, cell_edges, sign, u, f2q, f3q, Je, Fe, (cudaStream_t)stream);
}

int cm_mixed_m1_errors(int32_t nq, const double* rule_host, int32_t full_grad, int32_t ncells, int32_t nedges,
                       int32_t nverts,
                       const int32_t* cells, const double* coords, const int32_t* cell_edges,
                       const double* sign, const double* u, const double* exact_q, double* work,
                       void* stream) {
  CM_CHECK_ARG(ncells >= 1 && nedges >= 3 && nverts >= 3, "bad size");
  CM_CHECK_ARG(rule_host && cells && coords && cell_edges && sign && u && exact_q && work, "null pointer");
  return mixed_m1_errors(nq, rule_host, full_grad, ncells, nedges, nverts, cells, coords, cell_edges, sign, u,
                         exact_q, work, (cudaStream_t)stream);
}

int cm_cg2_exterior_facets(int32_t nq, const double* rule_host, double stab2, int32_t nfe,
                           const int32_t* fverts, const int32_t* fcell, const int32_t* kind,
                           const int32_t* cells, const double* coords, const double* qex, const double* pex,
                           double* Jx, double* cx, double* nat, void* stream) {
  CM_CHECK_ARG(nfe >= 0, "bad size");
  CM_CHECK_ARG(nfe == 0 || (rule_host && fverts && fcell && kind && cells && coords && qex && pex && Jx && cx && nat),
               "null pointer");
  return cg2_exterior_facets(nq, rule_host, stab2, nfe, fverts, fcell, kind, cells, coords, qex, pex, Jx, cx, nat,
                             (cudaStream_t)stream);
}

int cm_cg2_interior_facets(int32_t nq, const double* rule_host, double gamma, int32_t nfi,
                           const int32_t* fverts, const int32_t* fcell, const int32_t* cells,
                           const double* coords, const double* hbar, double* Ji, void* stream) {
  CM_CHECK_ARG(nfi >= 0, "bad size");
  CM_CHECK_ARG(nfi == 0 || (rule_host && fverts && fcell && cells && coords && Ji), "null pointer");
  return cg2_interior_facets(nq, rule_host, gamma, nfi, fverts, fcell, cells, coords, hbar, Ji,
                             (cudaStream_t)stream);
}

int cm_mixed_p2_cells(int32_t nq, const double* rule_host, double D1, double D2, double concentration,
                      double inv_dt, int32_t ncells, int32_t nedges, int32_t nverts, const int32_t* cells,
                      const double* coords, const int32_t* cell_edges, const double* sign, const double* u,
                      const double* p_old, double* Je, double* Fe, void* stream) {
  CM_CHECK_ARG(ncells >= 1 && nedges >= 3 && nverts >= 3 && concentration > 0.0, "bad size or concentration");
  CM_CHECK_ARG(rule_host && cells && coords && cell_edges && sign && u && Fe, "null pointer");
  CM_CHECK_ARG(inv_dt == 0.0 || p_old != nullptr, "transient form needs p_old");
  return mixed_p2_cells(nq, rule_host, D1, D2, concentration, inv_dt, ncells, nedges, nverts, cells, coords,
                        cell_edges, sign, u, p_old, Je, Fe, (cudaStream_t)stream);
}

int64_t cm_p1_sparsity_workspace_bytes(int32_t nv, int32_t ncells, int32_t nextra) {
  if (nv < 1 || ncells < 1 || nextra < 0) {
    set_error("cm_p1_sparsity_workspace_bytes: bad argument");
    return CM_E_INVALID;
  }
  return (int64_t)p1_sparsity_workspace_bytes(nv, ncells, nextra);
}

int cm_build_p1_sparsity(int32_t nv, int32_t ncells, const int32_t* cells, int32_t nextra,
                         const int32_t* extra_pairs, int32_t* rowptr, int32_t* col, int64_t col_capacity,
                         int64_t* nnz_host, void* workspace, int64_t workspace_bytes, void* stream) {
  CM_CHECK_ARG(nv >= 1 && ncells >= 1 && nextra >= 0, "bad size");
  CM_CHECK_ARG(cells && rowptr && col && workspace && (nextra == 0 || extra_pairs), "null pointer");
  long long nnz = 0;
  const int rc = build_p1_sparsity(nv, ncells, cells, nextra, extra_pairs, rowptr, col, col_capacity, &nnz,
                                   workspace, (size_t)workspace_bytes, (cudaStream_t)stream);
  if (nnz_host) *nnz_host = nnz;
  return rc;
}

int64_t cm_vec_scratch_doubles(void) { return (int64_t)vec_scratch_doubles(); }

int cm_dot_batch(int64_t n, const double* V, int64_t vstride, int32_t nvec, const double* w, double* out,
                 double* scratch, void* stream) {
  CM_CHECK_ARG(n >= 1 && nvec >= 1 && vstride >= n, "bad size");
  CM_CHECK_ARG(V && w && out && scratch, "null pointer");
  return multi_dot(n, V, vstride, nvec, w, out, scratch, (cudaStream_t)stream);
}

int cm_axpy_batch(int64_t n, const double* V, int64_t vstride, int32_t nvec, const double* h, double sign,
                  double* w, double* norm2_out, double* scratch, void* stream) {
  CM_CHECK_ARG(n >= 1 && nvec >= 1 && vstride >= n, "bad size");
  CM_CHECK_ARG(V && h && w && scratch, "null pointer");
  return multi_axpy(n, V, vstride, nvec, h, sign, w, norm2_out, scratch, (cudaStream_t)stream);
}

}  // extern "C"
