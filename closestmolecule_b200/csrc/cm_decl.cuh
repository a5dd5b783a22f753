// Internal (C++) launchers behind the C ABI.
#pragma once
#include "cm_common.cuh"

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
