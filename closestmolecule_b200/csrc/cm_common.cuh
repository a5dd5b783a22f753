// Shared device helpers: error handling, FIAT quadrature tables, P1 geometry, the G nonlinearity.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/closestmol_b200.h"

namespace cm {

void set_error(const char* fmt, ...);

#define CM_CHECK_ARG(cond, msg)                  \
  do {                                           \
    if (!(cond)) {                               \
      cm::set_error("%s: %s", __func__, msg);    \
      return CM_E_INVALID;                       \
    }                                            \
  } while (0)

#define CM_CUDA(call)                                                                  \
  do {                                                                                 \
    cudaError_t e_ = (call);                                                           \
    if (e_ != cudaSuccess) {                                                           \
      cm::set_error("%s: CUDA error %s at %s:%d", __func__, cudaGetErrorString(e_), __FILE__, \
                    __LINE__);                                                         \
      return CM_E_CUDA;                                                                \
    }                                                                                  \
  } while (0)

extern std::atomic<long long> g_launches;  // kernels launched by this library (bench.py's gpu_launches);
                                           // atomic: the two chains of the direct solver launch from two threads
extern double g_alg_bytes;    // algorithmic (compulsory) bytes of the kernels launched so far (Newton-step roofline)
#define CM_BYTES(x) (cm::g_alg_bytes += (double)(x))
#define CM_LAUNCH_CHECK()      \
  do {                         \
    ++cm::g_launches;          \
    CM_CUDA(cudaGetLastError()); \
  } while (0)

// ---- optional per-kernel-class timing with CUDA events on the launching stream (bench.py's roofline
// object: average launch duration of the dominant kernel measured inside the timed region)
enum ProfClass { kProfZebraFine = 0, kProfSpmv = 1, kProfAssembly = 2, kProfClasses = 3 };
void prof_begin(int cls, double bytes, cudaStream_t st);
void prof_end(int cls, cudaStream_t st);
bool g_prof_is_on();

// ---- FIAT "default" triangle rules (15-digit constants as stored by FIAT).
// Barycentric (l0,l1,l2) = (1-xi-eta, xi, eta) and weight (sums to 1/2).
template <int NQ>
struct TriRule;

template <>
struct TriRule<3> {  // degree 2
  static __device__ __forceinline__ void get(int k, double& l0, double& l1, double& l2, double& w) {
    constexpr double a = 1.0 / 6.0, b = 2.0 / 3.0;
    const double X[3] = {a, a, b}, Y[3] = {a, b, a};
    l1 = X[k];
    l2 = Y[k];
    l0 = 1.0 - l1 - l2;
    w = 1.0 / 6.0;
  }
};

template <>
struct TriRule<6> {  // degree 4
  static __device__ __forceinline__ void get(int k, double& l0, double& l1, double& l2, double& w) {
    constexpr double a1 = 0.816847572980459, b1 = 0.091576213509771;
    constexpr double a2 = 0.108103018168070, b2 = 0.445948490915965;
    const double X[6] = {a1, b1, b1, a2, b2, b2}, Y[6] = {b1, a1, b1, b2, a2, b2};
    l1 = X[k];
    l2 = Y[k];
    l0 = 1.0 - l1 - l2;
    w = (k < 3) ? 0.109951743655322 / 2.0 : 0.223381589678011 / 2.0;
  }
};

template <>
struct TriRule<12> {  // degree 6
  static __device__ __forceinline__ void get(int k, double& l0, double& l1, double& l2, double& w) {
    constexpr double a1 = 0.873821971016996, b1 = 0.063089014491502;
    constexpr double a2 = 0.501426509658179, b2 = 0.249286745170910;
    constexpr double c1 = 0.636502499121399, c2 = 0.310352451033785, c3 = 0.053145049844816;
    const double X[12] = {a1, b1, b1, a2, b2, b2, c1, c1, c2, c2, c3, c3};
    const double Y[12] = {b1, a1, b1, b2, a2, b2, c2, c3, c1, c3, c1, c2};
    l1 = X[k];
    l2 = Y[k];
    l0 = 1.0 - l1 - l2;
    w = (k < 3) ? 0.050844906370207 / 2.0 : (k < 6 ? 0.116786275726379 / 2.0 : 0.082851075618374 / 2.0);
  }
};

__host__ __device__ inline int nq_of_degree(int degree) {
  if (degree <= 2) return 3;
  if (degree <= 4) return 6;
  return 12;
}

// Gauss-Legendre on [0,1], (degree+2)/2 points (3 at degree 4, 4 at degree 6)
__device__ __forceinline__ int facet_rule(int degree, double* s, double* w) {
  if (degree <= 4) {
    const double r = 0.7745966692414834;  // sqrt(3/5)
    s[0] = 0.5 * (1.0 - r);
    s[1] = 0.5;
    s[2] = 0.5 * (1.0 + r);
    w[0] = w[2] = 0.5 * (5.0 / 9.0);
    w[1] = 0.5 * (8.0 / 9.0);
    return 3;
  }
  const double r1 = 0.3399810435848563, r2 = 0.8611363115940526;
  const double w1 = 0.6521451548625461, w2 = 0.3478548451374538;
  s[0] = 0.5 * (1.0 - r2);
  s[1] = 0.5 * (1.0 - r1);
  s[2] = 0.5 * (1.0 + r1);
  s[3] = 0.5 * (1.0 + r2);
  w[0] = w[3] = 0.5 * w2;
  w[1] = w[2] = 0.5 * w1;
  return 4;
}

constexpr double kFourPi = 12.566370614359172953850573533118;

// P1 geometry of a triangle with the row-owner's vertex as local 0 (cyclic relabelling keeps the
// orientation; the FIAT rules are symmetric point sets, so the integrals are unchanged).
struct TriGeom {
  double x0, y0, x1, y1, x2, y2;
  double adet;                      // |det J|
  double gx0, gx1, gx2, gy0, gy1, gy2;  // grad phi_j
  __device__ __forceinline__ void build() {
    const double j00 = x1 - x0, j01 = x2 - x0, j10 = y1 - y0, j11 = y2 - y0;
    const double det = j00 * j11 - j01 * j10;
    const double inv = 1.0 / det;
    adet = fabs(det);
    gx1 = j11 * inv;
    gy1 = -j01 * inv;
    gx2 = -j10 * inv;
    gy2 = j00 * inv;
    gx0 = -(gx1 + gx2);
    gy0 = -(gy1 + gy2);
  }
  __device__ __forceinline__ double hmax() const {  // CellDiameter = longest edge
    const double e0 = (x1 - x0) * (x1 - x0) + (y1 - y0) * (y1 - y0);
    const double e1 = (x2 - x0) * (x2 - x0) + (y2 - y0) * (y2 - y0);
    const double e2 = (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1);
    return sqrt(fmax(e0, fmax(e1, e2)));
  }
};

// G, dG/dp, dG/dq (both branches of the conditional evaluated eagerly and selected, like uflacs)
template <int GKIND>
__device__ __forceinline__ void eval_G(const cm_pq_params& P, double x2, double p, double q,
                                       double& G, double& Gp, double& Gq) {
  if (GKIND == CM_G_NONE) {
    G = Gp = Gq = 0.0;
  } else if (GKIND == CM_G_POWER) {
    const double iq = 1.0 / (q + P.geps);
    const double t = P.gcoef * x2 * p * iq;  // gcoef x^2 p/(q+eps)
    G = t * p;
    Gp = 2.0 * t;
    Gq = -G * iq;
  } else {
    const bool cond = fabs(P.gs_cond_scale * p - q) > P.gs_thr;
    const double iq = 1.0 / q;               // may be inf/nan when cond is false: discarded
    const double t = p * iq * x2;
    G = cond ? t * p : P.gs_lin * p * x2;
    Gp = cond ? 2.0 * t : P.gs_lin * x2;
    Gq = cond ? -t * p * iq : 0.0;
  }
}

}  // namespace cm
