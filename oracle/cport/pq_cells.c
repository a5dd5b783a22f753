/* CPU restatement (plain C, one core) of the P-Q cell assembly -- TEST INFRASTRUCTURE / CPU BASELINE.
 *
 * What DOLFIN's serial Assembler::assemble_cells does [ext] when solver.solve() is called at
 * advection_diffusion_convergence.py:219: for every cell in ascending order, tabulate the 6x6
 * Jacobian block and the 6-vector residual of the form at :196-204 (degree-4 FIAT rule, weight
 * (4*pi*r1*r2)^2 of :135, Gstar of :21-27 or the power-law G of …MixedPrimal_convergence.py:43-44)
 * the way FFC writes it (a plain loop over quadrature points), then add the block into the CSR
 * matrix (MatSetValues ADD_VALUES; here through a precomputed cell->slot map, which is cheaper
 * than PETSc's search) and the vector.  Used by tests (cross-check of oracle/pq.py) and by
 * bench.py's cpu_baseline / --impl reference legs only.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef struct {
  double D1, D2, inv_dt, gcoef, geps, gs_thr, gs_cond_scale, gs_lin;
  double q_react, q_adv, q_x2p, q_x2q, q_divr, q_ibp, supg_stab;
  int32_t gkind, nq;
} pq_c_params;

static const double X4[6] = {0.816847572980459, 0.091576213509771, 0.091576213509771,
                             0.108103018168070, 0.445948490915965, 0.445948490915965};
static const double Y4[6] = {0.091576213509771, 0.816847572980459, 0.091576213509771,
                             0.445948490915965, 0.108103018168070, 0.445948490915965};
static const double W4[6] = {0.109951743655322 / 2, 0.109951743655322 / 2, 0.109951743655322 / 2,
                             0.223381589678011 / 2, 0.223381589678011 / 2, 0.223381589678011 / 2};

static void eval_G(const pq_c_params* P, double x2, double p, double q, double* G, double* Gp,
                   double* Gq) {
  if (P->gkind == 0) {
    *G = *Gp = *Gq = 0.0;
  } else if (P->gkind == 1) {
    double qe = q + P->geps;
    *G = P->gcoef * x2 * p * p / qe;
    *Gp = 2.0 * P->gcoef * x2 * p / qe;
    *Gq = -P->gcoef * x2 * p * p / (qe * qe);
  } else {
    int cond = fabs(P->gs_cond_scale * p - q) > P->gs_thr;
    *G = cond ? (p / q) * p * x2 : P->gs_lin * p * x2;
    *Gp = cond ? 2.0 * p * x2 / q : P->gs_lin * x2;
    *Gq = cond ? -(p * p) * x2 / (q * q) : 0.0;
  }
}

/* cells[3*C], coords[2*V], u[2*V] interleaved, p_old[V] or NULL, fq[C*6*2] or NULL,
 * slots[C*36] positions in vals of the (i,j) entries (cell-local order [p0,p1,p2,q0,q1,q2]),
 * vals (zeroed by the caller) or NULL, b[2*V] (zeroed by the caller). */
void pq_assemble_cells(const pq_c_params* P, int64_t C, const int32_t* cells, const double* coords,
                       const double* u, const double* p_old, const double* fq, const int64_t* slots,
                       double* vals, double* b) {
  const double fourpi = 12.566370614359172953850573533118;
  for (int64_t c = 0; c < C; ++c) {
    const int32_t* cv = cells + 3 * c;
    double x[3], y[3], pe[3], qe[3], po[3] = {0, 0, 0};
    for (int i = 0; i < 3; ++i) {
      x[i] = coords[2 * cv[i]];
      y[i] = coords[2 * cv[i] + 1];
      pe[i] = u[2 * cv[i]];
      qe[i] = u[2 * cv[i] + 1];
      if (p_old) po[i] = p_old[cv[i]];
    }
    double j00 = x[1] - x[0], j01 = x[2] - x[0], j10 = y[1] - y[0], j11 = y[2] - y[0];
    double det = j00 * j11 - j01 * j10, adet = fabs(det);
    double gx[3], gy[3];
    gx[1] = j11 / det; gy[1] = -j01 / det; gx[2] = -j10 / det; gy[2] = j00 / det;
    gx[0] = -(gx[1] + gx[2]); gy[0] = -(gy[1] + gy[2]);
    double e0 = hypot(x[1] - x[0], y[1] - y[0]), e1 = hypot(x[2] - x[0], y[2] - y[0]),
           e2 = hypot(x[2] - x[1], y[2] - y[1]);
    double hK = fmax(e0, fmax(e1, e2));
    double tau = 0.5 * P->supg_stab * hK;
    double px = 0, py = 0, qx = 0;
    for (int i = 0; i < 3; ++i) { px += pe[i] * gx[i]; py += pe[i] * gy[i]; qx += qe[i] * gx[i]; }
    double Fe[6] = {0, 0, 0, 0, 0, 0}, Je[36];
    memset(Je, 0, sizeof(Je));
    for (int k = 0; k < 6; ++k) {
      double phi[3] = {1.0 - X4[k] - Y4[k], X4[k], Y4[k]};
      double xq = 0, yq = 0, p = 0, q = 0, pold = 0;
      for (int i = 0; i < 3; ++i) {
        xq += phi[i] * x[i]; yq += phi[i] * y[i]; p += phi[i] * pe[i]; q += phi[i] * qe[i];
        pold += phi[i] * po[i];
      }
      double t = fourpi * yq * xq, a = W4[k] * adet * t * t, x2 = xq * xq;
      double G, Gp, Gq;
      eval_G(P, x2, p, q, &G, &Gp, &Gq);
      double f2 = 0, f3 = 0;
      if (fq) { f2 = fq[(c * 6 + k) * 2]; f3 = fq[(c * 6 + k) * 2 + 1]; }
      double resid = P->q_react * q + P->q_adv * qx + x2 * (P->q_x2p * p + P->q_x2q * q) - f3;
      if (P->q_divr != 0.0) resid -= P->q_divr * (2.0 / xq) * q;
      double dpt = (p - pold) * P->inv_dt;
      for (int i = 0; i < 3; ++i) {
        double wt = phi[i] + tau * gx[i];
        Fe[i] += a * (dpt * phi[i] + P->D2 * (px + G) * gx[i] + P->D1 * py * gy[i] - f2 * phi[i]);
        Fe[3 + i] += a * (resid * wt - P->q_ibp * q * gx[i]);
        if (!vals) continue;
        for (int j = 0; j < 3; ++j) {
          Je[6 * i + j] += a * (P->inv_dt * phi[j] * phi[i] + P->D2 * (gx[j] + Gp * phi[j]) * gx[i] +
                                P->D1 * gy[j] * gy[i]);
          Je[6 * i + 3 + j] += a * (P->D2 * Gq * phi[j] * gx[i]);
          Je[6 * (3 + i) + j] += a * (x2 * P->q_x2p * phi[j] * wt);
          double dq = P->q_react * phi[j] + P->q_adv * gx[j] + x2 * P->q_x2q * phi[j];
          if (P->q_divr != 0.0) dq -= P->q_divr * (2.0 / xq) * phi[j];
          Je[6 * (3 + i) + 3 + j] += a * (dq * wt - P->q_ibp * phi[j] * gx[i]);
        }
      }
    }
    for (int i = 0; i < 3; ++i) {
      b[2 * cv[i]] += Fe[i];
      b[2 * cv[i] + 1] += Fe[3 + i];
    }
    if (vals) {
      const int64_t* s = slots + 36 * c;
      for (int k = 0; k < 36; ++k) vals[s[k]] += Je[k];
    }
  }
}
