"""Named weak forms of the reference scripts (there is no UFL compiler here: each form the scripts
declare is a parameterised object whose element arithmetic lives in the CUDA library).

PQPrimalForm   advection_diffusion_convergence.py:196-204 (+ the SUPG / C0IP q-equation variants of
               …MixedPrimal_convergence_SUPG.py:118-129, …_C0IP.py:141-147,
               test_for_paper_convergence.py:140-153 transplanted to the CG1 x CG1 space).
ScalarADRForm  SphericalAdvectionDiffusionReaction_convergence.py:92-100 and
               SphericalLaplace_convergence.py:128.
"""
from __future__ import annotations

import math

import torch

from . import _capi
from .mesh import DOLFIN_EPS

G_NONE, G_POWER, G_GSTAR = 0, 1, 2
EF_ADV, EF_DATA, EF_PENALTY = 1, 2, 4


class PQPrimalForm:
    """Residual FF and Jacobian derivative(FF,u,du) of the coupled P-Q system on CG1 x CG1.

    p-equation: (p-p_old)/dt*v1*W + D2*(dp/dx + G)*dv1/dx*W + D1*dp/dy*dv1/dy*W - f2*v1*W
    q-equation: [q_react*q + q_adv*dq/dx + x^2*(q_x2p*p + q_x2q*q) - q_divr*(2/x)*q - f3]
                *(w + supg_stab*hK/2*dw/dx)*W - q_ibp*q*dw/dx*W  (+ C0IP facet terms)
    """

    def __init__(self, V, D1, D2, *, G=("gstar_conv", 1.0), dt=None, p_old=None, forcing=None,
                 q_react=0.0, q_adv=1.0, q_x2p=1.0, q_x2q=0.0, q_divr=0.0, q_ibp=0.0,
                 supg_stab=0.0, c0ip_gamma=0.0, c0ip_h="cell", pen_stab2=0.0, ext_terms=(),
                 q_data=None, r2vec=(1.0, 0.0), degree=4):
        if V.bs != 2:
            raise ValueError("PQPrimalForm needs the mixed CG1 x CG1 space")
        self.space = V
        p = _capi.PQParamsC()
        p.D1, p.D2 = float(D1), float(D2)
        p.inv_dt = 0.0 if dt is None else 1.0 / float(dt)
        kind = G[0]
        p.geps = DOLFIN_EPS
        p.gcoef = 1.0
        if kind == "none":
            p.gkind = G_NONE
        elif kind == "power":            # gcoef*x^2*p^2/(q+DOLFIN_EPS)
            p.gkind, p.gcoef = G_POWER, float(G[1])
        elif kind == "gstar_conv":       # advection_diffusion_convergence.py:21-27
            c = float(G[1])
            p.gkind, p.gs_thr, p.gs_cond_scale, p.gs_lin = G_GSTAR, 1e-6, 1.0 / (4.0 * math.pi), 4.0 * math.pi * c
        elif kind == "gstar_prod":       # advection_diffusion_mixed.py:17-23
            c = float(G[1])
            p.gkind, p.gs_thr, p.gs_cond_scale, p.gs_lin = G_GSTAR, 1e-3, 1.0 / (4.0 * c * math.pi), 4.0 * math.pi * c
        else:
            raise ValueError(f"unknown G kind {kind}")
        p.q_react, p.q_adv, p.q_x2p, p.q_x2q = float(q_react), float(q_adv), float(q_x2p), float(q_x2q)
        p.q_divr, p.q_ibp, p.supg_stab = float(q_divr), float(q_ibp), float(supg_stab)
        p.c0ip_gamma, p.pen_stab2 = float(c0ip_gamma), float(pen_stab2)
        p.c0ip_h = {"cell": 0, "facet": 1}[c0ip_h]
        p.r2x, p.r2y = float(r2vec[0]), float(r2vec[1])
        p.degree = int(degree)
        self.params = p
        self.p_old = p_old              # Function on the collapsed p space (transient only)
        self.forcing = forcing          # f(x,y) -> (f2, f3) torch tensors, or None
        self.ext_terms = list(ext_terms)  # [(MeshFunction, marker | 'on_boundary', kind bits)]
        self.q_data = q_data            # q_ex(x,y) for EF_DATA / EF_PENALTY
        if dt is not None and p_old is None:
            raise ValueError("transient form needs p_old")

    @property
    def with_interior_facets(self):
        return self.params.c0ip_gamma != 0.0

    # ---- the variants the reference scripts declare -------------------------------------------
    @classmethod
    def flat_boundary(cls, V, concentration, D1, D2, dt=None, p_old=None, production_gstar=False):
        """advection_diffusion_convergence.py:196-204 (steady when dt is None)."""
        kind = "gstar_prod" if production_gstar else "gstar_conv"
        return cls(V, D1, D2, G=(kind, concentration), dt=dt, p_old=p_old)

    @classmethod
    def manufactured(cls, V, variant="galerkin", D1=1e-3, D2=1.5e-3, stab=None, sign=-1.0,
                     c0ip_h="cell", degree=4, bdry=None, left=None, ext_terms=(), pen_stab2=0.0, r2dir=1.0):
        """P1 x P1 restatement of the manufactured MixedPrimal scripts (p_ex, q_ex of :48-49).  r2dir = -1 (C0IP
        only): r2vec = (-1,0) in the q-equation as in …_C0IP_otherBCs.py:58 (G keeps its direction, :46-47)."""
        from .manufactured import forcing_torch, q_exact
        if variant == "galerkin":     # …MixedPrimal_convergence.py:120-123
            gc = D2 * 4.0 * math.pi
            f = forcing_torch(D1, D2, gc, q_react=1.0, x2_on_q=False)
            return cls(V, D1, D2, G=("power", gc), forcing=f, q_react=1.0, degree=degree)
        if variant == "supg":         # …MixedPrimal_convergence_SUPG.py:112,118-129 (x^2*q as written)
            gc = D2 * 4.0 * math.pi
            f = forcing_torch(D1, D2, gc, q_react=1.0, x2_on_q=True)
            return cls(V, D1, D2, G=("power", gc), forcing=f, q_react=1.0, q_x2p=0.0, q_x2q=1.0,
                       supg_stab=1.0 if stab is None else stab, degree=degree)
        if variant == "c0ip":         # …MixedPrimal_convergence_C0IP.py:141-147 / test_for_paper:140-153
            f = forcing_torch(D1, D2, 1.0, q_react=0.0, x2_on_q=False, r2dir=r2dir)
            gamma = sign * (7.5e-4 if stab is None else stab) / (1.0 ** 3.5)   # k_q = 1 for CG1
            return cls(V, D1, D2, G=("power", 1.0), forcing=f, q_adv=0.0, q_divr=float(r2dir), q_ibp=float(r2dir),
                       c0ip_gamma=gamma, c0ip_h=c0ip_h, pen_stab2=pen_stab2, ext_terms=ext_terms,
                       q_data=q_exact, r2vec=(float(r2dir), 0.0), degree=degree)
        raise ValueError(variant)


class ScalarADRForm:
    """F = u v + D2 u_x v_x + D1 u_y v_y - (2 D2/x) u_x v - (2 D1/y) u_y v + g(u) v_x - (2/x) g(u) v - f v,
    g = gcoef*x^2*u^2."""

    def __init__(self, V, D1, D2, gcoef=0.0, forcing=None, degree=4, load_degree=None):
        if V.bs != 1:
            raise ValueError("ScalarADRForm needs a scalar CG1 space")
        self.space = V
        p = _capi.ScalarParamsC()
        p.D1, p.D2, p.gcoef, p.degree = float(D1), float(D2), float(gcoef), int(degree)
        self.params = p
        self.forcing = forcing
        self.load_degree = degree if load_degree is None else load_degree

    @classmethod
    def nonlinear_adr(cls, V, D1=1e-3, D2=1e-4):
        """SphericalAdvectionDiffusionReaction_convergence.py:45-46,92-100 (density = 1)."""
        from .manufactured import scalar_forcing_fb
        return cls(V, D1, D2, gcoef=D2 * 4.0 * math.pi, forcing=scalar_forcing_fb(D1, D2), degree=4)

    @classmethod
    def laplace(cls, V, D=1e-4):
        """SphericalLaplace_convergence.py:90,128: bilinear form at UFL's estimated degree 2; the rhs
        (estimated degree 8 in FEniCS) is integrated with the degree-6 rule here."""
        from .manufactured import scalar_forcing_fa
        return cls(V, D, D, gcoef=0.0, forcing=scalar_forcing_fa(D), degree=2, load_degree=6)
