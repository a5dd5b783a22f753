"""Manufactured data of the convergence scripts as torch callables.

p_ex = sin(pi x) sin(pi y), q_ex = 1.5 + cos(pi x y)  (…MixedPrimal_convergence.py:48-49),
u_ex = sin(pi x) sin(pi y)                             (SphericalLaplace_convergence.py:86).
The scripts build the forcing symbolically from degree-5/7 Expression interpolants
(…MixedPrimal_convergence.py:109-113); here the same formulas are differentiated with sympy and
evaluated analytically (the difference is the interpolation error of B.6, invisible for h < 0.02).
"""
from __future__ import annotations

import math

import torch

from .mesh import DOLFIN_EPS


def _torchify(expr, syms):
    import sympy as sm
    mods = [{"sin": torch.sin, "cos": torch.cos, "exp": torch.exp, "sqrt": torch.sqrt, "pi": math.pi}, "math"]
    return sm.lambdify(syms, expr, modules=mods)


def p_exact(x, y):
    return torch.sin(math.pi * x) * torch.sin(math.pi * y)


def q_exact(x, y):
    return 1.5 + torch.cos(math.pi * x * y)


def forcing_torch(D1, D2, gcoef, q_react=1.0, x2_on_q=False, r2dir=1.0):
    """(f2_ex, f3_ex) of …MixedPrimal_convergence.py:111-113: f2 = div_rad(D*sig_ex),
    sig_ex = -grad(p_ex) - G(p_ex,q_ex) (1,0), f3 = q_react*q_ex + grad(q_ex).r2vec + x^2*(p_ex|q_ex) with
    r2vec = (r2dir, 0) (-1: …_C0IP_otherBCs.py:58,117)."""
    import sympy as sm
    xs, ys = sm.symbols("x y", positive=True)
    pe = sm.sin(sm.pi * xs) * sm.sin(sm.pi * ys)
    qe = sm.Rational(3, 2) + sm.cos(sm.pi * xs * ys)
    G = gcoef * xs ** 2 * pe ** 2 / (qe + DOLFIN_EPS)
    sx = -sm.diff(pe, xs) - G
    sy = -sm.diff(pe, ys)
    f2 = sm.diff(xs ** 2 * D2 * sx, xs) / xs ** 2 + sm.diff(ys ** 2 * D1 * sy, ys) / ys ** 2
    f3 = q_react * qe + r2dir * sm.diff(qe, xs) + xs ** 2 * (qe if x2_on_q else pe)
    f2t, f3t = _torchify(f2, (xs, ys)), _torchify(f3, (xs, ys))
    return lambda x, y: (f2t(x, y), f3t(x, y))


def scalar_forcing_fb(D1, D2):
    """f_ex of SphericalAdvectionDiffusionReaction_convergence.py:85 (density = 1)."""
    pi = math.pi

    def f(x, y):
        u = torch.sin(pi * x) * torch.sin(pi * y)
        ux = pi * torch.cos(pi * x) * torch.sin(pi * y)
        uy = pi * torch.sin(pi * x) * torch.cos(pi * y)
        lap = D2 * (-pi ** 2 * u) + D2 * 2.0 / x * ux + D1 * (-pi ** 2 * u) + D1 * 2.0 / y * uy
        g = 4.0 * pi * x ** 2 * u ** 2
        gx = 4.0 * pi * (2.0 * x * u ** 2 + x ** 2 * 2.0 * u * ux)
        return u - lap - (D2 * gx + D2 * 2.0 / x * g)
    return f


def scalar_forcing_fa(D):
    """f_ex of SphericalLaplace_convergence.py:121."""
    pi = math.pi

    def f(x, y):
        u = torch.sin(pi * x) * torch.sin(pi * y)
        ux = pi * torch.cos(pi * x) * torch.sin(pi * y)
        uy = pi * torch.sin(pi * x) * torch.cos(pi * y)
        return u - D * (-2.0 * pi ** 2 * u + 2.0 / x * ux + 2.0 / y * uy)
    return f


def u_exact(x, y):
    return torch.sin(math.pi * x) * torch.sin(math.pi * y)
