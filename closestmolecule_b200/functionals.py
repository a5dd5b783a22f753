"""Scalar functionals the convergence drivers report: error norms via
assemble((u_ex-u_h)**2*dx) (SphericalAdvectionDiffusionReaction_convergence.py:117-118,
…MixedPrimal_convergence.py:148-150) and mesh.hmax().  Vectorised torch on the mesh device
(cell-parallel reduction; setup-level cost, not on the Newton hot loop)."""
from __future__ import annotations

import math

import numpy as np
import torch

from .quadrature import triangle_rule


def _cell_quadrature(mesh, degree):
    pts, wts = triangle_rule(degree)
    dev = mesh.device
    phi = torch.tensor(np.stack([1 - pts[:, 0] - pts[:, 1], pts[:, 0], pts[:, 1]], 1), device=dev)
    w = torch.tensor(wts, device=dev)
    X = mesh.coords[mesh.cells.long()]
    x0, x1, x2 = X[:, 0], X[:, 1], X[:, 2]
    j00, j01 = x1[:, 0] - x0[:, 0], x2[:, 0] - x0[:, 0]
    j10, j11 = x1[:, 1] - x0[:, 1], x2[:, 1] - x0[:, 1]
    det = j00 * j11 - j01 * j10
    g1 = torch.stack([j11, -j01], 1) / det[:, None]
    g2 = torch.stack([-j10, j00], 1) / det[:, None]
    grads = torch.stack([-(g1 + g2), g1, g2], 1)
    P = torch.einsum("qi,cid->cqd", phi, X)
    return phi, w[None, :] * det.abs()[:, None], grads, P


def error_norms(mesh, uh, exact, grad_exact, degree=4, weighted=False, x_derivative_only=False):
    """(L2 error, H1-seminorm error) of the nodal P1 function uh against exact(x,y) /
    grad_exact(x,y) -> (gx, gy); weighted: with W = (4 pi x y)^2 (…MixedPrimal_convergence.py:148-150)."""
    phi, a, grads, P = _cell_quadrature(mesh, degree)
    x, y = P[:, :, 0], P[:, :, 1]
    if weighted:
        a = a * (4.0 * math.pi * x * y) ** 2
    ue = uh[mesh.cells.long()]
    uq = ue @ phi.T
    ux = (ue * grads[:, :, 0]).sum(1)[:, None]
    uy = (ue * grads[:, :, 1]).sum(1)[:, None]
    gx, gy = grad_exact(x, y)
    e0 = torch.sqrt((a * (exact(x, y) - uq) ** 2).sum())
    d = (gx - ux) ** 2 if x_derivative_only else (gx - ux) ** 2 + (gy - uy) ** 2
    e1 = torch.sqrt((a * d).sum())
    return float(e0), float(e1)
