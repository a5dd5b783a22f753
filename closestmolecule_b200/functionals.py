"""Scalar functionals and flux post-processing the drivers report, computed by the
CUDA kernels of csrc/cm_functional.cu through the C ABI (cm_functional_error_norms,
cm_project_flux_dg0, cm_functional_boundary_flux):

* error norms via assemble((u_ex-u_h)**2*dx) (SphericalAdvectionDiffusionReaction_convergence.py:117-118,
  …MixedPrimal_convergence.py:148-150);
* the DG0 flux project(dMatrix*grad(p_h) + D2*r2*r2*p_h/q_h*p_h*r2_vec, VectorFunctionSpace(mesh,"DG",0))
  (convergence/advection_diffusion_convergence.py:222-223);
* the boundary flux V1*assemble(dot(flux,n)*4*pi*r1**2*ds(bottom)) with its r1 / r2 parts
  (advection_diffusion_mixed.py:261-263).

Exact data is evaluated at the kernel's quadrature points with torch on the mesh device (setup
plumbing, like the forcing of cm_assemble_load); there is no CPU fallback.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _capi
from .quadrature import triangle_rule


def _rule_host(degree):
    pts, wts = triangle_rule(degree)
    r = np.ascontiguousarray(np.concatenate([pts, wts[:, None]], 1), dtype=np.float64)   # (xi, eta, w)
    return r, r.ctypes.data, pts


def _quad_points(mesh, pts):
    dev = mesh.device
    phi = torch.tensor(np.stack([1.0 - pts[:, 0] - pts[:, 1], pts[:, 0], pts[:, 1]], 1), device=dev)
    X = mesh.coords[mesh.cells.long()]
    P = torch.einsum("qi,cid->cqd", phi, X)
    return P[:, :, 0], P[:, :, 1]


def error_norms(mesh, uh, exact, grad_exact, degree=4, weighted=False, x_derivative_only=False):
    """(L2 error, H1-seminorm error) of the nodal P1 function uh against exact(x,y) /
    grad_exact(x,y) -> (gx, gy); weighted: with W = (4 pi x y)^2 (…MixedPrimal_convergence.py:148-150).
    uh may be a strided view (a component of the interleaved P-Q vector)."""
    _capi.require_cuda(mesh.coords)
    lib = _capi.lib()
    rule, rule_ptr, pts = _rule_host(degree)
    x, y = _quad_points(mesh, pts)
    exq = exact(x, y).to(torch.float64).contiguous()
    gx, gy = grad_exact(x, y)
    grq = torch.stack([gx, gy], -1).to(torch.float64).contiguous()
    if uh.dim() == 1 and uh.stride(0) in (1, 2):
        stride, base = uh.stride(0), uh
    else:
        stride, base = 1, uh.contiguous()
    work = torch.empty(int(lib.cm_functional_workspace_doubles()), dtype=torch.float64, device=mesh.device)
    _capi.check(lib.cm_functional_error_norms(
        rule.shape[0], rule_ptr, int(bool(weighted)), int(bool(x_derivative_only)), mesh.num_cells(),
        _capi.ptr(mesh.cells), _capi.ptr(mesh.coords), base.data_ptr(), stride, _capi.ptr(exq),
        _capi.ptr(grq), _capi.ptr(work), _capi.stream_ptr()))
    s = work[:2].cpu()
    return math.sqrt(float(s[0])), math.sqrt(float(s[1]))


def project_flux_dg0(mesh, u, D1, D2, degree=5):
    """[C,2] cell-wise flux dMatrix*grad(p_h) + D2*r2^2*p_h^2/q_h*r2_vec projected onto DG0 vectors
    (convergence/advection_diffusion_convergence.py:222-223; UFL estimates quadrature degree 5 for the
    quotient term).  u: interleaved P-Q vector [2V]."""
    _capi.require_cuda(mesh.coords)
    lib = _capi.lib()
    rule, rule_ptr, _ = _rule_host(degree)
    flux = torch.empty((mesh.num_cells(), 2), dtype=torch.float64, device=mesh.device)
    _capi.check(lib.cm_project_flux_dg0(rule.shape[0], rule_ptr, float(D1), float(D2), mesh.num_cells(),
                                        _capi.ptr(mesh.cells), _capi.ptr(mesh.coords), _capi.ptr(u),
                                        _capi.ptr(flux), _capi.stream_ptr()))
    return flux


def boundary_flux(mesh, flux, bdry, marker):
    """(r2 part, r1 part, total) of assemble(dot(flux,n)*4*pi*r1**2*ds(marker))
    (advection_diffusion_mixed.py:261-263; multiply by V1 there) for a DG0 flux [C,2]."""
    _capi.require_cuda(mesh.coords)
    lib = _capi.lib()
    fc = mesh.facets()
    sel = torch.nonzero((fc.cell1 < 0) & (bdry.values == int(marker))).flatten()
    efverts = fc.verts[sel].contiguous()
    efcell = fc.cell0[sel].contiguous()
    efopp = mesh.cells.long()[fc.cell0[sel].long(), fc.loc0[sel].long()].to(torch.int32).contiguous()
    out = torch.zeros(3, dtype=torch.float64, device=mesh.device)
    _capi.check(lib.cm_functional_boundary_flux(sel.numel(), _capi.ptr(efverts), _capi.ptr(efopp),
                                                _capi.ptr(efcell), _capi.ptr(mesh.coords),
                                                _capi.ptr(flux), _capi.ptr(out), _capi.stream_ptr()))
    o = out.cpu()
    return float(o[0]), float(o[1]), float(o[2])
