"""Function spaces, functions and Dirichlet conditions (host-side mirror of the FEniCS names used at
advection_diffusion_convergence.py:147-183 and …MixedPrimal_convergence.py:93-117).

Only what the P1 hot path needs: scalar CG1 and the mixed CG1 x CG1 space with DOLFIN's interleaved
dof layout dof = 2*vertex + component (reorder_dofs_serial=False).
"""
from __future__ import annotations

import torch

from .mesh import facet_vertices


class FiniteElement:
    def __init__(self, family, cell=None, degree=1):
        if family not in ("CG", "Lagrange", "P") or degree != 1:
            raise NotImplementedError(
                f"element {family}{degree}: only CG1 is on the B200 hot path (RT/DG/CG2+ are next rows)")
        self.family, self.cell, self.degree = "CG", cell, degree


class MixedElement:
    def __init__(self, elements):
        self.elements = list(elements)
        if len(self.elements) != 2:
            raise NotImplementedError("only the CG1 x CG1 (p,q) space is supported")


class FunctionSpace:
    def __init__(self, mesh, family_or_element, degree=None):
        self.mesh = mesh
        if isinstance(family_or_element, MixedElement):
            self.bs = 2
        else:
            if isinstance(family_or_element, str):
                FiniteElement(family_or_element, None, degree)
            self.bs = 1
        self.component = None
        self.parent = None

    def dim(self):
        return self.bs * self.mesh.num_vertices()

    def sub(self, i):
        if self.bs == 1:
            raise ValueError("scalar space has no sub-spaces")
        s = FunctionSpace.__new__(FunctionSpace)
        s.mesh, s.bs, s.component, s.parent = self.mesh, 1, int(i), self
        return s

    def collapse(self):
        return FunctionSpace(self.mesh, "CG", 1)


class Constant:
    def __init__(self, v):
        self.value = v

    def __call__(self, x, y):
        return torch.full_like(x, float(self.value))


class Function:
    """Nodal vector on the mesh device; zero-initialised like dolfin.Function."""

    def __init__(self, V):
        self.space = V
        self._vec = torch.zeros(V.dim(), dtype=torch.float64, device=V.mesh.device)

    def vector(self):
        return self._vec

    def assign(self, other):
        self._vec.copy_(other._vec if isinstance(other, Function) else other)

    def split(self, deepcopy=False):
        if self.space.bs == 1:
            return (self._vec,)
        return self._vec[0::2], self._vec[1::2]

    def interpolate(self, f):
        """P1 nodal interpolation of f(x,y) (interpolate(), advection_diffusion_convergence.py:166)."""
        X = self.space.mesh.coords
        self._vec.copy_(_evaluate(f, X[:, 0], X[:, 1]))
        return self


def interpolate(f, V):
    return Function(V).interpolate(f)


def _evaluate(g, x, y):
    if isinstance(g, (int, float)):
        return torch.full_like(x, float(g))
    out = g(x, y)
    if not torch.is_tensor(out):
        out = torch.full_like(x, float(out))
    return out.to(torch.float64)


class DirichletBC:
    """DirichletBC(V|V.sub(i), g, bdry, id) or DirichletBC(V, g, 'on_boundary') -- topological
    method: the closure dofs of every marked facet, g evaluated at the vertex."""

    def __init__(self, V, g, bdry, marker=None):
        self.space = V
        mesh = V.mesh
        if isinstance(bdry, str):
            verts = facet_vertices(mesh, None, "on_boundary")
        else:
            verts = facet_vertices(mesh, bdry, marker)
        X = mesh.coords[verts]
        self.values = _evaluate(g, X[:, 0], X[:, 1])
        if V.parent is not None:
            self.dofs = 2 * verts + V.component
        else:
            self.dofs = verts if V.bs == 1 else None
            if self.dofs is None:
                raise ValueError("DirichletBC on the mixed space needs .sub(i)")

    @staticmethod
    def merge(bcs, n):
        """BCs apply in list order, the last one wins on shared dofs.  Returns sorted
        unique dofs (int32) and their values."""
        if not bcs:
            dev = None
            return (torch.zeros(0, dtype=torch.int32), torch.zeros(0, dtype=torch.float64))
        dev = bcs[0].dofs.device
        val = torch.zeros(n, dtype=torch.float64, device=dev)
        has = torch.zeros(n, dtype=torch.bool, device=dev)
        for bc in bcs:
            val[bc.dofs] = bc.values
            has[bc.dofs] = True
        dofs = torch.nonzero(has).flatten()
        return dofs.to(torch.int32).contiguous(), val[dofs].contiguous()
