"""Run this WITH FEniCS 2019.1 (not available in the build image) to dump golden arrays for the P-Q
primal P1 path that the reference never records: mesh, dof maps, the assembled Jacobian
in PETSc AIJ CSR order, the residual and the Newton solution.  Compare with closestmolecule_b200
through tests/ by loading the .npz (entries within 1e-12 relative, solutions within 1e-8).

    python3 tools/dump_fenics_golden.py 16 16 golden_pq_16.npz
"""
import sys

import numpy as np


def main(nx, ny, out):
    from fenics import (Constant, DirichletBC, FiniteElement, Function, FunctionSpace, MixedElement, Point,
                        RectangleMesh, SpatialCoordinate, TestFunctions, TrialFunction, as_backend_type,
                        as_tensor, assemble, conditional, derivative, dot, grad, gt, Dx, dx, parameters, pi,
                        split, vertex_to_dof_map)
    parameters["form_compiler"]["representation"] = "uflacs"
    parameters["form_compiler"]["quadrature_degree"] = 4
    parameters["reorder_dofs_serial"] = False          # dof = 2*vertex + component
    mesh = RectangleMesh(Point(0, 0.1), Point(5, 5), nx, ny)
    r2, r1 = SpatialCoordinate(mesh)
    weight = (4 * pi * r1 * r2) ** 2
    P = FiniteElement('CG', mesh.ufl_cell(), 1)
    W = FunctionSpace(mesh, MixedElement([P, P]))
    v1, v2 = TestFunctions(W)
    u = Function(W)
    du = TrialFunction(W)
    p, q = split(u)
    rng = np.random.default_rng(0)
    u.vector().set_local(np.stack([0.3 + rng.random(mesh.num_vertices()),
                                   1.0 + rng.random(mesh.num_vertices())], 1).ravel())
    c, D1, D2 = 1.0, 2.0, 1.5
    dM = as_tensor([[D2, 0], [0, D1]])
    G = conditional(gt(abs((1 / (4 * np.pi) * p) - q), 0.000001), (p / q) * p * r2 ** 2, 4 * np.pi * c * p * r2 ** 2)
    FF = dot(dM * (grad(p) + G * Constant((1, 0))), grad(v1)) * weight * dx \
        + Dx(q, 0) * v2 * weight * dx + r2 ** 2 * p * v2 * weight * dx
    A = as_backend_type(assemble(derivative(FF, u, du))).mat()
    indptr, indices, data = A.getValuesCSR()
    b = assemble(FF).get_local()
    np.savez(out, coords=mesh.coordinates(), cells=mesh.cells(), indptr=indptr, indices=indices, data=data,
             b=b, u=u.vector().get_local(), v2d=vertex_to_dof_map(W))


if __name__ == "__main__":
    main(int(sys.argv[1]), int(sys.argv[2]), sys.argv[3])
