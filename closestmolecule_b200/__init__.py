"""closestmolecule_b200 -- B200-native drop-in for the FEM hot path of ruizbaier/closestMolecule:
Newton Jacobian/residual assembly of the P-Q system and the sparse solve of every Newton step.

The names mirror the FEniCS API surface the reference scripts use; forms are named
objects (forms.py) instead of UFL expressions.  Compute runs in hand-written sm_100a CUDA behind the
C ABI of include/closestmol_b200.h; there is no CPU fallback.
"""
from .mesh import (CompiledSubDomain, MappedRectangleMesh, Mesh, MeshFunction, Point, RectangleMesh,
                   UnitSquareMesh, flat_boundary_markers, structured_boundary_markers, DOLFIN_EPS)
from .space import (Constant, DirichletBC, FiniteElement, Function, FunctionSpace, MixedElement,
                    interpolate)
from .meshio import read_facet_function, read_mesh, write_facet_function, write_mesh
from .forms import EF_ADV, EF_DATA, EF_PENALTY, PQPrimalForm, ScalarADRForm

__all__ = [n for n in dir() if not n.startswith("_")]
from .solver import (LinearVariationalSolver, NonlinearVariationalProblem,  # noqa: E402
                     NonlinearVariationalSolver, solve)
__all__ = [n for n in dir() if not n.startswith("_")]
