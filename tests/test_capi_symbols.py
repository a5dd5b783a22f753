"""The C-ABI library loads on a CPU-only box and exports every symbol include/closestmol_b200.h
declares (no compute call is made here)."""
import ctypes
import os

from closestmolecule_b200 import _capi, build


def test_library_builds_and_exports_declared_symbols():
    build.build()
    assert os.path.exists(_capi.LIB_PATH)
    L = ctypes.CDLL(_capi.LIB_PATH)
    names = _capi.declared_symbols()
    assert "cm_assemble_pq_cells" in names and "cm_apply_dirichlet" in names
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    lib = _capi.lib()
    assert lib.cm_version() >= 100
    for n in names:
        assert n in _capi._PROTOS, f"{n} has no ctypes prototype"


def test_no_cpu_fallback():
    import pytest
    import torch
    import closestmolecule_b200 as cm
    from closestmolecule_b200.assembler import PQAssembler
    if torch.cuda.is_available():
        pytest.skip("GPU box")
    mesh = cm.RectangleMesh(cm.Point(0, 0), cm.Point(1, 1), 2, 2, device="cpu")
    V = cm.FunctionSpace(mesh, cm.MixedElement([cm.FiniteElement("CG", "triangle", 1)] * 2))
    with pytest.raises(_capi.CMError):
        PQAssembler(cm.PQPrimalForm.flat_boundary(V, 1.0, 2.0, 1.5))
    # the functionals and the mixed RT1 x DG0 x CG1 path refuse CPU tensors as well
    from closestmolecule_b200 import functionals
    from closestmolecule_b200.mixed import PaperMixedSystem
    u = torch.zeros(2 * mesh.num_vertices(), dtype=torch.float64)
    with pytest.raises(_capi.CMError):
        functionals.project_flux_dg0(mesh, u, 2.0, 1.5)
    with pytest.raises(_capi.CMError):
        functionals.error_norms(mesh, u[0::2], lambda x, y: x, lambda x, y: (x, y))
    with pytest.raises(_capi.CMError):
        PaperMixedSystem(mesh, cm.MeshFunction("size_t", mesh, 1, 0))


def test_argument_validation_fails_loudly_without_touching_the_gpu():
    """Bad arguments are rejected before any CUDA call (return < 0 + cm_last_error), so this runs on the CPU."""
    from closestmolecule_b200 import _capi
    L = _capi.lib()
    assert L.cm_blu_workspace_bytes(3, 10, 4) < 0 and b"bad argument" in L.cm_last_error()
    assert L.cm_blu_workspace_bytes(2, 1000, 40) > 4 * (2 * 40) ** 2 * 8 * 25      # 4 nb^2 doubles per block
    assert L.cm_blu_factor(None, 2, 10, 4, None, None, None, None, None, None, 0, None, None) == -1
    assert b"null context" in L.cm_last_error()
    assert L.cm_blu_solve(None, 2, 10, 4, None, None, None, None, 0, None) == -1
    assert L.cm_functional_error_norms(6, None, 0, 0, 10, None, None, None, 1, None, None, None, None) == -1
    assert L.cm_mixed_m0_cells(12, None, 1.0, 1.5, 3e-16, 0, None, None, None, None, 5, None, None, None, None,
                               None, None) == -1
    assert L.cm_csr_gather(-1, None, None, None, None, 0, None) == -1
    assert L.cm_functional_workspace_doubles() > 2 and L.cm_mixed_workspace_doubles() > 3
