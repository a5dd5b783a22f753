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
