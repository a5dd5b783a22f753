"""ctypes binding of libclosestmol_b200.so (include/closestmol_b200.h).

There is no CPU fallback: every compute entry point raises if the library is missing or the call
fails.  torch tensors are only used as owners of device memory; the C ABI sees raw pointers.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libclosestmol_b200.so")

_lib = None


class CMError(RuntimeError):
    pass


class PQParamsC(C.Structure):
    """struct cm_pq_params"""
    _fields_ = [(n, C.c_double) for n in (
        "D1", "D2", "inv_dt", "gcoef", "geps", "gs_thr", "gs_cond_scale", "gs_lin", "q_react",
        "q_adv", "q_x2p", "q_x2q", "q_divr", "q_ibp", "supg_stab", "c0ip_gamma", "pen_stab2", "r2x",
        "r2y")] + [(n, C.c_int32) for n in ("gkind", "degree", "c0ip_h", "reserved")]


class ScalarParamsC(C.Structure):
    """struct cm_scalar_params"""
    _fields_ = [("D1", C.c_double), ("D2", C.c_double), ("gcoef", C.c_double),
                ("degree", C.c_int32), ("reserved", C.c_int32)]


class DistC(C.Structure):
    """struct cm_dist"""
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("peer_base", C.c_void_p * 8),
                ("halo_seq", C.c_uint64), ("coll_seq", C.c_uint64), ("row_begin", C.c_int32 * 9),
                ("reserved", C.c_int32)]


class KrylovParamsC(C.Structure):
    """struct cm_krylov_params"""
    _fields_ = [("rtol", C.c_double), ("atol", C.c_double), ("maxit", C.c_int32),
                ("restart", C.c_int32), ("method", C.c_int32), ("precond", C.c_int32),
                ("check_every", C.c_int32), ("verbose", C.c_int32)]


_P = C.c_void_p
_I = C.c_int32
_D = C.c_double

_PROTOS = {
    "cm_version": ([], C.c_int),
    "cm_last_error": ([], C.c_char_p),
    "cm_device_available": ([], C.c_int),
    "cm_profile_enable": ([_I], C.c_int),
    "cm_profile_report": ([_I, C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_double)], C.c_int),
    "cm_profile_reset": ([], C.c_int),
    "cm_launch_count": ([], C.c_int64),
    "cm_algorithmic_bytes": ([], C.c_double),
    "cm_assemble_pq_cells": ([C.POINTER(PQParamsC), _I, _I] + [_P] * 7 + [_I] + [_P] * 7 + [_I, _P], C.c_int),
    "cm_assemble_pq_cells_structured": ([C.POINTER(PQParamsC), _I, _I] + [_P] * 7 + [_P], C.c_int),
    "cm_assemble_scalar_cells": ([C.POINTER(ScalarParamsC), _I, _I] + [_P] * 6 + [_I] + [_P] * 4 + [_P], C.c_int),
    "cm_assemble_load": ([_I, _I, _D, _I, _I, _I] + [_P] * 6 + [_P], C.c_int),
    "cm_assemble_c0ip_facets": ([C.POINTER(PQParamsC), _I, _I] + [_P] * 9 + [_P], C.c_int),
    "cm_assemble_exterior_facets": ([C.POINTER(PQParamsC), _I, _I] + [_P] * 11 + [_P], C.c_int),
    "cm_apply_dirichlet": ([_I, _I] + [_P] * 7 + [_P], C.c_int),
    "cm_spmv": ([_I, _I] + [_P] * 5 + [_P], C.c_int),
    "cm_pqs_workspace_bytes": ([_I, _I, _I], C.c_int64),
    "cm_pqs_setup": ([_I, _I, _I, _P, _P, _P, _P, C.c_int64, _P], C.c_int),
    "cm_pqs_solve": ([_I, _I, _I, _P, _P, _D, _D, _I, _P, C.c_int64, C.POINTER(C.c_int32),
                      C.POINTER(C.c_double), _P], C.c_int),
    "cm_pqs_create": ([_I, _I, _I, _P, C.c_int64, C.POINTER(DistC)], C.c_void_p),
    "cm_pqs_destroy": ([_P], C.c_int),
    "cm_pqs_ctx_setup": ([_P, _P, _P, _P, _P], C.c_int),
    "cm_pqs_ctx_solve": ([_P, _P, _P, _D, _D, _I, C.POINTER(C.c_int32), C.POINTER(C.c_double), _P], C.c_int),
    "cm_pqs_ctx_info": ([_P, C.POINTER(C.c_int32)], C.c_int),
    "cm_mg_zebra": ([_I, _I, _I] + [_P] * 6 + [_I, _P], C.c_int),
    "cm_mg_pack_f32_bytes": ([_I, _I], C.c_int64),
    "cm_mg_pack_f32": ([_I, _I] + [_P] * 5 + [_P], C.c_int),
    "cm_mg_zebra_f32": ([_I, _I, _I] + [_P] * 6 + [_I, _P, _P], C.c_int),
    "cm_mg_zebra_mid": ([_I, _I, _I, _I] + [_P] * 8 + [_P], C.c_int),
    "cm_mg_ring_debug_times": ([_P, _I], C.c_int),
    "cm_mg_smem_debug_times": ([_P, _I], C.c_int),
    "cm_mg_resid_restrict": ([_I, _I] + [_P] * 4 + [_P], C.c_int),
    "cm_mg_prolong_odd": ([_I, _I, _P, _P, _P], C.c_int),
    "cm_ipc_alloc": ([C.c_int64], C.c_void_p),
    "cm_ipc_free": ([_P], C.c_int),
    "cm_ipc_get_handle": ([_P, _P], C.c_int),
    "cm_ipc_open": ([_P], C.c_void_p),
    "cm_ipc_close": ([_P], C.c_int),
    "cm_pqs_state_offset": ([_I, _I, _I], C.c_int64),
    "cm_assemble_pq_cells_structured_rows": ([C.POINTER(PQParamsC), _I, _I] + [_P] * 7 + [_I, _I, _P], C.c_int),
    "cm_assemble_pq_cells_structured_blocks": ([C.POINTER(PQParamsC), _I, _I] + [_P] * 7 + [_P, _I, _I, _P], C.c_int),
    "cm_apply_dirichlet_blocks": ([_I, _P, _I, _P, _P], C.c_int),
    "cm_pqs_ctx_blocks": ([_P, _P], C.c_int),
    "cm_pqs_ctx_setup_from_blocks": ([_P, _P], C.c_int),
    "cm_dist_exchange_rows": ([C.POINTER(DistC), C.c_int64, _I, _I, _I, _P], C.c_int),
    "cm_dist_allgather_rows": ([C.POINTER(DistC), C.c_int64, _I, _I, _I, _P], C.c_int),
    "cm_dist_allreduce_sum": ([C.POINTER(DistC), _P, _I, _P], C.c_int),
    "cm_pqs_setup_dist": ([_I, _I, _I, _P, _P, _P, _P, C.c_int64, C.POINTER(DistC), _P], C.c_int),
    "cm_pqs_solve_dist": ([_I, _I, _I, _P, _P, _D, _D, _I, _P, C.c_int64, C.POINTER(DistC),
                           C.POINTER(C.c_int32), C.POINTER(C.c_double), _P], C.c_int),
    "cm_stencil_line_factor": ([_I, _I] + [_P] * 5 + [_P], C.c_int),
    "cm_stencil_zebra": ([_I, _I, _I] + [_P] * 6 + [_I, _P, _P], C.c_int),
    "cm_stencil_residual": ([_I, _I] + [_P] * 4 + [_P], C.c_int),
    "cm_blu_workspace_bytes": ([_I, _I, _I], C.c_int64),
    "cm_blu_create": ([], C.c_void_p),
    "cm_blu_destroy": ([_P], C.c_int),
    "cm_blu_factor": ([_P, _I, _I, _I] + [_P] * 6 + [C.c_int64, C.POINTER(C.c_int32), _P], C.c_int),
    "cm_blu_solve": ([_P, _I, _I, _I] + [_P] * 4 + [C.c_int64, _P], C.c_int),
    "cm_functional_workspace_doubles": ([], C.c_int64),
    "cm_functional_error_norms": ([_I, _P, _I, _I, _I, _P, _P, _P, _I, _P, _P, _P, _P], C.c_int),
    "cm_project_flux_dg0": ([_I, _P, _D, _D, _I, _P, _P, _P, _P, _P], C.c_int),
    "cm_functional_boundary_flux": ([_I, _P, _P, _P, _P, _P, _P, _P], C.c_int),
    "cm_mixed_m0_cells": ([_I, _P, _D, _D, _D, _I, _P, _P, _P, _P, _I, _P, _P, _P, _P, _P, _P], C.c_int),
    "cm_mixed_m1_cells": ([_I, _P, _D, _D, _D, _D, _I, _D, _D, _D, _D, _I, _I, _I] + [_P] * 10, C.c_int),
    "cm_mixed_m1_errors": ([_I, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _P, _P, _P], C.c_int),
    "cm_cg2_exterior_facets": ([_I, _P, _D, _I] + [_P] * 10 + [_P], C.c_int),
    "cm_cg2_interior_facets": ([_I, _P, _D, _I] + [_P] * 6 + [_P], C.c_int),
    "cm_mixed_p2_cells": ([_I, _P, _D, _D, _D, _D, _I, _I, _I] + [_P] * 8 + [_P], C.c_int),
    "cm_p1_sparsity_workspace_bytes": ([_I, _I, _I], C.c_int64),
    "cm_build_p1_sparsity": ([_I, _I, _P, _I, _P, _P, _P, C.c_int64, C.POINTER(C.c_int64), _P, C.c_int64, _P], C.c_int),
    "cm_vec_scratch_doubles": ([], C.c_int64),
    "cm_dot_batch": ([C.c_int64, _P, C.c_int64, _I, _P, _P, _P, _P], C.c_int),
    "cm_axpy_batch": ([C.c_int64, _P, C.c_int64, _I, _P, _D, _P, _P, _P, _P], C.c_int),
    "cm_csr_gather": ([_I, _P, _P, _P, _P, _I, _P], C.c_int),
    "cm_csr_add_mapped": ([_I, _P, _P, _P, _P], C.c_int),
    "cm_mixed_workspace_doubles": ([], C.c_int64),
    "cm_mixed_m0_errors": ([_I, _P, _I, _P, _P, _P, _P, _I, _P, _P, _P, _P], C.c_int),
    "cm_krylov_workspace_bytes": ([_I, _I, _I], C.c_int64),
    "cm_krylov_solve": ([_I, _I, _P, _P, _P, _P, _P, _P, _D, _D, _I, _I, _P, C.c_int64,
                         C.POINTER(C.c_int32), C.POINTER(C.c_double), _P], C.c_int),
    "cm_bicgstab_workspace_bytes": ([_I, _I], C.c_int64),
    "cm_bicgstab_solve": ([_I, _I, _P, _P, _P, _P, _P, _P, _D, _D, _I, _P, C.c_int64,
                           C.POINTER(C.c_int32), C.POINTER(C.c_double), _P], C.c_int),
}


def declared_symbols():
    """Names of every function include/closestmol_b200.h declares (parsed from the header)."""
    import re
    hdr = os.path.join(os.path.dirname(_HERE), "include", "closestmol_b200.h")
    txt = open(hdr).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(cm_[a-z0-9_]+)\s*\(", txt)))


def lib():
    """Load the shared library (building it is __graft_entry__.build()'s job)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CMError(f"{LIB_PATH} is missing: run `python -m closestmolecule_b200.build` "
                      "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    for name, (args, res) in _PROTOS.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = res
    _lib = L
    return L


def register(name, args, res=C.c_int):
    _PROTOS[name] = (args, res)
    if _lib is not None:
        fn = getattr(_lib, name)
        fn.argtypes = args
        fn.restype = res


def check(rc):
    if rc != 0:
        raise CMError(f"closestmol_b200 error {rc}: {lib().cm_last_error().decode()}")


class PQBlocksC(C.Structure):
    """struct cm_pq_blocks (include/closestmol_b200.h): stencil form of the P-Q Jacobian inside a solver workspace"""
    _fields_ = [(n, C.c_void_p) for n in ("PP", "PQ", "QP", "QQ", "sp", "sq")]


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return None
    assert t.is_contiguous()
    return t.data_ptr()


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(t):
    if not t.is_cuda:
        raise CMError("closestmolecule_b200 computes on CUDA tensors only (no CPU fallback)")
