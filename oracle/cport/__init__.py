"""ctypes wrapper of the plain-C restatement (test infrastructure / CPU baseline only)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libpq_cport.so")


class CParams(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "D1", "D2", "inv_dt", "gcoef", "geps", "gs_thr", "gs_cond_scale", "gs_lin", "q_react",
        "q_adv", "q_x2p", "q_x2q", "q_divr", "q_ibp", "supg_stab")] + [("gkind", C.c_int32), ("nq", C.c_int32)]


def build():
    src = os.path.join(_HERE, "pq_cells.c")
    if not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        # -march=native is avoided on purpose: the .so travels to a different host
        subprocess.check_call(["gcc", "-O3", "-fPIC", "-shared", "-std=c11", "-o", _LIB, src, "-lm"])
    return _LIB


def lib():
    L = C.CDLL(build())
    L.pq_assemble_cells.argtypes = [C.POINTER(CParams), C.c_int64] + [C.c_void_p] * 8
    L.pq_assemble_cells.restype = None
    return L


def assemble_cells(prm, cells, coords, u, slots, nnz, p_old=None, fq=None, want_jac=True):
    """prm: oracle.pq.PQParams (degree 4 only).  Returns (vals or None, b)."""
    assert prm.degree == 4
    cp = CParams()
    for n, _ in CParams._fields_[:-2]:
        setattr(cp, n, getattr(prm, n))
    cp.gkind, cp.nq = prm.gkind, 6
    cells = np.ascontiguousarray(cells, dtype=np.int32)
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    slots = np.ascontiguousarray(slots, dtype=np.int64)
    b = np.zeros(u.size)
    vals = np.zeros(nnz) if want_jac else None
    p = lambda a: None if a is None else a.ctypes.data
    if p_old is not None:
        p_old = np.ascontiguousarray(p_old, dtype=np.float64)
    if fq is not None:
        fq = np.ascontiguousarray(fq, dtype=np.float64)
    lib().pq_assemble_cells(C.byref(cp), cells.shape[0], p(cells), p(coords), p(u), p(p_old), p(fq),
                            p(slots), p(vals), p(b))
    return vals, b
