"""CPU (numpy / scipy) checks of the two identities the fused multigrid kernels rest on (csrc/cm_mgmid.cu,
csrc/cm_mgsmem.cu) -- no GPU involved: the kernels themselves are compared with these operations in
tests/test_gpu_mg.py.
  1. right after a zebra sweep (colour 0, then colour 1) the residual vanishes on the odd rows and, on the even
     rows, equals  -(A - T)(x_odd_new - x_odd_old)  (T = tridiagonal x-line part of the 7-point stencil): the
     restricted residual needs the four off-line diagonals and the CHANGE of the odd rows only;
  2. the coarse-grid correction is only ever read through the odd rows next to the first colour-0 sweep of the up
     leg: a colour-0 sweep that reads x_odd + P e on the fly equals "x += P e, then the sweep" on the even rows, and
     the odd rows are overwritten by the colour-1 sweep that follows without being read.
"""
import numpy as np

from test_gpu_mg import OFFS, _apply, _random_stencil, _zebra_ref


def _prolong(ec, nx, ny):
    """P1 interpolation on the nested right-diagonal meshes (weights 1 and 1/2 on the six neighbours)"""
    P = np.zeros((ny + 1, nx + 1))
    for iy in range(ny + 1):
        for ix in range(nx + 1):
            X, Y, ox, oy = ix // 2, iy // 2, ix & 1, iy & 1
            if not ox and not oy:
                e = ec[Y, X]
            elif ox and not oy:
                e = 0.5 * (ec[Y, X] + ec[Y, X + 1])
            elif not ox and oy:
                e = 0.5 * (ec[Y, X] + ec[Y + 1, X])
            else:
                e = 0.5 * (ec[Y, X] + ec[Y + 1, X + 1])
            P[iy, ix] = e
    return P


def test_even_row_residual_from_the_change_of_the_odd_rows():
    nx, ny = 24, 14
    A, rng = _random_stencil(nx, ny, 11)
    b = rng.standard_normal((ny + 1, nx + 1))
    x0 = rng.standard_normal((ny + 1, nx + 1))
    x1 = _zebra_ref(A, b, x0, 0, False)          # colour 0: the even lines are solved against the OLD odd rows
    x2 = _zebra_ref(A, b, x1, 1, False)          # colour 1
    R = b - _apply(A, x2)
    assert np.max(np.abs(R[1::2])) < 1e-12
    off = A.copy()
    off[2:5] = 0.0                               # A - T: the couplings to the neighbouring lines
    d = np.zeros_like(x2)
    d[1::2] = x2[1::2] - x1[1::2]
    assert np.max(np.abs(R[0::2] + _apply(off, d)[0::2])) < 1e-12
    # from a zero guess the old odd rows are zero: the change is the new odd rows themselves
    z1 = _zebra_ref(A, b, np.zeros_like(b), 0, True)
    z2 = _zebra_ref(A, b, z1, 1, False)
    dz = np.zeros_like(z2)
    dz[1::2] = z2[1::2]
    assert np.max(np.abs((b - _apply(A, z2))[0::2] + _apply(off, dz)[0::2])) < 1e-12


def test_prolongation_folded_into_the_first_colour0_sweep():
    nx, ny = 20, 12
    A, rng = _random_stencil(nx, ny, 23)
    b = rng.standard_normal((ny + 1, nx + 1))
    x = rng.standard_normal((ny + 1, nx + 1))
    ec = rng.standard_normal((ny // 2 + 1, nx // 2 + 1))
    P = _prolong(ec, nx, ny)
    # reference: full prolongation, then colour 0, then colour 1
    ref = _zebra_ref(A, b, x + P, 0, False)
    ref = _zebra_ref(A, b, ref, 1, False)
    # fused: only the odd rows carry the correction into the colour-0 sweep; nothing of it is stored
    xin = x.copy()
    xin[1::2] += P[1::2]
    fused = _zebra_ref(A, b, xin, 0, False)
    fused[1::2] = x[1::2]                        # the kernel leaves the (stale) odd rows untouched ...
    fused = _zebra_ref(A, b, fused, 1, False)    # ... and the colour-1 sweep overwrites them without reading them
    assert np.max(np.abs(fused - ref)) < 1e-12 * (1 + np.max(np.abs(ref)))
