"""Host restatement of the block-tridiagonal pivoted LU of csrc/cm_blu.cu (same data flow: panel getrf,
pivot sequence -> row permutation, fused gather, trsm, Schur update, permuted forward/backward
substitution), checked against SuperLU on Galerkin P-Q Jacobians whose q-q diagonal is zero.
It pins the index conventions of the CUDA kernels; the GPU parity test is tests/test_gpu_direct.py."""
import numpy as np
import scipy.linalg as sla
import scipy.sparse as sp
import scipy.sparse.linalg as spla

from oracle import fem, pq


def _src_from_ipiv(ipiv, m):
    s = np.arange(m)
    for i, j in enumerate(ipiv):
        if j != i:
            s[i], s[j] = s[j], s[i]
    return s


def blu_factor(A, nb):
    """A: dense (n x n) block tridiagonal with block size nb, n = nblk*nb."""
    n = A.shape[0]
    nblk = n // nb
    blk = lambda i, j: A[i * nb:(i + 1) * nb, j * nb:(j + 1) * nb].copy()
    P = [np.zeros((2 * nb, nb)) for _ in range(nblk)]
    U = [np.zeros((nb, 2 * nb)) for _ in range(nblk)]
    src = [None] * nblk
    P[0][:nb] = blk(0, 0)
    T = blk(0, 1) if nblk > 1 else None
    for k in range(nblk):
        if k == nblk - 1:
            lu, piv = sla.lu_factor(P[k][:nb])
            P[k][:nb] = lu
            src[k] = _src_from_ipiv(piv, nb)
            break
        P[k][nb:] = blk(k + 1, k)
        orig = np.zeros((nb, 2 * nb))
        orig[:, :nb] = blk(k + 1, k + 1)
        if k + 2 < nblk:
            orig[:, nb:] = blk(k + 1, k + 2)
        # getrf of the 2nb x nb panel (LAPACK semantics through scipy)
        lu, piv = sla.lu_factor(P[k])
        P[k][:] = lu
        s = _src_from_ipiv(piv, 2 * nb)
        src[k] = s
        stacked = np.vstack([np.hstack([T, np.zeros((nb, nb))]), orig])[s]
        U[k][:] = stacked[:nb]
        C = stacked[nb:]
        L11 = np.tril(P[k][:nb], -1) + np.eye(nb)
        U[k][:] = sla.solve_triangular(L11, U[k], lower=True, unit_diagonal=True)
        C = C - P[k][nb:] @ U[k]
        P[k + 1][:nb] = C[:, :nb]
        T = C[:, nb:]
    return P, U, src


def blu_solve(P, U, src, b, nb):
    nblk = len(P)
    y = np.zeros((nblk + 2) * nb)
    y[:nblk * nb] = b
    for k in range(nblk):
        m = nb if k == nblk - 1 else 2 * nb
        y[k * nb:k * nb + m] = y[k * nb:k * nb + m][src[k]]
        yk = y[k * nb:(k + 1) * nb]
        yk[:] = sla.solve_triangular(P[k][:nb], yk, lower=True, unit_diagonal=True)
        if k < nblk - 1:
            y[(k + 1) * nb:(k + 2) * nb] -= P[k][nb:] @ yk
    for k in range(nblk - 1, -1, -1):
        yk = y[k * nb:(k + 1) * nb]
        if k < nblk - 1:
            yk -= U[k] @ y[(k + 1) * nb:(k + 3) * nb]
        yk[:] = sla.solve_triangular(P[k][:nb], yk, lower=False)
    return y[:nblk * nb]


def _galerkin_jacobian(nx, ny, seed=0):
    m = fem.unit_square_mesh(nx, ny)
    prm = pq.PQParams(D1=1e-3, D2=1.5e-3, gkind=pq.G_POWER, gcoef=1.5e-3)
    S = pq.PQSystem(m, prm)
    rng = np.random.default_rng(seed)
    u = np.empty(S.N)
    u[0::2] = 0.3 + rng.random(S.N // 2)
    u[1::2] = 1.0 + rng.random(S.N // 2)
    vals, F = S.assemble(u, True, None, None)
    A = sp.csr_matrix((vals, S.col, S.rowptr), shape=(S.N, S.N)).tolil()
    # p Dirichlet on the whole boundary, q Dirichlet on x = 1 (…MixedPrimal_convergence.py:108-110)
    X = m.coords
    onb = (X[:, 0] < 1e-14) | (X[:, 0] > 1 - 1e-14) | (X[:, 1] < 1e-14) | (X[:, 1] > 1 - 1e-14)
    rows = np.concatenate([2 * np.nonzero(onb)[0], 2 * np.nonzero(X[:, 0] > 1 - 1e-14)[0] + 1])
    for r in rows:
        A.rows[r], A.data[r] = [r], [1.0]
    return m, A.tocsr(), F


def test_block_lu_matches_superlu_on_zero_diagonal_system():
    nx, ny = 12, 7
    m, A, F = _galerkin_jacobian(nx, ny)
    V = (nx + 1) * (ny + 1)
    # transpose renumbering: position = ix*(ny+1) + iy, bandwidth ny+2 nodes
    ix, iy = np.arange(V) % (nx + 1), np.arange(V) // (nx + 1)
    pos = ix * (ny + 1) + iy
    nbn = ny + 3            # deliberately not a divisor of V: exercises the padding block
    nb = 2 * nbn
    nblk = -(-V // nbn)
    n = nblk * nb
    dof = np.empty(2 * V, dtype=np.int64)
    dof[0::2], dof[1::2] = 2 * pos, 2 * pos + 1
    Ad = np.eye(n)
    Ad[np.ix_(dof, dof)] = A.toarray()
    P, U, src = blu_factor(Ad, nb)
    b = np.zeros(n)
    b[dof] = F
    x = blu_solve(P, U, src, b, nb)[dof]
    ref = spla.splu(A.tocsc()).solve(F)
    assert np.linalg.norm(x - ref) / np.linalg.norm(ref) < 1e-10


def test_ordering_bandwidth_structured_and_rcm():
    """Host-side renumbering used by BandedLULinearSolver: grid lines along the shorter direction on
    RectangleMesh numbering, reverse Cuthill-McKee otherwise; nbn bounds the node distance of every couple."""
    import torch

    import closestmolecule_b200 as cm
    from closestmolecule_b200.solver import BandedLULinearSolver
    mesh = cm.RectangleMesh(cm.Point(0, 0), cm.Point(1, 1), 40, 12, device="cpu")
    for with_facets, expect in ((False, 16), (True, 32)):
        t = mesh.topology(with_facets)
        perm, iperm, nbn = BandedLULinearSolver.ordering(mesh, t)
        assert nbn == expect and torch.equal(perm[iperm.long()].long(), torch.arange(t.nv))
    g = torch.Generator().manual_seed(1)
    new = torch.randperm(mesh.num_vertices(), generator=g)
    coords = torch.empty_like(mesh.coords)
    coords[new] = mesh.coords
    cells, _ = torch.sort(new[mesh.cells.long()], dim=1)
    um = cm.Mesh(coords, cells.to(torch.int32), device="cpu")
    t = um.topology(False)
    perm, iperm, nbn = BandedLULinearSolver.ordering(um, t)
    rows = torch.repeat_interleave(torch.arange(t.nv), (t.nrowptr[1:] - t.nrowptr[:-1]).long())
    assert int((perm[rows].long() - perm[t.ncol.long()].long()).abs().max()) <= nbn <= 40


# ---- two-sided ("twisted") variant: chain A eliminates blocks 0..mA-1 from the top, chain B the blocks
# nblk-1..mA+2 from the bottom (= the same algorithm on the matrix with reversed node order), the two
# leftover block rows meet in a 2nb x 2nb system for the blocks mA, mA+1 ------------------------------
def _chain(A, nb, nsteps):
    """blu_factor without the final block: returns P, U, src and the leftover [S | T] of block `nsteps`."""
    n = A.shape[0]
    nblk = n // nb
    blk = lambda i, j: A[i * nb:(i + 1) * nb, j * nb:(j + 1) * nb].copy()
    P = [np.zeros((2 * nb, nb)) for _ in range(nsteps)]
    U = [np.zeros((nb, 2 * nb)) for _ in range(nsteps)]
    src = [None] * nsteps
    S, T = blk(0, 0), blk(0, 1)
    for k in range(nsteps):
        P[k][:nb], P[k][nb:] = S, blk(k + 1, k)
        orig = np.zeros((nb, 2 * nb))
        orig[:, :nb] = blk(k + 1, k + 1)
        if k + 2 < nblk:
            orig[:, nb:] = blk(k + 1, k + 2)
        lu, piv = sla.lu_factor(P[k])
        P[k][:] = lu
        s = src[k] = _src_from_ipiv(piv, 2 * nb)
        stacked = np.vstack([np.hstack([T, np.zeros((nb, nb))]), orig])[s]
        L11 = np.tril(P[k][:nb], -1) + np.eye(nb)
        U[k][:] = sla.solve_triangular(L11, stacked[:nb], lower=True, unit_diagonal=True)
        C = stacked[nb:] - P[k][nb:] @ U[k]
        S, T = C[:, :nb], C[:, nb:]
    return P, U, src, S, T


def _chain_forward(P, src, y, nb):
    for k in range(len(P)):
        y[k * nb:(k + 2) * nb] = y[k * nb:(k + 2) * nb][src[k]]
        yk = y[k * nb:(k + 1) * nb]
        yk[:] = sla.solve_triangular(P[k][:nb], yk, lower=True, unit_diagonal=True)
        y[(k + 1) * nb:(k + 2) * nb] -= P[k][nb:] @ yk


def _chain_backward(P, U, y, nb):
    for k in range(len(P) - 1, -1, -1):
        yk = y[k * nb:(k + 1) * nb]
        yk -= U[k] @ y[(k + 1) * nb:(k + 3) * nb]
        yk[:] = sla.solve_triangular(P[k][:nb], yk, lower=False)


def twisted_solve(A, b, nb, bs):
    n = A.shape[0]
    nblk = n // nb
    nbn = nb // bs
    mA = (nblk - 2) // 2
    mB = nblk - 2 - mA
    # node reversal, component order kept: rev(bs*l + c) within a block, flip of the block order
    loc = np.arange(nb)
    rev = bs * (nbn - 1 - loc // bs) + loc % bs
    flip = np.concatenate([(nblk - 1 - k) * nb + rev for k in range(nblk)])     # flipped index -> original index
    Af, bf = A[np.ix_(flip, flip)], b[flip]
    PA, UA, sA, SA, TA = _chain(A, nb, mA)
    PB, UB, sB, SB, TB = _chain(Af, nb, mB)
    M = np.block([[SA, TA[:, rev]], [TB[:, rev], SB]])          # unknowns [x_mA (orig order); x'_mB (flipped order)]
    yA = np.zeros((nblk + 2) * nb)
    yB = np.zeros((nblk + 2) * nb)
    yA[:n], yB[:n] = b, bf
    _chain_forward(PA, sA, yA, nb)
    _chain_forward(PB, sB, yB, nb)
    z = np.linalg.solve(M, np.concatenate([yA[mA * nb:(mA + 1) * nb], yB[mB * nb:(mB + 1) * nb]]))
    yA[mA * nb:(mA + 1) * nb], yA[(mA + 1) * nb:(mA + 2) * nb] = z[:nb], z[nb:][rev]
    yB[mB * nb:(mB + 1) * nb], yB[(mB + 1) * nb:(mB + 2) * nb] = z[nb:], z[:nb][rev]
    _chain_backward(PA, UA, yA, nb)
    _chain_backward(PB, UB, yB, nb)
    x = np.empty(n)
    x[:(mA + 1) * nb] = yA[:(mA + 1) * nb]
    xb = np.empty(n)
    xb[flip] = yB[:n]                                           # back to the original order
    x[(mA + 1) * nb:] = xb[(mA + 1) * nb:]
    return x


def test_twisted_block_lu_matches_superlu():
    for nx, ny, extra in ((12, 7, 3), (9, 5, 2), (16, 3, 4)):
        m, A, F = _galerkin_jacobian(nx, ny, seed=nx)
        V = (nx + 1) * (ny + 1)
        ix, iy = np.arange(V) % (nx + 1), np.arange(V) // (nx + 1)
        pos = ix * (ny + 1) + iy
        nbn = ny + extra
        nb = 2 * nbn
        nblk = -(-V // nbn)
        assert nblk >= 4
        n = nblk * nb
        dof = np.empty(2 * V, dtype=np.int64)
        dof[0::2], dof[1::2] = 2 * pos, 2 * pos + 1
        Ad = np.eye(n)
        Ad[np.ix_(dof, dof)] = A.toarray()
        b = np.zeros(n)
        b[dof] = F
        x = twisted_solve(Ad, b, nb, 2)[dof]
        ref = spla.splu(A.tocsc()).solve(F)
        assert np.linalg.norm(x - ref) / np.linalg.norm(ref) < 1e-9, (nx, ny)
