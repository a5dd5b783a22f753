"""Second-generation multigrid kernels (csrc/cm_mg.cu) against numpy / scipy restatements of the same
operations: fused zebra half sweep (shared-memory ring fed by cp.async.bulk on long lines, one CTA per line
on short ones), restricted residual on even rows, correction on odd rows; and the solver-level equivalence
of the two kernel generations (same GMRES iteration counts, same solution)."""
import numpy as np
import pytest
import scipy.linalg as sla
import torch

import closestmolecule_b200 as cm
from closestmolecule_b200 import _capi

pytestmark = pytest.mark.gpu

OFFS = [(-1, -1), (0, -1), (-1, 0), (0, 0), (1, 0), (0, 1), (1, 1)]


def _random_stencil(nx, ny, seed):
    rng = np.random.default_rng(seed)
    n1 = nx + 1
    A = np.zeros((7, ny + 1, n1))
    for k, (dx, dy) in enumerate(OFFS):
        v = (1.0 + rng.random((ny + 1, n1))) if k == 3 else -(0.1 + 0.05 * rng.random((ny + 1, n1)))
        if dx == -1: v[:, 0] = 0
        if dx == 1: v[:, -1] = 0
        if dy == -1: v[0, :] = 0
        if dy == 1: v[-1, :] = 0
        A[k] = v
    return A, rng


def _apply(A, x):
    """y = A x for the 7-point stencil stored as diagonals (numpy)."""
    ny1, n1 = x.shape
    y = np.zeros_like(x)
    xp = np.pad(x, 1)
    for k, (dx, dy) in enumerate(OFFS):
        y += A[k] * xp[1 + dy:1 + dy + ny1, 1 + dx:1 + dx + n1]
    return y


def _line_solve(A, j, rhs):
    n1 = rhs.size
    ab = np.zeros((3, n1))
    ab[0, 1:] = A[4, j, :-1]
    ab[1] = A[3, j]
    ab[2, :-1] = A[2, j, 1:]
    return sla.solve_banded((1, 1), ab, rhs)


def _zebra_ref(A, b, x, colour, x_zero):
    """one half sweep in numpy (only a few lines are checked by the caller: this does all of them)"""
    ny1, n1 = b.shape
    xin = np.zeros_like(x) if x_zero else x
    T = np.zeros_like(A)
    T[2:5] = A[2:5]
    off = _apply(A - T, xin)
    out = x.copy()
    for j in range(colour, ny1, 2):
        out[j] = _line_solve(A, j, b[j] - off[j])
    return out


@pytest.mark.parametrize("nx,ny,shift", [(200, 37, 0), (1024, 40, 0), (1024, 1300, 1), (1500, 21, 1), (2048, 330, 0),
                                         (3000, 9, 1), (4096, 310, 1), (4479, 6, 0)])
def test_fused_zebra_half_sweeps(nx, ny, shift):
    """colour 0 from a zero guess, then colour 1 with couplings, then colour 0 with couplings: every line
    equals scipy's tridiagonal solve with the couplings of the current iterate.  `shift` places b and x one
    double off a 16-byte boundary (the q-part of a split vector with an odd vertex count): the bulk copies
    then start one element early and fetch array ends by plain loads."""
    n1, n = nx + 1, (nx + 1) * (ny + 1)
    dev = torch.device("cuda")
    A, rng = _random_stencil(nx, ny, nx + ny)
    At = torch.tensor(A.reshape(7, n), device=dev)
    lf, iu, cu = (torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3))
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    bh = rng.standard_normal((ny + 1, n1))
    # one buffer holds [pad | b | x] so that neighbouring memory is valid and the alignment is controlled
    buf = torch.full((2 * n + 8,), float("nan"), dtype=torch.float64, device=dev)
    b = buf[shift:shift + n]
    xo = shift + n + 2
    xo += (xo - shift) % 2          # x starts with the same parity as b
    x = buf[xo:xo + n]
    assert (b.data_ptr() // 8) % 2 == shift % 2 and (x.data_ptr() // 8) % 2 == shift % 2
    b.copy_(torch.tensor(bh.reshape(-1)))
    x.zero_()
    lib, P, st = _capi.lib(), _capi.ptr, _capi.stream_ptr()
    _capi.check(lib.cm_stencil_line_factor(nx, ny, P(At), P(lf), P(iu), P(cu), P(err), st))
    assert int(err.item()) == 0
    lines = sorted({0, 1, 2, 3, ny // 2, ny // 2 + 1, ny - 1, ny})
    xr = np.zeros((ny + 1, n1))
    for colour, x_zero in ((0, 1), (1, 0), (0, 0), (1, 0)):
        _capi.check(lib.cm_mg_zebra(nx, ny, colour, P(At), P(lf), P(iu), P(cu), P(b), P(x), x_zero, st))
        X = x.cpu().numpy().reshape(ny + 1, n1)
        assert np.isfinite(X[colour::2]).all()
        # reference on the checked lines only (the full numpy sweep is slow on the big grids)
        xin = np.zeros_like(xr) if x_zero else xr
        for j in lines:
            if j % 2 != colour:
                continue
            rhs = bh[j].copy()
            for k, (dx, dy) in enumerate(OFFS):
                if dy == 0 or not (0 <= j + dy <= ny):
                    continue
                src = np.zeros(n1)
                lo, hi = max(0, -dx), min(n1, n1 - dx)
                src[lo:hi] = xin[j + dy, lo + dx:hi + dx]
                rhs -= A[k, j] * src
            ref = _line_solve(A, j, rhs)
            assert np.max(np.abs(X[j] - ref)) < 1e-12 * (1 + np.max(np.abs(ref))), (colour, j)
        # continue from the GPU iterate (all lines of the colour were updated there)
        xr = X.copy()
    # after a sweep that ends with colour 1 the residual vanishes on the odd rows
    R = bh - _apply(A, xr)
    assert np.max(np.abs(R[1::2])) < 1e-11 and np.max(np.abs(R[0::2])) > 1e-4


@pytest.mark.parametrize("nx,ny,shift", [(1024, 40, 0), (1500, 21, 1), (2048, 330, 0), (4096, 310, 1), (4099, 7, 0)])
def test_fused_zebra_half_sweeps_fp32_operator(nx, ny, shift):
    """cm_mg_zebra_f32: the long-line half sweep with its operator streams read from the fp32 pack equals the
    sequential line solve (two first-order recurrences, fp64 arithmetic) with the fp32-ROUNDED couplings and line
    factors, to fp64 rounding; and it differs from the fp64-operator sweep by fp32 rounding only."""
    n1, n = nx + 1, (nx + 1) * (ny + 1)
    dev = torch.device("cuda")
    A, rng = _random_stencil(nx, ny, 7 * nx + ny)
    At = torch.tensor(A.reshape(7, n), device=dev)
    lf, iu, cu = (torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3))
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    bh = rng.standard_normal((ny + 1, n1))
    buf = torch.full((2 * n + 8,), float("nan"), dtype=torch.float64, device=dev)
    b = buf[shift:shift + n]
    xo = shift + n + 2
    xo += (xo - shift) % 2
    x = buf[xo:xo + n]
    b.copy_(torch.tensor(bh.reshape(-1)))
    x.zero_()
    lib, P, st = _capi.lib(), _capi.ptr, _capi.stream_ptr()
    _capi.check(lib.cm_stencil_line_factor(nx, ny, P(At), P(lf), P(iu), P(cu), P(err), st))
    assert int(err.item()) == 0
    pack = torch.empty(int(lib.cm_mg_pack_f32_bytes(nx, ny)) // 8, dtype=torch.float64, device=dev)
    _capi.check(lib.cm_mg_pack_f32(nx, ny, P(At), P(lf), P(iu), P(cu), P(pack), st))
    r32 = lambda a: a.astype(np.float32).astype(np.float64)
    A32 = r32(A)
    L32, I32, C32 = (r32(v.cpu().numpy().reshape(ny + 1, n1)) for v in (lf, iu, cu))
    lines = sorted({0, 1, 2, 3, ny // 2, ny // 2 + 1, ny - 1, ny})
    xr = np.zeros((ny + 1, n1))
    x64 = x.clone()
    for colour, x_zero in ((0, 1), (1, 0), (0, 0), (1, 0)):
        x64.copy_(x)
        _capi.check(lib.cm_mg_zebra(nx, ny, colour, P(At), P(lf), P(iu), P(cu), P(b), P(x64), x_zero, st))
        _capi.check(lib.cm_mg_zebra_f32(nx, ny, colour, P(At), P(lf), P(iu), P(cu), P(b), P(x), x_zero, P(pack), st))
        X = x.cpu().numpy().reshape(ny + 1, n1)
        assert np.isfinite(X[colour::2]).all()
        xin = np.zeros_like(xr) if x_zero else xr
        for j in lines:
            if j % 2 != colour:
                continue
            rhs = bh[j].copy()
            for k, (dx, dy) in enumerate(OFFS):
                if dy == 0 or not (0 <= j + dy <= ny):
                    continue
                src = np.zeros(n1)
                lo, hi = max(0, -dx), min(n1, n1 - dx)
                src[lo:hi] = xin[j + dy, lo + dx:hi + dx]
                rhs -= A32[k, j] * src
            z = np.empty(n1)
            acc = 0.0
            for i in range(n1):
                acc = rhs[i] - L32[j, i] * acc
                z[i] = acc
            w = z * I32[j]
            ref = np.empty(n1)
            acc = 0.0
            for i in range(n1 - 1, -1, -1):
                acc = w[i] - C32[j, i] * acc
                ref[i] = acc
            assert np.max(np.abs(X[j] - ref)) < 1e-11 * (1 + np.max(np.abs(ref))), (colour, j)
        X64 = x64.cpu().numpy().reshape(ny + 1, n1)
        d = np.max(np.abs(X[colour::2] - X64[colour::2])) / (1 + np.max(np.abs(X64[colour::2])))
        assert d < 1e-5, (colour, d)
        xr = X.copy()


@pytest.mark.parametrize("nx,ny", [(8, 6), (200, 38), (1024, 20), (514, 130)])
def test_restricted_residual_and_odd_row_correction(nx, ny):
    n1, n = nx + 1, (nx + 1) * (ny + 1)
    nxc, nyc = nx // 2, ny // 2
    m1 = nxc + 1
    dev = torch.device("cuda")
    A, rng = _random_stencil(nx, ny, 5 * nx + ny)
    bh = rng.standard_normal((ny + 1, n1))
    xh = rng.standard_normal((ny + 1, n1))
    # make the residual vanish on the odd rows (what a colour-1 half sweep leaves behind)
    xh = _zebra_ref(A, bh, xh, 1, False)
    R = bh - _apply(A, xh)
    assert np.max(np.abs(R[1::2])) < 1e-11
    # full-weighting restriction of the nested right-diagonal meshes: weights 1 and 1/2 on the six neighbours
    Rp = np.pad(R, 1)
    full = np.zeros((nyc + 1, m1))
    for Y in range(nyc + 1):
        for k, (dx, dy) in enumerate(OFFS):
            w = 1.0 if k == 3 else 0.5
            full[Y] += w * Rp[1 + 2 * Y + dy, 1 + dx:1 + dx + 2 * nxc + 1:2]
    lib, P, st = _capi.lib(), _capi.ptr, _capi.stream_ptr()
    At = torch.tensor(A.reshape(7, n), device=dev)
    b = torch.tensor(bh.reshape(-1), device=dev)
    x = torch.tensor(xh.reshape(-1), device=dev)
    rc = torch.full(((nyc + 1) * m1,), float("nan"), dtype=torch.float64, device=dev)
    _capi.check(lib.cm_mg_resid_restrict(nx, ny, P(At), P(b), P(x), P(rc), st))
    got = rc.cpu().numpy().reshape(nyc + 1, m1)
    assert np.max(np.abs(got - full)) < 1e-11 * (1 + np.max(np.abs(full)))
    # prolongation: odd fine rows get the interpolated correction, even rows are left alone
    ec = rng.standard_normal((nyc + 1, m1))
    P2 = np.zeros((ny + 1, n1))
    for iy in range(ny + 1):
        for ix in range(n1):
            X, Y, ox, oy = ix // 2, iy // 2, ix & 1, iy & 1
            if not ox and not oy: e = ec[Y, X]
            elif ox and not oy: e = 0.5 * (ec[Y, X] + ec[Y, X + 1])
            elif not ox and oy: e = 0.5 * (ec[Y, X] + ec[Y + 1, X])
            else: e = 0.5 * (ec[Y, X] + ec[Y + 1, X + 1])
            P2[iy, ix] = e
    x0 = xh.copy()
    ect = torch.tensor(ec.reshape(-1), device=dev)
    _capi.check(lib.cm_mg_prolong_odd(nxc, nyc, P(ect), P(x), st))
    got = x.cpu().numpy().reshape(ny + 1, n1)
    assert np.max(np.abs(got[1::2] - (x0 + P2)[1::2])) < 1e-13
    assert np.array_equal(got[0::2], x0[0::2])


@pytest.mark.parametrize("nx,ny", [(8, 6), (64, 32), (128, 64), (200, 38), (256, 128), (512, 20), (514, 130),
                                   (1024, 36), (1278, 10)])
def test_mid_level_half_sweeps(nx, ny):
    """cm_mgmid.cu: mode 0 / 1 half sweeps equal scipy's tridiagonal solves; mode 2 (colour 1 + restricted
    residual in one launch) equals the separate colour-1 sweep followed by cm_mg_resid_restrict BIT FOR BIT in x
    and to rounding in the coarse right-hand side; mode 3 (colour 0 reading x + P e on the odd rows) equals
    cm_mg_prolong_odd followed by a colour-0 sweep."""
    n1, n = nx + 1, (nx + 1) * (ny + 1)
    nxc, nyc = nx // 2, ny // 2
    m1 = nxc + 1
    dev = torch.device("cuda")
    A, rng = _random_stencil(nx, ny, 3 * nx + ny)
    At = torch.tensor(A.reshape(7, n), device=dev)
    lf, iu, cu = (torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3))
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    bh = rng.standard_normal((ny + 1, n1))
    b = torch.tensor(bh.reshape(-1), device=dev)
    x = torch.full((n,), float("nan"), dtype=torch.float64, device=dev)
    lib, P, st = _capi.lib(), _capi.ptr, _capi.stream_ptr()
    _capi.check(lib.cm_stencil_line_factor(nx, ny, P(At), P(lf), P(iu), P(cu), P(err), st))
    assert int(err.item()) == 0

    def mid(mode, colour, xx, ec=None, bc=None):
        _capi.check(lib.cm_mg_zebra_mid(mode, nx, ny, colour, P(At), P(lf), P(iu), P(cu), P(b), P(xx),
                                        P(ec) if ec is not None else None, P(bc) if bc is not None else None, st))

    # colour 0 from a zero guess (x holds NaN: it must not be read), colour 1 with couplings, colour 0 with couplings
    mid(0, 0, x)
    x[torch.isnan(x)] = 0.0
    xr = np.zeros((ny + 1, n1))
    xr = _zebra_ref(A, bh, xr, 0, True)
    assert np.max(np.abs(x.cpu().numpy().reshape(ny + 1, n1) - xr)) < 1e-12 * (1 + np.max(np.abs(xr)))
    for colour in (1, 0):
        xr = x.cpu().numpy().reshape(ny + 1, n1).copy()
        mid(1, colour, x)
        ref = _zebra_ref(A, bh, xr, colour, False)
        got = x.cpu().numpy().reshape(ny + 1, n1)
        assert np.max(np.abs(got - ref)) < 1e-12 * (1 + np.max(np.abs(ref))), colour
    fused_ok = nx % 2 == 0 and ny % 2 == 0
    if fused_ok and nx + 1 <= 640:
        # mode 2 against the two separate launches
        x1 = x.clone()
        mid(1, 1, x1)
        rc1 = torch.full(((nyc + 1) * m1,), float("nan"), dtype=torch.float64, device=dev)
        _capi.check(lib.cm_mg_resid_restrict(nx, ny, P(At), P(b), P(x1), P(rc1), st))
        x2 = x.clone()
        rc2 = torch.full(((nyc + 1) * m1,), float("nan"), dtype=torch.float64, device=dev)
        mid(2, 1, x2, bc=rc2)
        assert torch.equal(x1, x2)
        assert torch.isfinite(rc2).all()
        assert float((rc1 - rc2).abs().max()) < 1e-12 * (1 + float(rc1.abs().max()))
        x.copy_(x2)
    if fused_ok:
        ec = torch.tensor(rng.standard_normal((nyc + 1) * m1), device=dev)
        x1 = x.clone()
        _capi.check(lib.cm_mg_prolong_odd(nxc, nyc, P(ec), P(x1), st))
        mid(1, 0, x1)
        x2 = x.clone()
        mid(3, 0, x2, ec=ec)
        X1, X2 = (v.cpu().numpy().reshape(ny + 1, n1) for v in (x1, x2))
        assert np.max(np.abs(X1[0::2] - X2[0::2])) < 1e-12 * (1 + np.max(np.abs(X1)))
        assert np.array_equal(X2[1::2], x.cpu().numpy().reshape(ny + 1, n1)[1::2])   # odd rows untouched


def _supg_solver(nx, ny, restart=40):
    from closestmolecule_b200.manufactured import p_exact, q_exact
    from closestmolecule_b200.solver import StructuredPQLinearSolver
    P1 = cm.FiniteElement("CG", "triangle", 1)
    mesh = cm.RectangleMesh(cm.Point(0, 0), cm.Point(1, 1), nx, ny)
    V = cm.FunctionSpace(mesh, cm.MixedElement([P1, P1]))
    bd = cm.MeshFunction("size_t", mesh, 1, 0)
    cm.CompiledSubDomain("near(x[0],0) && on_boundary").mark(bd, 31)
    form = cm.PQPrimalForm.manufactured(V, "supg")
    bcs = [cm.DirichletBC(V.sub(0), p_exact, "on_boundary"), cm.DirichletBC(V.sub(1), q_exact, bd, 31)]
    s = cm.NonlinearVariationalSolver(cm.NonlinearVariationalProblem(form, cm.Function(V), bcs))
    asm = s.asm
    X = mesh.coords
    u = torch.empty(2 * mesh.num_vertices(), dtype=torch.float64, device=mesh.device)
    u[0::2] = p_exact(X[:, 0], X[:, 1])
    u[1::2] = q_exact(X[:, 0], X[:, 1]) * (1 + 1e-3 * torch.sin(7 * X[:, 0] + 3 * X[:, 1]))
    bc = asm._bc_tables(bcs)
    asm.assemble(u, True)
    asm.apply_bcs(bc, u, True)
    return asm, (lambda: StructuredPQLinearSolver(asm, restart=restart))


@pytest.mark.parametrize("nx,ny", [(40, 24), (256, 128), (300, 130)])
def test_sweep_writes_the_solver_blocks(nx, ny):
    """The assembly sweep can write the stencil form of the Jacobian and the row scalings straight into the solver
    workspace (cm_assemble_pq_cells_structured_blocks + cm_apply_dirichlet_blocks): bit-identical to what
    cm_pqs_ctx_setup extracts from the CSR values, and the solve through cm_pqs_ctx_setup_from_blocks gives the
    same iterations and the same answer."""
    import ctypes as C
    from closestmolecule_b200.manufactured import p_exact, q_exact
    from closestmolecule_b200.solver import StructuredPQLinearSolver
    P1 = cm.FiniteElement("CG", "triangle", 1)
    mesh = cm.RectangleMesh(cm.Point(0, 0), cm.Point(1, 1), nx, ny)
    V = cm.FunctionSpace(mesh, cm.MixedElement([P1, P1]))
    bd = cm.MeshFunction("size_t", mesh, 1, 0)
    cm.CompiledSubDomain("near(x[0],0) && on_boundary").mark(bd, 31)
    form = cm.PQPrimalForm.manufactured(V, "supg")
    bcs = [cm.DirichletBC(V.sub(0), p_exact, "on_boundary"), cm.DirichletBC(V.sub(1), q_exact, bd, 31)]
    s = cm.NonlinearVariationalSolver(cm.NonlinearVariationalProblem(form, cm.Function(V), bcs))
    asm = s.asm
    X = mesh.coords
    u = torch.empty(2 * mesh.num_vertices(), dtype=torch.float64, device=mesh.device)
    u[0::2] = p_exact(X[:, 0], X[:, 1])
    u[1::2] = q_exact(X[:, 0], X[:, 1]) * (1 + 1e-3 * torch.sin(7 * X[:, 0] + 3 * X[:, 1]))
    bc = asm._bc_tables(bcs)
    lin = StructuredPQLinearSolver(asm, restart=40)
    blk = lin.blocks()
    assert blk is not None and asm.writes_blocks()
    nv = mesh.num_vertices()
    base = lin.ws.data_ptr()

    def view(p, n):
        off = p - base
        assert 0 <= off and off + 8 * n <= lin.ws.numel() and off % 8 == 0
        return lin.ws[off:off + 8 * n].view(torch.float64)

    names = ("PP", "PQ", "QP", "QQ", "sp", "sq")
    arrs = {k: view(getattr(blk, k), (7 if len(k) == 2 and k[0] in "PQ" and k[1] in "PQ" else 1) * nv) for k in names}
    for a in arrs.values():
        a.fill_(float("nan"))
    asm.assemble(u, True, blocks=blk)
    asm.apply_bcs(bc, u, True, blocks=blk)
    got = {k: a.clone() for k, a in arrs.items()}
    assert all(torch.isfinite(a).all() for a in got.values())
    lin.setup(blocks_ready=True)
    dx1 = torch.zeros_like(asm.b)
    assert lin.solve(asm.b, dx1, 1e-11, 0.0, 200)
    it1 = lin.last_iterations
    # the extraction path on the same CSR values
    for a in arrs.values():
        a.fill_(float("nan"))
    lin.setup()
    for k in names:
        assert torch.equal(got[k], arrs[k]), k
    dx2 = torch.zeros_like(asm.b)
    assert lin.solve(asm.b, dx2, 1e-11, 0.0, 200)
    assert lin.last_iterations == it1
    assert torch.equal(dx1, dx2)


@pytest.mark.parametrize("nx,ny", [(64, 32), (256, 128), (1024, 512)])
def test_kernel_generations_agree(nx, ny, monkeypatch):
    """The same Jacobian solved by the first-generation path (two-kernel sweeps, full residual / restriction /
    prolongation, host-driven GMRES with two reductions) and by the second-generation path, with the coarse
    levels as the shared-memory resident cycle (default), as one cluster kernel with global-memory phases, as one
    single-CTA kernel of that kind, and unfused: same iteration counts, same solution."""
    asm, make = _supg_solver(nx, ny)
    res = {}
    for name, env in (("v1", {"CM_SOLVER_V1": "1"}), ("v2", {}),
                      ("v2_cluster", {"CM_COARSE_SMEM": "0", "CM_COARSE_N1": "257"}),
                      ("v2_cta1", {"CM_COARSE_SMEM": "0", "CM_COARSE_N1": "65", "CM_COARSE_NCTA": "1"}),
                      ("v2_unfused", {"CM_COARSE_SMEM": "0"}), ("v2_nograph", {"CM_GRAPH": "0"}),
                      ("v2_nomid", {"CM_MID_MAX_N1": "0"}), ("v2_mid_unfused", {"CM_MID_FUSE": "0"}),
                      ("v2_f64_operator", {"CM_F32_OP": "0"}), ("v2_host_loop", {"CM_GRAPH_WHILE": "0"}),
                      ("v2_unscale_pass", {"CM_FUSE_UNSCALE": "0"})):
        for k in ("CM_SOLVER_V1", "CM_COARSE_NCTA", "CM_COARSE_N1", "CM_GRAPH", "CM_COARSE_SMEM", "CM_MID_MAX_N1",
                  "CM_MID_FUSE", "CM_F32_OP", "CM_GRAPH_WHILE", "CM_FUSE_UNSCALE"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        lin = make()
        info = lin.info()
        assert info["generation"] == (1 if name == "v1" else 2), info
        dx = torch.zeros_like(asm.b)
        for rep in range(2):          # the second solve replays the cached graphs
            lin.setup()
            assert lin.solve(asm.b, dx, 1e-11, 0.0, 200)
        res[name] = (lin.last_iterations, dx.cpu().numpy().copy(), lin.last_residual)
    ref_it, ref_x, _ = res["v1"]
    for name, (it, xx, rr) in res.items():
        assert abs(it - ref_it) <= 1, {k: v[0] for k, v in res.items()}
        assert np.linalg.norm(xx - ref_x) / np.linalg.norm(ref_x) < 1e-8, name
    assert res["v2"][0] == res["v2_cta1"][0] == res["v2_unfused"][0] == res["v2_nograph"][0] == res["v2_cluster"][0]
    assert np.linalg.norm(res["v2"][1] - res["v2_nograph"][1]) == 0.0      # graph replay is bit-identical


def test_gmres_restart_path(monkeypatch):
    """restart = 3 forces several cycles: x += M^{-1} V y, true residual, new cycle (all on the device)."""
    asm, make = _supg_solver(96, 48, restart=3)
    lin = make()
    lin.setup()
    dx = torch.zeros_like(asm.b)
    assert lin.solve(asm.b, dx, 1e-10, 0.0, 300)
    assert lin.last_iterations > 3
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    rp, col, vals = asm.csr()
    n = dx.numel()
    Amat = sp.csr_matrix((vals.cpu().numpy(), col.cpu().numpy(), rp.cpu().numpy()), shape=(n, n))
    ref = spla.splu(Amat.tocsc()).solve(asm.b.cpu().numpy())
    assert np.linalg.norm(dx.cpu().numpy() - ref) / np.linalg.norm(ref) < 1e-7
