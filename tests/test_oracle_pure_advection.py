"""oracle/pure_advection.py (scalar CG2 restatement of the two pure-advection scripts): consistency of the forms with
the manufactured solution, the orders of convergence, and agreement with the q-q block of oracle/paper_m1.py's
17 x 17 element tensors (the block the GPU path reuses)."""
import numpy as np
import pytest

from oracle import paper_m1, pure_advection as pa


def _interpolant(m):
    X, fc = m.mesh.coords, m.fc
    XM = 0.5 * (X[fc.verts[:, 0]] + X[fc.verts[:, 1]])
    return np.concatenate([pa.q_ex(X[:, 0], X[:, 1]), pa.q_ex(XM[:, 0], XM[:, 1])])


@pytest.mark.parametrize("kind", ["supg", "c0ip"])
def test_forms_are_consistent_with_the_manufactured_solution(kind):
    res = []
    for n in (4, 8, 16):
        m = pa.PureAdvection(n, kind)
        A, b = m.assemble()
        free = np.setdiff1d(np.arange(m.N), m.bcs()[0])
        res.append(np.abs((A @ _interpolant(m) - b)[free]).max() / np.abs(b).max())
    assert res[2] < res[1] / 3 < res[0] / 9          # the residual of the interpolant vanishes under refinement


def test_supg_table_has_second_order_and_dof_column():
    rows = pa.table("supg", 5)
    assert [int(r.split()[0]) for r in rows] == [25, 81, 289, 1089, 4225]       # (2n+1)^2
    assert 1.9 < float(rows[-1].replace("&", " ").split()[3]) < 2.1


def test_c0ip_sign_of_the_penalty():
    """The script's +0.005 works against the negative form (errors grow under refinement); with the sign of
    ..._MixedPrimal_convergence_C0IP.py:147 the same form converges."""
    err = lambda stab: [m.error(m.solve()) for m in (pa.PureAdvection(n, "c0ip", stab=stab) for n in (8, 16, 32))]
    plus, minus = err(0.005), err(-0.005)
    assert plus[2] > 100 * plus[0]
    assert minus[2] < minus[1] / 2.5 < minus[0] / 6


@pytest.mark.parametrize("kind", ["supg", "c0ip"])
def test_scalar_matrix_is_the_qq_block_of_the_mixed_element_tensors(kind):
    n = 5
    m = pa.PureAdvection(n, kind)
    A, b = m.assemble()
    if kind == "supg":
        O = paper_m1.MixedPrimalM1(n, supg=True)
        O.supg_stab = pa.STAB["supg"]
        O.nat_markers = None
    else:
        O = paper_m1.MixedPrimalC0IPM1(n, stab=-pa.STAB["c0ip"])
        O.q_react, O.q_x2p = 1.0, 0.0
        O.k_adv = np.isin(O.markers[O.ext], (31, 33)).astype(float)
    J, F0 = O.assemble(np.zeros(O.N), True)
    Jqq = J[O.q0:, O.q0:]
    assert abs(Jqq - A).max() < 1e-12 * abs(A).max()
    if kind == "supg":          # same forcing q_ex + dq_ex/dx + r2^2 q_ex
        assert np.abs(-F0[O.q0:] - b).max() < 1e-12 * np.abs(b).max()


@pytest.mark.parametrize("kind", ["supg", "c0ip"])
def test_host_tables_of_the_gpu_port_match_the_oracle_pattern(kind):
    """PureAdvectionCG2 cuts the q-q block out of the RT2 x DG1 x CG2 pattern: its CSR pattern must contain the scalar
    oracle's matrix pattern, diag_slot must point at the diagonal INSIDE each row (what cm_apply_dirichlet expects),
    and the boundary dofs / values must be the oracle's."""
    import torch
    import closestmolecule_b200 as cm
    from closestmolecule_b200.mixed import PureAdvectionCG2
    n = 4
    mesh = cm.UnitSquareMesh(n, n, device="cpu")
    bd = cm.MeshFunction("size_t", mesh, 1, 0)
    cm.CompiledSubDomain("near(x[0],0) && on_boundary").mark(bd, 31)
    cm.CompiledSubDomain("near(x[0],1) && on_boundary").mark(bd, 32)
    cm.CompiledSubDomain("(near(x[1],0) || near(x[1],1)) && on_boundary").mark(bd, 33)
    S = PureAdvectionCG2(mesh, bd, kind, tables_only=True)
    O = pa.PureAdvection(n, kind)
    assert S.N == O.N
    rp, col = S.rowptr.numpy().astype(np.int64), S.col.numpy().astype(np.int64)
    assert rp[0] == 0 and rp[-1] == col.size and np.all(np.diff(rp) > 0)
    rows = np.repeat(np.arange(S.N), np.diff(rp))
    assert np.all(np.diff(rows * S.N + col) > 0)                         # sorted, unique
    ds = S.diag_slot.numpy().astype(np.int64)
    assert np.array_equal(col[rp[:-1] + ds], np.arange(S.N))
    A, _ = O.assemble()
    A = A.tocoo()
    have = set((rows * S.N + col).tolist())
    nz = np.abs(A.data) > 0
    assert set((A.row[nz].astype(np.int64) * S.N + A.col[nz]).tolist()) <= have
    d, g = O.bcs()
    assert np.array_equal(S.bc_dofs.numpy(), d) and np.allclose(S.bc_vals.numpy(), g, rtol=0, atol=1e-15)
    assert np.array_equal(S.bc_diag.numpy(), ds[d])
