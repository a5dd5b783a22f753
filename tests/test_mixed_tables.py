"""Host tables of the RT1 x DG0 x CG1 path (closestmolecule_b200/mixed.py) against the oracle
(oracle/paper_m0.py): edge numbering, orientation signs, CSR pattern, and the per-slot gather lists that
replace DOLFIN's serial add_local (checked by emulating cm_csr_gather with torch)."""
import numpy as np
import scipy.sparse as sp
import torch

import closestmolecule_b200 as cm
from closestmolecule_b200.mixed import PaperMixedSystem
from oracle import paper_m0


def _system(n):
    mesh = cm.UnitSquareMesh(n, n, device="cpu")
    bd = cm.MeshFunction("size_t", mesh, 1, 0)
    cm.CompiledSubDomain("near(x[0],0) && on_boundary").mark(bd, 30)
    cm.CompiledSubDomain("near(x[0],1) && on_boundary").mark(bd, 31)
    cm.CompiledSubDomain("(near(x[1],0) || near(x[1],1)) && on_boundary").mark(bd, 32)
    return PaperMixedSystem(mesh, bd, tables_only=True)


def _emulate_gather(ptr, idx, src):
    out = torch.zeros(ptr.numel() - 1, dtype=torch.float64)
    seg = torch.repeat_interleave(torch.arange(ptr.numel() - 1), (ptr[1:] - ptr[:-1]).long())
    out.index_add_(0, seg, src[idx.long()])
    return out


def test_tables_match_oracle_and_gather_lists_assemble():
    n = 6
    S, O = _system(n), paper_m0.PaperM0(n)
    assert (S.E, S.C, S.V, S.N) == (O.E, O.C, O.V, O.N)
    assert np.array_equal(S.cell_edges.numpy(), O.cell_edges)
    assert np.array_equal(S.sign.numpy(), O.sign)
    assert np.array_equal(S.cell_dofs.numpy(), O.cell_dofs)
    assert np.array_equal(np.sort((S.bdry.values.numpy() == 30).nonzero()[0]), np.sort((O.markers == 30).nonzero()[0]))
    # the oracle's Jacobian pattern (cells + facet terms) is contained in ours, slot for slot
    J, _ = O.assemble(np.linspace(0.5, 1.5, O.N), True)
    A = sp.csr_matrix((np.ones(S.nnz), S.col.numpy(), S.rowptr.numpy()), shape=(S.N, S.N))
    assert (abs(J) > 0).multiply(A == 0).nnz == 0
    # gather lists: random element tensors, entry-major storage like the kernel writes them
    rng = np.random.default_rng(0)
    Je = rng.standard_normal((S.C, 49))
    Fe = rng.standard_normal((S.C, 7))
    vals = _emulate_gather(S.mat_ptr, S.mat_idx, torch.tensor(Je.T.copy().ravel()))
    cd = S.cell_dofs.numpy()
    ref = sp.coo_matrix((Je.ravel(), (np.repeat(cd, 7, axis=1).ravel(), np.tile(cd, (1, 7)).ravel())),
                        shape=(S.N, S.N)).tocsr()
    got = sp.csr_matrix((vals.numpy(), S.col.numpy(), S.rowptr.numpy()), shape=(S.N, S.N))
    assert abs(got - ref).max() < 1e-13
    b = _emulate_gather(S.vec_ptr, S.vec_idx, torch.tensor(Fe.T.copy().ravel()))
    bref = np.zeros(S.N)
    np.add.at(bref, cd.ravel(), Fe.ravel())
    assert np.max(np.abs(b.numpy() - bref)) < 1e-13
    # q-q facet operator map and diagonal slots
    rp, col = S.rowptr.numpy(), S.col.numpy()
    assert all(col[rp[i] + S.diag_slot_all[i]] == i for i in range(S.N))
    t = S.qtopo
    qrows = np.repeat(np.arange(S.V), np.diff(t.nrowptr.numpy()))
    rows_of_slot = np.repeat(np.arange(S.N), np.diff(rp))
    assert np.array_equal(rows_of_slot[S.kq_map.numpy()], S.E + S.C + qrows)
    assert np.array_equal(col[S.kq_map.numpy()], S.E + S.C + t.ncol.numpy())


def test_k1_tables_and_constant_facet_operator_match_oracle():
    """RT2 x DG1 x CG2 host tables of PaperMixedSystemK1 against oracle/paper_m1.py: dof numbering, the
    constant facet matrix / vector (tabulated with torch at setup) and the gather lists."""
    from closestmolecule_b200.mixed import PaperMixedSystemK1
    from oracle import paper_m1
    n = 4
    mesh = cm.UnitSquareMesh(n, n, device="cpu")
    bd = cm.MeshFunction("size_t", mesh, 1, 0)
    cm.CompiledSubDomain("near(x[0],0) && on_boundary").mark(bd, 30)
    cm.CompiledSubDomain("near(x[0],1) && on_boundary").mark(bd, 31)
    cm.CompiledSubDomain("(near(x[1],0) || near(x[1],1)) && on_boundary").mark(bd, 32)
    S, O = PaperMixedSystemK1(mesh, bd, tables_only=True), paper_m1.PaperM1(n)
    assert S.N == O.N == 3 * O.E + 5 * O.C + O.V
    assert np.array_equal(S.cell_dofs.numpy(), O.cell_dofs)
    assert np.array_equal(S.sign.numpy(), O.sign)
    # facet-only contribution of the oracle = full assembly minus the cell part: evaluate through linearity
    u = np.linspace(0.5, 1.5, O.N)
    rows, cols, vals = [], [], []
    b = np.zeros(O.N)
    O._exterior(u, b, rows, cols, vals, True)
    O._interior(u, b, rows, cols, vals, True)
    import scipy.sparse as sp
    Jf = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(O.N, O.N)).tocsr()
    Kf = sp.csr_matrix((S.Kf.numpy(), S.col.numpy(), S.rowptr.numpy()), shape=(S.N, S.N))
    assert abs(Kf - Jf).max() < 1e-12 * abs(Jf).max()
    assert np.max(np.abs(Kf @ u + S.cvec.numpy() - b)) < 1e-12 * np.max(np.abs(b))
    # gather lists
    rng = np.random.default_rng(0)
    Je = rng.standard_normal((S.C, 289))
    got = _emulate_gather(S.mat_ptr, S.mat_idx, torch.tensor(Je.T.copy().ravel()))
    cd = S.cell_dofs.numpy()
    ref = sp.coo_matrix((Je.ravel(), (np.repeat(cd, 17, axis=1).ravel(), np.tile(cd, (1, 17)).ravel())),
                        shape=(S.N, S.N)).tocsr()
    assert abs(sp.csr_matrix((got.numpy(), S.col.numpy(), S.rowptr.numpy()), shape=(S.N, S.N)) - ref).max() < 1e-13


def test_c0ip_variant_facet_operator_matches_oracle_and_oracle_converges():
    """...MixedPrimal_convergence_C0IP.py variant: the host tabulation of its facet terms (inflow term on x = 0 only,
    penalty with avg CellDiameter and the script's minus sign, natural p data on the whole boundary) against
    oracle/paper_m1.MixedPrimalC0IPM1; the oracle's own table keeps e_0(p) at second order."""
    from closestmolecule_b200.mixed import PaperMixedSystemK1
    from oracle import paper_m1
    import scipy.sparse as sp
    n = 4
    mesh = cm.UnitSquareMesh(n, n, device="cpu")
    bd = cm.MeshFunction("size_t", mesh, 1, 0)
    cm.CompiledSubDomain("near(x[0],0) && on_boundary").mark(bd, 31)
    cm.CompiledSubDomain("near(x[0],1) && on_boundary").mark(bd, 32)
    cm.CompiledSubDomain("(near(x[1],0) || near(x[1],1)) && on_boundary").mark(bd, 33)
    S = PaperMixedSystemK1(mesh, bd, D1=0.01, D2=0.015, stab=-0.00075, stab2=0.0, degree=4, tables_only=True,
                           variant="c0ip")
    O = paper_m1.MixedPrimalC0IPM1(n)
    assert S.gamma == O.gamma < 0
    u = np.linspace(0.5, 1.5, O.N)
    rows, cols, vals = [], [], []
    b = np.zeros(O.N)
    O._exterior(u, b, rows, cols, vals, True)
    O._interior(u, b, rows, cols, vals, True)
    Jf = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(O.N, O.N)).tocsr()
    Kf = sp.csr_matrix((S.Kf.numpy(), S.col.numpy(), S.rowptr.numpy()), shape=(S.N, S.N))
    assert abs(Kf - Jf).max() < 1e-12 * abs(Jf).max()
    assert np.max(np.abs(Kf @ u + S.cvec.numpy() - b)) < 1e-12 * np.max(np.abs(b))
    rows = paper_m1.mixed_primal_c0ip_table(3)
    assert [int(r.split()[0]) for r in rows] == [3 * m.E + 5 * m.C + m.V for m in
                                                 (paper_m1.MixedPrimalC0IPM1(2 ** (k + 1)) for k in range(3))]
    assert 1.8 < float(rows[-1].replace("&", " ").split()[5]) < 2.2


def test_c0ip_other_bcs_variant_facet_operator_matches_oracle():
    """..._C0IP_otherBCs.py variant (r2vec = (-1,0) in the q-equation, penalty +stab with avg CellDiameter, natural p
    data on right / rest): host tabulation of the facet terms against oracle/paper_m1.MixedPrimalC0IPOtherM1."""
    from closestmolecule_b200.mixed import PaperMixedSystemK1
    from oracle import paper_m1
    import scipy.sparse as sp
    n = 4
    mesh = cm.UnitSquareMesh(n, n, device="cpu")
    bd = cm.MeshFunction("size_t", mesh, 1, 0)
    cm.CompiledSubDomain("near(x[0],0) && on_boundary").mark(bd, 30)
    cm.CompiledSubDomain("near(x[0],1) && on_boundary").mark(bd, 31)
    cm.CompiledSubDomain("(near(x[1],0) || near(x[1],1)) && on_boundary").mark(bd, 32)
    S = PaperMixedSystemK1(mesh, bd, D1=2.0, D2=1.5, stab=0.0075, stab2=0.0, degree=4, tables_only=True,
                           variant="c0ip_other")
    O = paper_m1.MixedPrimalC0IPOtherM1(n)
    assert S.gamma == O.gamma > 0
    u = np.linspace(0.5, 1.5, O.N)
    rows, cols, vals = [], [], []
    b = np.zeros(O.N)
    O._exterior(u, b, rows, cols, vals, True)
    O._interior(u, b, rows, cols, vals, True)
    Jf = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(O.N, O.N)).tocsr()
    Kf = sp.csr_matrix((S.Kf.numpy(), S.col.numpy(), S.rowptr.numpy()), shape=(S.N, S.N))
    assert abs(Kf - Jf).max() < 1e-12 * abs(Jf).max()
    assert np.max(np.abs(Kf @ u + S.cvec.numpy() - b)) < 1e-12 * np.max(np.abs(b))
    f3 = S.data["f3"](torch.tensor([0.3], dtype=torch.float64), torch.tensor([0.7], dtype=torch.float64))
    assert abs(float(f3) - float(O.sym["f3"](0.3, 0.7))) < 1e-14
    rows = paper_m1.mixed_primal_c0ip_table(3, other_bcs=True)
    assert 1.8 < float(rows[-1].replace("&", " ").split()[5]) < 2.0
