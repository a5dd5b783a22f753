"""Oracle building blocks: DOLFIN-numbered meshes, FIAT quadrature, facet
topology, DOLFIN-pattern CSR and serial-order insertion.  (Test infrastructure;
see oracle/__init__.py.)

Conventions restated from FEniCS 2019.1 [ext] as listed in SURVEY.md App. B;
the reference call sites that trigger them are cited per function.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

DOLFIN_EPS = 3.0e-16

# ----------------------------------------------------------------------------------------------
# meshes


class Mesh:
    """coords [V,2] (x=r2, y=r1), cells [C,3] int32 with ascending vertex ids per cell."""

    def __init__(self, coords, cells, nx=None, ny=None):
        self.coords = np.ascontiguousarray(coords, dtype=np.float64)
        self.cells = np.ascontiguousarray(cells, dtype=np.int32)
        self.nx, self.ny = nx, ny
        self._facets = None

    @property
    def num_vertices(self):
        return self.coords.shape[0]

    @property
    def num_cells(self):
        return self.cells.shape[0]

    def hmax(self):
        """mesh.hmax() (e.g. SphericalAdvectionDiffusionReaction_convergence.py:68): largest
        vertex distance over cells."""
        X = self.coords[self.cells]
        e = np.stack([X[:, 1] - X[:, 0], X[:, 2] - X[:, 0], X[:, 2] - X[:, 1]], 1)
        return float(np.sqrt((e ** 2).sum(-1)).max())

    def facets(self):
        if self._facets is None:
            self._facets = build_facets(self)
        return self._facets


def rectangle_mesh(a, c, b, d, nx, ny):
    """RectangleMesh(Point(a,c),Point(b,d),nx,ny), diagonal "right"
    (create_flat_boundary.py:65; SphericalAdvectionDiffusionReaction_convergence.py:66).
    Vertex iy*(nx+1)+ix at (a + ix*(b-a)/nx, c + iy*(d-c)/ny); quad (ix,iy) -> cells
    (v0,v1,v3), (v0,v2,v3)."""
    ix = np.arange(nx + 1, dtype=np.float64)
    iy = np.arange(ny + 1, dtype=np.float64)
    x = a + (ix * (b - a)) / nx
    y = c + (iy * (d - c)) / ny
    X, Y = np.meshgrid(x, y, indexing="xy")
    coords = np.stack([X.ravel(), Y.ravel()], 1)
    qx, qy = np.meshgrid(np.arange(nx), np.arange(ny), indexing="xy")
    v0 = (qy * (nx + 1) + qx).ravel()
    v1 = v0 + 1
    v2 = v0 + nx + 1
    v3 = v1 + nx + 1
    cells = np.empty((2 * nx * ny, 3), dtype=np.int32)
    cells[0::2] = np.stack([v0, v1, v3], 1)
    cells[1::2] = np.stack([v0, v2, v3], 1)
    return Mesh(coords, cells, nx, ny)


def unit_square_mesh(nx, ny):
    """UnitSquareMesh (test_for_paper_convergence.py:88)."""
    return rectangle_mesh(0.0, 0.0, 1.0, 1.0, nx, ny)


def mapped_rectangle_mesh(nx, ny, r2_max, r1_max, sigma, gamma, grading=1.0):
    """Boundary-fitted structured stand-in for the FreeFem++ production mesh
    (meshes/automated_exp_meshing.edp:17-26): r2 in [0,r2_max], r1 in [f(r2), r1_max] with
    f = sigma*exp(-4/3*pi*gamma*r2^3); same topology/numbering as rectangle_mesh.  grading > 1 clusters the grid
    lines towards r2 = 0 and the reaction boundary."""
    m = rectangle_mesh(0.0, 0.0, 1.0, 1.0, nx, ny)
    xi = m.coords[:, 0].copy()
    eta = m.coords[:, 1].copy()
    x = r2_max * xi ** grading
    f = sigma * np.exp(-4.0 / 3.0 * np.pi * gamma * x ** 3)
    m.coords[:, 0] = x
    m.coords[:, 1] = f + eta ** grading * (r1_max - f)
    return m


# ----------------------------------------------------------------------------------------------
# quadrature (FIAT create_quadrature "default") -- constants as stored (15 digits)

_TRI = {
    1: (np.array([[1.0 / 3.0, 1.0 / 3.0]]), np.array([0.5])),
    2: (np.array([[1.0 / 6.0, 1.0 / 6.0], [1.0 / 6.0, 2.0 / 3.0], [2.0 / 3.0, 1.0 / 6.0]]),
        np.array([1.0 / 6.0, 1.0 / 6.0, 1.0 / 6.0])),
    4: (np.array([[0.816847572980459, 0.091576213509771],
                  [0.091576213509771, 0.816847572980459],
                  [0.091576213509771, 0.091576213509771],
                  [0.108103018168070, 0.445948490915965],
                  [0.445948490915965, 0.108103018168070],
                  [0.445948490915965, 0.445948490915965]]),
        np.array([0.109951743655322 / 2.0] * 3 + [0.223381589678011 / 2.0] * 3)),
    6: (np.array([[0.873821971016996, 0.063089014491502],
                  [0.063089014491502, 0.873821971016996],
                  [0.063089014491502, 0.063089014491502],
                  [0.501426509658179, 0.249286745170910],
                  [0.249286745170910, 0.501426509658179],
                  [0.249286745170910, 0.249286745170910],
                  [0.636502499121399, 0.310352451033785],
                  [0.636502499121399, 0.053145049844816],
                  [0.310352451033785, 0.636502499121399],
                  [0.310352451033785, 0.053145049844816],
                  [0.053145049844816, 0.636502499121399],
                  [0.053145049844816, 0.310352451033785]]),
        np.array([0.050844906370207 / 2.0] * 3 + [0.116786275726379 / 2.0] * 3
                 + [0.082851075618374 / 2.0] * 6)),
}


def _gauss_jacobi_tri(degree):
    """FIAT collapsed Gauss-Jacobi rule for degree >= 7 (B.4): m=(degree+2)//2 points/direction."""
    from scipy.special import roots_jacobi
    m = (degree + 2) // 2
    ptx, wx = roots_jacobi(m, 0.0, 0.0)
    pty, wy = roots_jacobi(m, 1.0, 0.0)
    pts, wts = [], []
    for i in range(m):
        for j in range(m):
            # collapsed point on the (-1,-1),(1,-1),(-1,1) triangle, then to the UFC triangle
            xr = 0.5 * (1.0 + ptx[i]) * (1.0 - pty[j]) - 1.0
            yr = pty[j]
            pts.append([0.5 * (xr + 1.0), 0.5 * (yr + 1.0)])
            wts.append(0.5 * wx[i] * wy[j] * 0.25)
    return np.array(pts), np.array(wts)


def triangle_rule(degree):
    """(points [nq,2] in (xi,eta), weights summing to 1/2)."""
    if degree <= 1:
        return _TRI[1]
    if degree == 2:
        return _TRI[2]
    if degree in (3, 4):
        # degree 3 has its own 6-point rule in FIAT; no reference script uses it
        return _TRI[4]
    if degree in (5, 6):
        return _TRI[6]
    return _gauss_jacobi_tri(degree)


def facet_rule(degree):
    """Gauss-Legendre on [0,1] with (degree+2)//2 points (B.4): (s, w) with sum w = 1."""
    m = (degree + 2) // 2
    p, w = np.polynomial.legendre.leggauss(m)
    return 0.5 * (p + 1.0), 0.5 * w


# ----------------------------------------------------------------------------------------------
# geometry


def cell_geometry(mesh):
    """Per cell: detJ (signed), grads [C,3,2] of the P1 basis, vertex coords [C,3,2]
    (UFC reference triangle)."""
    X = mesh.coords[mesh.cells]
    x0, x1, x2 = X[:, 0], X[:, 1], X[:, 2]
    j00 = x1[:, 0] - x0[:, 0]
    j01 = x2[:, 0] - x0[:, 0]
    j10 = x1[:, 1] - x0[:, 1]
    j11 = x2[:, 1] - x0[:, 1]
    det = j00 * j11 - j01 * j10
    g1 = np.stack([j11, -j01], 1) / det[:, None]
    g2 = np.stack([-j10, j00], 1) / det[:, None]
    g0 = -(g1 + g2)
    grads = np.stack([g0, g1, g2], 1)
    return det, grads, X


def cell_diameter(mesh):
    """CellDiameter = longest edge for triangles (…_SUPG.py:81)."""
    X = mesh.coords[mesh.cells]
    e = np.stack([X[:, 1] - X[:, 0], X[:, 2] - X[:, 0], X[:, 2] - X[:, 1]], 1)
    return np.sqrt((e ** 2).sum(-1)).max(1)


# ----------------------------------------------------------------------------------------------
# facets


class Facets:
    """Edges of the mesh.  verts [F,2] ascending; cell0/cell1 (cell1=-1 on the boundary; cell0 is the
    lower cell index = the '+' side, B.2); loc0/loc1 = local facet number (opposite local vertex)."""

    def __init__(self, verts, cell0, cell1, loc0, loc1):
        self.verts, self.cell0, self.cell1, self.loc0, self.loc1 = verts, cell0, cell1, loc0, loc1

    @property
    def interior(self):
        return np.nonzero(self.cell1 >= 0)[0]

    @property
    def exterior(self):
        return np.nonzero(self.cell1 < 0)[0]


def build_facets(mesh):
    C = mesh.num_cells
    V = mesh.num_vertices
    cells = mesh.cells.astype(np.int64)
    # local facet i is opposite local vertex i: (v1,v2), (v0,v2), (v0,v1)
    fa = np.stack([cells[:, 1], cells[:, 0], cells[:, 0]], 1).ravel()
    fb = np.stack([cells[:, 2], cells[:, 2], cells[:, 1]], 1).ravel()
    key = np.minimum(fa, fb) * V + np.maximum(fa, fb)
    cid = np.repeat(np.arange(C, dtype=np.int64), 3)
    loc = np.tile(np.arange(3, dtype=np.int64), C)
    order = np.lexsort((cid, key))
    key_s, cid_s, loc_s = key[order], cid[order], loc[order]
    first = np.ones(key_s.size, dtype=bool)
    first[1:] = key_s[1:] != key_s[:-1]
    start = np.nonzero(first)[0]
    count = np.diff(np.append(start, key_s.size))
    F = start.size
    verts = np.stack([key_s[start] // V, key_s[start] % V], 1).astype(np.int32)
    cell0 = cid_s[start].astype(np.int32)
    loc0 = loc_s[start].astype(np.int32)
    cell1 = np.full(F, -1, dtype=np.int32)
    loc1 = np.full(F, -1, dtype=np.int32)
    two = count == 2
    cell1[two] = cid_s[start[two] + 1]
    loc1[two] = loc_s[start[two] + 1]
    return Facets(verts, cell0, cell1, loc0, loc1)


def mark_facets(mesh, predicates, default=0):
    """MeshFunction("size_t",mesh,1); set_all(default); then CompiledSubDomain(...).mark(bdry,id) in
    order (…MixedPrimal_convergence.py:75-80; create_flat_boundary.py:66-77).  ``predicates`` is a
    list of (callable(x[...,2]) -> bool, id); a facet is marked when both vertices and the midpoint
    satisfy the predicate.  All predicates here include ``on_boundary``."""
    fc = mesh.facets()
    markers = np.full(fc.verts.shape[0], default, dtype=np.int64)
    ext = fc.cell1 < 0
    A = mesh.coords[fc.verts[:, 0]]
    B = mesh.coords[fc.verts[:, 1]]
    M = 0.5 * (A + B)
    for pred, mid in predicates:
        ok = ext & pred(A) & pred(B) & pred(M)
        markers[ok] = mid
    return markers


def near(a, b, eps=DOLFIN_EPS):
    return np.abs(a - b) < eps


def facet_geometry(mesh, fidx):
    """For facets fidx: length, outward unit normal of cell0 [n,2], end points."""
    fc = mesh.facets()
    A = mesh.coords[fc.verts[fidx, 0]]
    B = mesh.coords[fc.verts[fidx, 1]]
    t = B - A
    L = np.sqrt((t ** 2).sum(1))
    nrm = np.stack([t[:, 1], -t[:, 0]], 1) / L[:, None]
    # orient outward from cell0: opposite vertex is on the inside
    c0 = mesh.cells[fc.cell0[fidx]]
    opp = mesh.coords[c0[np.arange(fidx.size), fc.loc0[fidx]]]
    flip = ((opp - A) * nrm).sum(1) > 0
    nrm[flip] *= -1.0
    return L, nrm, A, B


# ----------------------------------------------------------------------------------------------
# DOLFIN-pattern CSR and serial-order insertion


def p1_node_pattern(mesh, with_interior_facets=False):
    """Node-level CSR (rowptr, col) of the P1 pattern: union over cells of vertex x vertex, plus --
    when the form has dS integrals -- union over interior facets of the 4 macro vertices (DOLFIN
    SparsityPatternBuilder [ext]).  Columns ascending per row."""
    V = mesh.num_vertices
    c = mesh.cells.astype(np.int64)
    rows = [np.repeat(c, 3, axis=1).ravel()]
    cols = [np.tile(c, (1, 3)).ravel()]
    if with_interior_facets:
        fc = mesh.facets()
        fi = fc.interior
        macro = np.concatenate([c[fc.cell0[fi]], c[fc.cell1[fi]]], 1)  # [Fi,6]
        rows.append(np.repeat(macro, 6, axis=1).ravel())
        cols.append(np.tile(macro, (1, 6)).ravel())
    r = np.concatenate(rows)
    cc = np.concatenate(cols)
    A = sp.coo_matrix((np.ones(r.size, dtype=np.int8), (r, cc)), shape=(V, V)).tocsr()
    A.sum_duplicates()
    A.sort_indices()
    return A.indptr.astype(np.int32), A.indices.astype(np.int32)


def blocked_pattern(rowptr, col, bs):
    """Scalar CSR of the interleaved block-size-bs space (dof = bs*node + comp)."""
    V = rowptr.size - 1
    deg = np.diff(rowptr).astype(np.int64)
    rp = np.zeros(bs * V + 1, dtype=np.int64)
    rp[1:] = np.cumsum(np.repeat(deg * bs, bs))
    cols = (bs * col.astype(np.int64)[:, None] + np.arange(bs)[None, :]).reshape(-1)  # per node-nnz
    # each node row is replicated bs times
    out = np.empty(rp[-1], dtype=np.int64)
    # build with a python-free trick: for node v the bs rows all equal cols[bs*rowptr[v]:bs*rowptr[v+1]]
    seg_len = deg * bs
    src_start = rowptr[:-1].astype(np.int64) * bs
    for comp in range(bs):
        dst_start = rp[comp:-1:bs]
        idx = _ranges(dst_start, seg_len)
        out[idx] = cols[_ranges(src_start, seg_len)]
    return rp.astype(np.int32) if rp[-1] < 2 ** 31 else rp, out.astype(np.int32)


def _ranges(starts, lens):
    """concatenate [s, s+l) for all (s,l)."""
    total = int(lens.sum())
    if total == 0:
        return np.zeros(0, dtype=np.int64)
    ends = np.cumsum(lens)
    base = np.repeat(starts - (ends - lens), lens)
    return base + np.arange(total, dtype=np.int64)


def csr_slots(rowptr, col, rows, cols):
    """Position in the CSR value array of every (row, col) pair (must exist)."""
    ncols = int(col.max()) + 1 if col.size else 1
    key = rows.astype(np.int64) * ncols + cols.astype(np.int64)
    deg = np.diff(rowptr)
    rid = np.repeat(np.arange(rowptr.size - 1, dtype=np.int64), deg)
    allkey = rid * ncols + col.astype(np.int64)
    pos = np.searchsorted(allkey, key)
    assert np.all(allkey[pos] == key), "entry outside the sparsity pattern"
    return pos


def add_serial(vals, slots, contrib):
    """vals[slots] += contrib one entry at a time in the given order (np.add.at is sequential) --
    the floating-point summation order of DOLFIN's serial loop + MatSetValues(ADD) [ext] (B.5)."""
    np.add.at(vals, slots.ravel(), contrib.ravel())
