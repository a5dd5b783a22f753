"""Meshes, facet markers and the P1 topology tables the kernels consume.

Host-side mirror of the FEniCS names the reference scripts use:
RectangleMesh / UnitSquareMesh (…MixedPrimal_convergence.py:74, test_for_paper_convergence.py:88,
create_flat_boundary.py:65), MeshFunction + CompiledSubDomain.mark (…MixedPrimal_convergence.py:75-80),
mesh.hmax() (:86).  All tables are torch tensors (any device); building them is setup plumbing, the
per-Newton-step work is in the CUDA library.
"""
from __future__ import annotations

import math
import re

import numpy as np
import torch

DOLFIN_EPS = 3.0e-16
# CUDA meshes build the P1 pattern with cm_build_p1_sparsity (env CM_LIBRARY_SPARSITY=0: torch sort/unique instead)
LIBRARY_SPARSITY = __import__("os").environ.get("CM_LIBRARY_SPARSITY", "1") != "0"


class Point:
    def __init__(self, x, y):
        self.x, self.y = float(x), float(y)


class Mesh:
    """coords [V,2] f64 (x=r2, y=r1), cells [C,3] int32, vertices ascending per cell."""

    def __init__(self, coords, cells, nx=None, ny=None, device=None):
        device = torch.device(device) if device is not None else _default_device()
        self.coords = torch.as_tensor(coords, dtype=torch.float64).to(device).contiguous()
        self.cells = torch.as_tensor(cells, dtype=torch.int32).to(device).contiguous()
        self.nx, self.ny = nx, ny
        self.device = device
        self._topo = {}

    # -- FEniCS-like accessors -------------------------------------------------------------
    def num_vertices(self):
        return self.coords.shape[0]

    def num_cells(self):
        return self.cells.shape[0]

    def hmax(self):
        X = self.coords[self.cells.long()]
        e = torch.stack([X[:, 1] - X[:, 0], X[:, 2] - X[:, 0], X[:, 2] - X[:, 1]], 1)
        return float(e.pow(2).sum(-1).sqrt().max())

    def ufl_cell(self):
        return "triangle"

    @property
    def structured(self):
        return self.nx is not None

    # -- topology ---------------------------------------------------------------------------
    def facets(self):
        if "facets" not in self._topo:
            self._topo["facets"] = _build_facets(self)
        return self._topo["facets"]

    def topology(self, with_interior_facets=False):
        key = ("topo", bool(with_interior_facets))
        if key not in self._topo:
            self._topo[key] = P1Topology(self, with_interior_facets)
        return self._topo[key]


def _default_device():
    return torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")


def RectangleMesh(p0, p1, nx, ny, diagonal="right", device=None):
    """DOLFIN RectangleMesh numbering: vertex iy*(nx+1)+ix at
    (a + ix*(b-a)/nx, c + iy*(d-c)/ny); quad (ix,iy) -> cells (v0,v1,v3), (v0,v2,v3)."""
    if diagonal != "right":
        raise NotImplementedError("only the default 'right' diagonal is used by the reference")
    device = torch.device(device) if device is not None else _default_device()
    a, c, b, d = p0.x, p0.y, p1.x, p1.y
    ix = torch.arange(nx + 1, dtype=torch.float64, device=device)
    iy = torch.arange(ny + 1, dtype=torch.float64, device=device)
    x = a + (ix * (b - a)) / nx
    y = c + (iy * (d - c)) / ny
    coords = torch.stack([x.repeat(ny + 1), y.repeat_interleave(nx + 1)], 1)
    cells = _structured_cells(nx, ny, device)
    return Mesh(coords, cells, nx, ny, device)


def UnitSquareMesh(nx, ny, device=None):
    return RectangleMesh(Point(0.0, 0.0), Point(1.0, 1.0), nx, ny, device=device)


def MappedRectangleMesh(nx, ny, r2_max, r1_max, sigma, gamma, device=None, grading=1.0):
    """Boundary-fitted structured stand-in for the FreeFem++ production mesh
    (meshes/automated_exp_meshing.edp:17-26; same labels as create_flat_boundary.py:70-77):
    r2 in [0,r2_max], r1 in [f(r2), r1_max], f = sigma*exp(-4/3*pi*gamma*r2^3).  grading > 1 clusters the grid
    lines towards r2 = 0 and the reaction boundary (r2 = r2_max*xi^g, r1 = f + eta^g*(r1_max - f)), where the
    solution has its exp(-4/3 pi c r2^3) layer."""
    m = RectangleMesh(Point(0.0, 0.0), Point(1.0, 1.0), nx, ny, device=device)
    xi = m.coords[:, 0].clone()
    eta = m.coords[:, 1].clone()
    x = r2_max * xi ** grading
    f = sigma * torch.exp(-4.0 / 3.0 * math.pi * gamma * x ** 3)
    m.coords[:, 0] = x
    m.coords[:, 1] = f + eta ** grading * (r1_max - f)
    return m


def _structured_cells(nx, ny, device):
    qx = torch.arange(nx, device=device, dtype=torch.int64).repeat(ny)
    qy = torch.arange(ny, device=device, dtype=torch.int64).repeat_interleave(nx)
    v0 = qy * (nx + 1) + qx
    v1, v2, v3 = v0 + 1, v0 + nx + 1, v0 + nx + 2
    cells = torch.empty((2 * nx * ny, 3), dtype=torch.int64, device=device)
    cells[0::2] = torch.stack([v0, v1, v3], 1)
    cells[1::2] = torch.stack([v0, v2, v3], 1)
    return cells.to(torch.int32)


# ---------------------------------------------------------------------------------------------


class Facets:
    """verts [F,2] ascending; cell0/cell1 (cell1=-1 on the boundary; cell0 = lower cell index = '+'
    side); loc0/loc1 local facet number (opposite local vertex)."""

    def __init__(self, verts, cell0, cell1, loc0, loc1):
        self.verts, self.cell0, self.cell1, self.loc0, self.loc1 = verts, cell0, cell1, loc0, loc1

    @property
    def num(self):
        return self.verts.shape[0]

    def exterior(self):
        return torch.nonzero(self.cell1 < 0).flatten()

    def interior(self):
        return torch.nonzero(self.cell1 >= 0).flatten()


def _build_facets(mesh):
    C = mesh.num_cells()
    V = mesh.num_vertices()
    c = mesh.cells.long()
    fa = torch.stack([c[:, 1], c[:, 0], c[:, 0]], 1).flatten()
    fb = torch.stack([c[:, 2], c[:, 2], c[:, 1]], 1).flatten()
    key = torch.minimum(fa, fb) * V + torch.maximum(fa, fb)
    # (cell, local) are already ascending in the flattened order -> a stable sort keeps cell0 < cell1
    key_s, order = torch.sort(key, stable=True)
    cid = order // 3
    loc = order % 3
    first = torch.ones_like(key_s, dtype=torch.bool)
    first[1:] = key_s[1:] != key_s[:-1]
    start = torch.nonzero(first).flatten()
    count = torch.diff(torch.cat([start, torch.tensor([key_s.numel()], device=start.device)]))
    F = start.numel()
    verts = torch.stack([key_s[start] // V, key_s[start] % V], 1).to(torch.int32)
    cell0 = cid[start].to(torch.int32)
    loc0 = loc[start].to(torch.int32)
    cell1 = torch.full((F,), -1, dtype=torch.int32, device=c.device)
    loc1 = torch.full((F,), -1, dtype=torch.int32, device=c.device)
    two = count == 2
    cell1[two] = cid[start[two] + 1].to(torch.int32)
    loc1[two] = loc[start[two] + 1].to(torch.int32)
    return Facets(verts, cell0, cell1, loc0, loc1)


def _library_sparsity(V, cells, extra):
    """(rowptr int32 [V+1], col int32 [nnz]) from cm_build_p1_sparsity."""
    import ctypes as C

    from . import _capi
    lib = _capi.lib()
    dev = cells.device
    ncells = cells.shape[0]
    nextra = 0 if extra is None else extra.shape[0]
    nbytes = int(lib.cm_p1_sparsity_workspace_bytes(V, ncells, nextra))
    if nbytes <= 0:
        raise _capi.CMError(lib.cm_last_error().decode())
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    cap = 9 * ncells + 2 * nextra
    rowptr = torch.empty(V + 1, dtype=torch.int32, device=dev)
    col = torch.empty(cap, dtype=torch.int32, device=dev)
    nnz = C.c_int64(0)
    _capi.check(lib.cm_build_p1_sparsity(V, ncells, _capi.ptr(cells), nextra, _capi.ptr(extra), _capi.ptr(rowptr),
                                         _capi.ptr(col), cap, C.byref(nnz), _capi.ptr(ws), nbytes,
                                         _capi.stream_ptr(dev)))
    return rowptr, col[:nnz.value].clone()


class P1Topology:
    """Node-level CSR pattern of the P1 space (DOLFIN SparsityPatternBuilder [ext]) and
    the row-owner gather tables of include/closestmol_b200.h."""

    def __init__(self, mesh, with_interior_facets=False):
        V = mesh.num_vertices()
        c = mesh.cells.long()
        dev = c.device
        o0 = o1 = None
        if with_interior_facets:
            # the macro element of an interior facet adds only the (opp0, opp1) couple: every other pair shares a cell
            fc = mesh.facets()
            fi = fc.interior()
            o0 = c[fc.cell0[fi].long(), fc.loc0[fi].long()]
            o1 = c[fc.cell1[fi].long(), fc.loc1[fi].long()]
        if c.is_cuda and LIBRARY_SPARSITY:
            # a9 through the library (cm_build_p1_sparsity: radix sort + unique of the row*V+col keys)
            extra = torch.stack([o0, o1], 1).to(torch.int32).contiguous() if with_interior_facets else None
            rp32, self.ncol = _library_sparsity(V, mesh.cells, extra)
            rp = rp32.long()
            deg = rp[1:] - rp[:-1]
            key = torch.repeat_interleave(torch.arange(V, device=dev), deg) * V + self.ncol.long()
        else:   # CPU tensors (host-logic tests): the same pattern with torch
            rows = c.repeat_interleave(3, dim=1).flatten()
            cols = c.repeat(1, 3).flatten()
            keys = [rows * V + cols]
            if with_interior_facets:
                keys += [o0 * V + o1, o1 * V + o0]
            key = torch.unique(torch.cat(keys))          # sorted
            r = key // V
            self.ncol = (key % V).to(torch.int32).contiguous()
            deg = torch.bincount(r, minlength=V)
            rp = torch.zeros(V + 1, dtype=torch.int64, device=dev)
            rp[1:] = torch.cumsum(deg, 0)
        assert int(rp[-1]) * 4 < 2 ** 31, "P-Q nnz exceeds int32 (DOLFIN la_index is 32 bit too)"
        self.nrowptr = rp.to(torch.int32).contiguous()
        self.maxdeg = int(deg.max())
        self.nnz = int(rp[-1])
        self.nv = V
        self.with_interior_facets = with_interior_facets
        self._key = key
        self._rp64 = rp
        # vertex -> incident cells (ascending cell index), entry = 4*cell + local
        vflat = c.flatten()
        _, order = torch.sort(vflat, stable=True)
        cid = order // 3
        loc = order % 3
        vs = vflat[order]
        self.v2c_ent = (4 * cid + loc).to(torch.int32).contiguous()
        cnt = torch.bincount(vflat, minlength=V)
        vp = torch.zeros(V + 1, dtype=torch.int64, device=dev)
        vp[1:] = torch.cumsum(cnt, 0)
        self.v2c_ptr = vp.to(torch.int32).contiguous()
        va = c[cid, (loc + 1) % 3]
        vb = c[cid, (loc + 2) % 3]
        k0 = self.slot(vs, vs)
        k1 = self.slot(vs, va)
        k2 = self.slot(vs, vb)
        assert self.maxdeg < 256
        self.v2c_slot = (k0 | (k1 << 8) | (k2 << 16)).to(torch.int32).contiguous()
        self.diag_slot = self.slot(torch.arange(V, device=dev), torch.arange(V, device=dev)).to(torch.int32)

    def slot(self, rows, cols):
        """Slot of column `cols` inside node row `rows` (entries must exist)."""
        k = rows.long() * self.nv + cols.long()
        pos = torch.searchsorted(self._key, k)
        return pos - self._rp64[rows.long()]

    # scalar CSR views (DOLFIN / PETSc AIJ layout) -------------------------------------------
    def scalar_csr(self, bs):
        """(rowptr, col) of the interleaved block-size-bs space: dof = bs*node + comp."""
        rp = self._rp64
        deg = rp[1:] - rp[:-1]
        V = self.nv
        dev = rp.device
        rowlen = (bs * deg).repeat_interleave(bs)
        srp = torch.zeros(bs * V + 1, dtype=torch.int64, device=dev)
        srp[1:] = torch.cumsum(rowlen, 0)
        # row 2v+c holds (w0,0),(w0,1),...: node-level cols expanded, each node row repeated bs times
        node_cols = (bs * self.ncol.long()[:, None] + torch.arange(bs, device=dev)[None, :]).flatten()
        # build index: for node v, block [bs*rp[v], bs*rp[v+1]) repeated bs times
        seg = torch.repeat_interleave(torch.arange(V, device=dev), bs * bs * deg)
        pos_in_node = torch.arange(seg.numel(), device=dev) - (bs * bs * rp[:-1])[seg]
        src = bs * rp[:-1][seg] + pos_in_node % (bs * deg[seg])
        return srp, node_cols[src]


# ---------------------------------------------------------------------------------------------
# facet markers


class MeshFunction:
    """MeshFunction("size_t", mesh, 1): one integer per facet (…MixedPrimal_convergence.py:75)."""

    def __init__(self, value_type, mesh, dim, value=0):
        assert dim == 1, "only facet functions are used by the reference"
        self.mesh = mesh
        self.values = torch.full((mesh.facets().num,), int(value), dtype=torch.int64, device=mesh.device)

    def set_all(self, v):
        self.values.fill_(int(v))

    def array(self):
        return self.values


_NEAR = re.compile(r"near\(\s*x\[(\d)\]\s*,\s*([-+0-9.eE]+)\s*\)")


class CompiledSubDomain:
    """CompiledSubDomain("near(x[0],1) && on_boundary") (…MixedPrimal_convergence.py:78-79).
    Supports the C++ subset the reference uses: near(x[i],c), &&, ||, parentheses, on_boundary.
    A python callable f(x[...,2]) -> bool tensor is accepted too (SubDomain.inside subclasses,
    create_flat_boundary.py:12-59)."""

    def __init__(self, expr):
        self.expr = expr

    def _eval(self, X, on_boundary):
        if callable(self.expr):
            return self.expr(X) & on_boundary
        s = self.expr

        def near_sub(m):
            return f"(torch.abs(X[:,{m.group(1)}]-({m.group(2)}))<{DOLFIN_EPS})"

        s = _NEAR.sub(near_sub, s).replace("&&", "&").replace("||", "|").replace("on_boundary", "OB")
        return eval(s, {"torch": torch, "X": X, "OB": on_boundary})  # noqa: S307 (trusted form text)

    def mark(self, mf, value):
        mesh = mf.mesh
        fc = mesh.facets()
        ob = fc.cell1 < 0
        A = mesh.coords[fc.verts[:, 0].long()]
        B = mesh.coords[fc.verts[:, 1].long()]
        M = 0.5 * (A + B)
        ok = self._eval(A, ob) & self._eval(B, ob) & self._eval(M, ob)
        mf.values[ok] = int(value)


def flat_boundary_markers(mesh, r1_min, r1_max, r2_min, r2_max):
    """create_flat_boundary.create_mesh (create_flat_boundary.py:62-77): right=21, top=22, left=23,
    bottom=24, marked in that order."""
    mf = MeshFunction("size_t", mesh, 1, 0)
    CompiledSubDomain(f"near(x[0],{r2_max}) && on_boundary").mark(mf, 21)
    CompiledSubDomain(f"near(x[1],{r1_max}) && on_boundary").mark(mf, 22)
    CompiledSubDomain(f"near(x[0],{r2_min}) && on_boundary").mark(mf, 23)
    CompiledSubDomain(f"near(x[1],{r1_min}) && on_boundary").mark(mf, 24)
    return mf


def structured_boundary_markers(mesh):
    """Markers 21/22/23/24 from the structured index (for mapped meshes whose bottom is curved)."""
    nx, ny = mesh.nx, mesh.ny
    mf = MeshFunction("size_t", mesh, 1, 0)
    fc = mesh.facets()
    a, b = fc.verts[:, 0].long(), fc.verts[:, 1].long()
    ax, ay, bx, by = a % (nx + 1), a // (nx + 1), b % (nx + 1), b // (nx + 1)
    ob = fc.cell1 < 0
    mf.values[ob & (ax == nx) & (bx == nx)] = 21
    mf.values[ob & (ay == ny) & (by == ny)] = 22
    mf.values[ob & (ax == 0) & (bx == 0)] = 23
    mf.values[ob & (ay == 0) & (by == 0)] = 24
    return mf


def facet_vertices(mesh, mf, marker):
    """Vertices of all facets carrying `marker` ('on_boundary' = every exterior facet)."""
    fc = mesh.facets()
    sel = (fc.cell1 < 0) if marker == "on_boundary" else (mf.values == int(marker))
    return torch.unique(fc.verts[sel].flatten().long())


def to_numpy(t):
    return t.detach().cpu().numpy()
