"""DOLFIN XML mesh I/O: `Mesh("x.xml")`, `MeshFunction("size_t", mesh, "x_facet_region.xml")`
(advection_diffusion_mixed.py:313-314) and the writer side `File("x.xml") << mesh`
(create_flat_boundary.py:79-80).  Host-side plumbing (xml.etree), any triangle mesh."""
from __future__ import annotations

import xml.etree.ElementTree as ET

import numpy as np
import torch

from .mesh import Mesh, MeshFunction

_SENTINEL = np.iinfo(np.int64).max   # untagged facets of a MeshValueCollection file


def read_mesh(path, device=None):
    root = ET.parse(path).getroot()
    m = root.find("mesh")
    if m is None or m.get("celltype") != "triangle":
        raise ValueError("only DOLFIN XML triangle meshes are supported")
    vs = m.find("vertices")
    coords = np.zeros((int(vs.get("size")), 2))
    for v in vs:
        coords[int(v.get("index"))] = (float(v.get("x")), float(v.get("y")))
    cs = m.find("cells")
    cells = np.zeros((int(cs.get("size")), 3), dtype=np.int32)
    for c in cs:
        cells[int(c.get("index"))] = (int(c.get("v0")), int(c.get("v1")), int(c.get("v2")))
    cells.sort(axis=1)               # DOLFIN orders cell vertices by global index (B.2)
    return Mesh(coords, cells, None, None, device)


def write_mesh(mesh, path):
    X = mesh.coords.cpu().numpy()
    C = mesh.cells.cpu().numpy()
    with open(path, "w") as f:
        f.write('<?xml version="1.0"?>\n<dolfin xmlns:dolfin="http://fenicsproject.org">\n')
        f.write('  <mesh celltype="triangle" dim="2">\n')
        f.write(f'    <vertices size="{X.shape[0]}">\n')
        for i, (x, y) in enumerate(X):
            f.write(f'      <vertex index="{i}" x="{float(x)!r}" y="{float(y)!r}" />\n')
        f.write(f'    </vertices>\n    <cells size="{C.shape[0]}">\n')
        for i, (a, b, c) in enumerate(C):
            f.write(f'      <triangle index="{i}" v0="{a}" v1="{b}" v2="{c}" />\n')
        f.write('    </cells>\n  </mesh>\n</dolfin>\n')


def read_facet_function(mesh, path):
    """`*_facet_region.xml`: either a <mesh_value_collection> of (cell_index, local_entity, value)
    (dolfin-convert) or a <mesh_function> with one <entity> per facet (File << MeshFunction)."""
    root = ET.parse(path).getroot()
    mf = MeshFunction("size_t", mesh, 1, 0)
    fc = mesh.facets()
    vals = np.zeros(fc.num, dtype=np.int64)
    cells = mesh.cells.cpu().numpy().astype(np.int64)
    V = mesh.num_vertices()
    fv = fc.verts.cpu().numpy().astype(np.int64)
    key = fv[:, 0] * V + fv[:, 1]
    mvc = root.find(".//mesh_value_collection")
    if mvc is not None:
        vals[:] = _SENTINEL
        for e in mvc:
            c, le, v = int(e.get("cell_index")), int(e.get("local_entity")), int(e.get("value"))
            a, b = np.delete(cells[c], le)          # local facet le is opposite local vertex le
            k = min(a, b) * V + max(a, b)
            vals[np.searchsorted(key, k)] = v
    else:
        raise ValueError("facet file without a mesh_value_collection: use the (cell, local facet) format")
    mf.values = torch.as_tensor(vals, device=mesh.device)
    return mf


def write_facet_function(mf, path):
    mesh = mf.mesh
    fc = mesh.facets()
    vals = mf.values.cpu().numpy()
    c0 = fc.cell0.cpu().numpy()
    l0 = fc.loc0.cpu().numpy()
    tagged = np.nonzero((vals != 0) & (vals != _SENTINEL))[0]
    with open(path, "w") as f:
        f.write('<?xml version="1.0"?>\n<dolfin xmlns:dolfin="http://fenicsproject.org">\n  <mesh_function>\n')
        f.write(f'    <mesh_value_collection name="m" type="uint" dim="1" size="{tagged.size}">\n')
        for k in tagged:
            f.write(f'      <value cell_index="{c0[k]}" local_entity="{l0[k]}" value="{vals[k]}" />\n')
        f.write('    </mesh_value_collection>\n  </mesh_function>\n</dolfin>\n')
