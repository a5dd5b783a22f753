"""DOLFIN XML round trip (host logic, CPU): mesh + facet region files in the format of
create_flat_boundary.py:79-80 / dolfin-convert."""
import numpy as np
import torch

import closestmolecule_b200 as cm


def test_xml_round_trip(tmp_path):
    mesh = cm.RectangleMesh(cm.Point(0, 0.1), cm.Point(5, 5), 7, 5, device="cpu")
    mf = cm.flat_boundary_markers(mesh, 0.1, 5, 0, 5)
    cm.write_mesh(mesh, tmp_path / "m.xml")
    cm.write_facet_function(mf, tmp_path / "m_facet_region.xml")
    m2 = cm.read_mesh(tmp_path / "m.xml", device="cpu")
    assert not m2.structured
    assert torch.equal(m2.cells, mesh.cells) and torch.equal(m2.coords, mesh.coords)
    f2 = cm.read_facet_function(m2, tmp_path / "m_facet_region.xml")
    tagged = mf.values != 0
    assert torch.equal(f2.values[tagged], mf.values[tagged])
    assert (f2.values[~tagged] > 10 ** 9).all()          # untagged facets hold the sentinel, not 0
    for marker, count in ((21, 5), (22, 7), (23, 5), (24, 7)):
        assert int((f2.values == marker).sum()) == count
    # same P1 topology as the structured constructor
    t1, t2 = mesh.topology(), m2.topology()
    assert torch.equal(t1.nrowptr, t2.nrowptr) and torch.equal(t1.ncol, t2.ncol)
