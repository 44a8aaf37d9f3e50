"""Field output: .vtu / .pvd files of GLL fields, written on the host."""
import os

import numpy as np
import pytest

from conftest import warp_vertices


@pytest.mark.parametrize("P,binary", [(1, True), (2, False), (4, True)])
def test_vtu_round_trip_and_geometry(fus, tmp_path, P, binary):
    from fenicsx_fus_b200 import vtkio
    m = fus.BoxMesh((3, 2, 2), (0, 0, 0), (1.5, 1.0, 0.5), warp=lambda x: warp_vertices(x, 0.05, 2))
    V = fus.FunctionSpace(m, P, numbering=1)
    X = V.tabulate_dof_coordinates()
    u = np.sin(3 * X[:, 0]) * np.cos(2 * X[:, 1]) + X[:, 2]
    v = X[:, 0] - 2 * X[:, 1]
    path = str(tmp_path / "u.vtu")
    npts, ncell = vtkio.write_vtu(path, V, {"u": u, "v": v}, binary=binary)
    assert (npts, ncell) == (V.ndofs, m.ncells * P ** 3)
    d = vtkio.read_vtu_arrays(path)
    assert np.array_equal(d["u"], u) and np.array_equal(d["v"], v)          # values kept exactly
    assert np.array_equal(d["Points"], X)
    conn = d["connectivity"].reshape(-1, 8)
    assert conn.min() == 0 and conn.max() == V.ndofs - 1
    assert np.array_equal(d["offsets"], 8 * np.arange(1, ncell + 1)) and (d["types"] == 12).all()
    # every sub-hexahedron is positively oriented in VTK vertex order and together they tile the mesh
    Y = X[conn]
    e1, e2, e3 = Y[:, 1] - Y[:, 0], Y[:, 3] - Y[:, 0], Y[:, 4] - Y[:, 0]
    assert (np.einsum("ij,ij->i", np.cross(e1, e2), e3) > 0).all()
    vol = 0.0
    for tet in ((0, 1, 3, 4), (1, 2, 3, 6), (1, 3, 4, 6), (1, 4, 5, 6), (3, 4, 6, 7)):   # 5-tet split
        a, b, c, dd = (Y[:, k] for k in tet)
        vol += np.abs(np.einsum("ij,ij->i", np.cross(b - a, c - a), dd - a)).sum() / 6.0
    # exact for planar faces; warped trilinear faces differ at second order in the warp
    assert abs(vol - 1.5 * 1.0 * 0.5) < 2e-2
    # each interior lattice node is shared by neighbouring sub-cells: all dofs are referenced
    assert np.unique(conn).size == V.ndofs


def test_pvd_time_series_on_the_reference_mesh(fus, tmp_path):
    """Snapshots every few steps, as the reference's examples do with VTXWriter::write(t), on the
    reference's unstructured test mesh (general hexahedra, conforming numbering)."""
    import xml.etree.ElementTree as ET
    from fenicsx_fus_b200 import vtkio
    from fenicsx_fus_b200.unstructured import HexFunctionSpace, HexMesh
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_mesh_hex6312.npz"))
    m = HexMesh(g["geometry"], g["topology_vtk"][:, (0, 1, 3, 2, 4, 5, 7, 6)], g["facet_quads"],
                g["facet_values"])
    V = HexFunctionSpace(m, 2)
    X = V.tabulate_dof_coordinates()
    w = vtkio.TimeSeriesWriter(str(tmp_path / "run.pvd"), V)
    for k in range(3):
        w.write(0.5 * k, u=np.cos(k + X[:, 0]), v=X[:, 1] * k)
    w.close()
    sets = list(ET.parse(str(tmp_path / "run.pvd")).getroot().iter("DataSet"))
    assert [float(s.get("timestep")) for s in sets] == [0.0, 0.5, 1.0]
    last = vtkio.read_vtu_arrays(str(tmp_path / sets[-1].get("file")))
    assert np.array_equal(last["u"], np.cos(2 + X[:, 0]))
    assert last["connectivity"].size == 8 * m.ncells * 8                  # P^3 = 8 sub-cells per cell
    with pytest.raises(ValueError):
        vtkio.write_vtu(str(tmp_path / "bad.vtu"), V, {"u": np.zeros(3)})


def test_vtu_2d_on_the_reference_example_mesh(fus, tmp_path):
    """The reference writes its output in the 2-D examples (VTXWriter in
    cpp/fenicsx-sf-naive/examples/linear_planewave2d_1/main.cpp:147-149): quadrilateral spaces are
    written as P^2 VTK_QUAD sub-cells per cell, here on that example's own mesh."""
    from fenicsx_fus_b200 import vtkio
    from fenicsx_fus_b200.unstructured2d import QuadFunctionSpace, QuadMesh
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_mesh_quad8400.npz"))
    m = QuadMesh(g["geometry"], g["topology_vtk"][:, (0, 1, 3, 2)], g["facet_lines"],
                 g["facet_values"], g["cell_values"])
    P = 3
    V = QuadFunctionSpace(m, P)
    X = V.tabulate_dof_coordinates()
    u = np.sin(50 * X[:, 0]) + X[:, 1]
    path = str(tmp_path / "u2d.vtu")
    npts, ncell = vtkio.write_vtu(path, V, {"u": u})
    assert (npts, ncell) == (V.ndofs, m.ncells * P * P)
    d = vtkio.read_vtu_arrays(path)
    assert np.array_equal(d["u"], u) and np.array_equal(d["Points"], X) and (d["types"] == 9).all()
    conn = d["connectivity"].reshape(-1, 4)
    assert np.array_equal(d["offsets"], 4 * np.arange(1, ncell + 1))
    Y = X[conn][:, :, :2]
    area = 0.5 * ((Y[:, :, 0] * np.roll(Y[:, :, 1], -1, 1)
                   - np.roll(Y[:, :, 0], -1, 1) * Y[:, :, 1]).sum(1))      # signed, counter-clockwise
    assert (area > 0).all() and abs(area.sum() - 0.12 * 0.07) < 1e-15
    # a structured RectMesh space goes through the same path
    Vr = fus.FunctionSpace(fus.RectMesh((3, 2)), 2)
    assert vtkio.write_vtu(str(tmp_path / "r.vtu"), Vr, {"one": np.ones(Vr.ndofs)}) == (Vr.ndofs, 24)
