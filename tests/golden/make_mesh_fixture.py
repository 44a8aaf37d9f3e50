"""Derive the unstructured-mesh fixture from the reference's own test mesh (run in the build
container, where /root/reference is mounted):

    python tests/golden/make_mesh_fixture.py

Source: cpp/fenicsx-sf/tests/test_operators3d/mesh.h5 (6 312 hexahedra, 7 939 vertices, unit cube,
facet tags; SURVEY.md section 4 "Fixtures"), parsed with fenicsx-fus_b200/hdf5min.py.  The fixture
keeps the raw datasets (VTK vertex order, as stored) plus reference outputs of the reference's own
acceptance set-up (tests/test_operators3d/main.cpp:59-79: P=4, u = sin(x) cos(pi y), c0 = 1.5e-3,
rho0 = 1e-3) computed by oracle/_ref, i.e. the cell loops of spectral_op.hpp on the reference's
contract<>/transpose<>: norms and 4096 sampled entries of the mass and stiffness vectors.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from fenicsx_fus_b200 import hdf5min  # noqa: E402
from fenicsx_fus_b200.unstructured import HexFunctionSpace, HexMesh  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402

SRC = "/root/reference/cpp/fenicsx-sf/tests/test_operators3d/mesh.h5"


def main():
    f = hdf5min.File(SRC)
    topo = f.read("/Mesh/hex/topology").astype(np.int32)
    geom = f.read("/Mesh/hex/geometry")
    fq = f.read("/MeshTags/hex_facets/topology").astype(np.int32)
    fv = f.read("/MeshTags/hex_facets/Values").astype(np.int32)
    mesh = HexMesh(geom, topo[:, (0, 1, 3, 2, 4, 5, 7, 6)], fq, fv)
    P = 4
    V = HexFunctionSpace(mesh, P)
    ref = Oracle(ref=True)
    ref.lib.fr_set_threads(1)
    G, dJ = ref.geometry(P, mesh.x, mesh.xdofmap)
    X = V.tabulate_dof_coordinates()
    u = np.sin(X[:, 0]) * np.cos(np.pi * X[:, 1])
    c0, rho0 = 1.5e-3, 1.0e-3
    m_coeff = np.full(mesh.ncells, 1.0 / rho0 / c0 / c0)      # main.cpp:86-88
    s_coeff = np.full(mesh.ncells, -1.0 / rho0)               # main.cpp:129-131
    ym = ref.mass_apply(P, V.dofmap, dJ, m_coeff, u, np.zeros(V.ndofs), use_ref_kernels=True)
    ys = ref.stiffness_apply(P, V.dofmap, G, ref.dphi(P), s_coeff, u, np.zeros(V.ndofs),
                             use_ref_kernels=True)
    rng = np.random.default_rng(4096)
    sample = np.sort(rng.choice(V.ndofs, 4096, replace=False))
    np.savez_compressed(
        os.path.join(HERE, "ref_mesh_hex6312.npz"), topology_vtk=topo, geometry=geom,
        facet_quads=fq, facet_values=fv, P=P, ndofs=V.ndofs,
        sample=sample, sample_xyz=X[sample], mass_sample=ym[sample], stiff_sample=ys[sample],
        mass_l2=np.linalg.norm(ym), stiff_l2=np.linalg.norm(ys), mass_sum=ym.sum(),
        stiff_sum=ys.sum(), volume=dJ.sum())
    print("ndofs", V.ndofs, "mass_l2", np.linalg.norm(ym), "stiff_l2", np.linalg.norm(ys),
          "bytes", os.path.getsize(os.path.join(HERE, "ref_mesh_hex6312.npz")))


SRC_2D = "/root/reference/cpp/fenicsx-sf-naive/examples/linear_planewave2d_1/mesh.h5"


def main_2d():
    """The mesh of the reference's 2-D example linear_planewave2d_1 (8 400 quadrilaterals on
    [0,0.12] x [-0.035,0.035], facet tags 1 = source edge, 2 = absorbing edge, 3 = walls) with
    outputs of the 2-D operators of cpp/fenicsx-sf-naive computed by oracle/_ref, i.e. the cell loop
    of spectral_op.hpp:275-318 on the reference's own 2-D contract<>/transpose<> (P = 4,
    u = sin(40 x) cos(30 pi y), the coefficients of the example: c0 = 1500, rho0 = 1000)."""
    from fenicsx_fus_b200.unstructured2d import QuadFunctionSpace, QuadMesh
    f = hdf5min.File(SRC_2D)
    name = "planewave_2d_1"
    topo = f.read(f"/Mesh/{name}/topology").astype(np.int32)
    geom = f.read(f"/Mesh/{name}/geometry")
    fl = f.read(f"/MeshTags/{name}_facets/topology").astype(np.int32)
    fv = f.read(f"/MeshTags/{name}_facets/Values").astype(np.int32)
    cv = f.read(f"/MeshTags/{name}_cells/Values").astype(np.int32)
    assert np.array_equal(f.read(f"/MeshTags/{name}_cells/topology"), topo)
    mesh = QuadMesh(geom, topo[:, (0, 1, 3, 2)], fl, fv, cv)
    P = 4
    V = QuadFunctionSpace(mesh, P)
    ref = Oracle(ref=True)
    G, dJ = ref.geometry_2d(P, mesh.x, mesh.xdofmap)
    X = V.tabulate_dof_coordinates()
    u = np.sin(40 * X[:, 0]) * np.cos(30 * np.pi * X[:, 1])
    c0, rho0 = 1500.0, 1000.0
    ym = ref.mass_apply_2d(P, V.dofmap, dJ, np.full(mesh.ncells, 1.0 / rho0 / c0 / c0), u,
                           np.zeros(V.ndofs))
    ys = ref.stiffness_apply_2d(P, V.dofmap, G, ref.dphi(P), np.full(mesh.ncells, -1.0 / rho0), u,
                                np.zeros(V.ndofs), use_ref_kernels=True)
    rng = np.random.default_rng(2048)
    sample = np.sort(rng.choice(V.ndofs, 2048, replace=False))
    out = os.path.join(HERE, "ref_mesh_quad8400.npz")
    np.savez_compressed(out, topology_vtk=topo, geometry=geom, facet_lines=fl, facet_values=fv,
                        cell_values=cv, P=P, ndofs=V.ndofs, sample=sample, sample_xy=X[sample, :2],
                        mass_sample=ym[sample], stiff_sample=ys[sample],
                        mass_l2=np.linalg.norm(ym), stiff_l2=np.linalg.norm(ys), area=dJ.sum())
    print("2-D: ndofs", V.ndofs, "mass_l2", np.linalg.norm(ym), "stiff_l2", np.linalg.norm(ys),
          "bytes", os.path.getsize(out))


if __name__ == "__main__":
    if "--2d-only" not in sys.argv:
        main()
    main_2d()
