"""Unstructured hexahedral meshes: the reference's own test mesh
(cpp/fenicsx-sf/tests/test_operators3d/mesh.h5, committed as tests/golden/ref_mesh_hex6312.npz by
tests/golden/make_mesh_fixture.py), conforming GLL numbering, oracle identities, and the CUDA
operators / models on it."""
import os

import numpy as np
import pytest

from conftest import rel_l2

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ref_mesh_hex6312.npz")
REF_H5 = "/root/reference/cpp/fenicsx-sf/tests/test_operators3d/mesh.h5"


@pytest.fixture(scope="module")
def ref_mesh(fus):
    from fenicsx_fus_b200.unstructured import HexMesh
    g = np.load(GOLD)
    return HexMesh(g["geometry"], g["topology_vtk"][:, (0, 1, 3, 2, 4, 5, 7, 6)],
                   g["facet_quads"], g["facet_values"]), g


def _match(X, xyz):
    """indices of the dofs located at the given coordinates"""
    key = {tuple(r): i for i, r in enumerate(np.round(X, 9))}
    return np.array([key[tuple(r)] for r in np.round(xyz, 9)])


def test_hdf5_reader_matches_fixture(fus):
    if not os.path.exists(REF_H5):
        pytest.skip("/root/reference not mounted")
    from fenicsx_fus_b200 import hdf5min
    f = hdf5min.File(REF_H5)
    g = np.load(GOLD)
    assert f.listdir("/") == ["Mesh", "MeshTags"]
    assert np.array_equal(f.read("/Mesh/hex/topology"), g["topology_vtk"])
    assert np.array_equal(f.read("/Mesh/hex/geometry"), g["geometry"])
    assert np.array_equal(f.read("/MeshTags/hex_facets/Values"), g["facet_values"])
    with pytest.raises(KeyError):
        f.read("/Mesh/hex/nothing")


def test_mesh_topology(ref_mesh):
    m, g = ref_mesh
    assert m.ncells == 6312 and m.x.shape == (7939, 3)
    assert m.facets.shape[0] == 2124 and np.all(m.facets[:, 2] == 1)     # every exterior facet is tagged
    # exterior facets lie on the surface of the unit cube
    lf_corners = {0: (0, 1, 2, 3), 1: (0, 1, 4, 5), 2: (0, 2, 4, 6), 3: (1, 3, 5, 7),
                  4: (2, 3, 6, 7), 5: (4, 5, 6, 7)}
    for c, lf, _ in m.facets[::17]:
        X = m.x[m.xdofmap[c, list(lf_corners[lf])]]
        on = [np.allclose(X[:, d], 0) or np.allclose(X[:, d], 1) for d in range(3)]
        assert any(on)


@pytest.mark.parametrize("P", [1, 2, 3, 5])
def test_conforming_numbering(fus, orc, ref_mesh, P):
    from fenicsx_fus_b200.unstructured import HexFunctionSpace
    m, _ = ref_mesh
    V = HexFunctionSpace(m, P)
    k = V.counts
    assert k["vertices"] - k["edges"] + k["faces"] - k["cells"] == 1            # Euler, a ball
    assert V.ndofs == k["vertices"] + k["edges"] * (P - 1) + k["faces"] * (P - 1) ** 2 \
        + k["cells"] * (P - 1) ** 3
    assert np.array_equal(np.unique(V.dofmap), np.arange(V.ndofs))
    X, spread = V.tabulate_dof_coordinates(return_spread=True)
    assert spread < 1e-14                       # all cells sharing a dof put it at the same point
    assert len(np.unique(np.round(X, 9), axis=0)) == V.ndofs
    if P <= 3:
        G, dJ = orc.geometry(P, m.x, m.xdofmap)
        assert (dJ > 0).all()
        if P >= 2:                  # the 2-point rule does not integrate a trilinear det J exactly
            assert abs(dJ.sum() - 1.0) < 1e-13
        nd = V.ndofs
        co = np.ones(m.ncells)
        assert np.abs(orc.stiffness_apply(P, V.dofmap, G, orc.dphi(P), co, np.ones(nd),
                                          np.zeros(nd))).max() < 1e-14
        lin = 2 * X[:, 0] - 3 * X[:, 1] + 0.5 * X[:, 2]
        kl = orc.stiffness_apply(P, V.dofmap, G, orc.dphi(P), co, lin, np.zeros(nd))
        fn, fs = orc.facet_data(P, m.x, m.xdofmap, m.facets)
        if P >= 2:
            assert abs(fs.sum() - 6.0) < 1e-12
        bnd = np.zeros(nd, bool)
        for f in range(m.facets.shape[0]):
            bnd[V.dofmap[m.facets[f, 0], fn[f]]] = True
        assert np.abs(kl[~bnd]).max() < 1e-13 and np.abs(kl[bnd]).max() > 1e-4


def test_box_as_unstructured_equals_box_path(fus, orc):
    """A structured box fed through the unstructured numbering gives the same operator (compared
    through dof coordinates, the numberings differ)."""
    from fenicsx_fus_b200.unstructured import HexFunctionSpace, HexMesh
    P, n = 3, (3, 2, 2)
    xg, xd = orc.box_mesh(n, (0, 0, 0), (1.0, 0.7, 0.9))
    rng = np.random.default_rng(1)
    perm = rng.permutation(xd.shape[0])                      # scramble the cell order too
    mesh = HexMesh(xg, xd[perm])
    V = HexFunctionSpace(mesh, P)
    dm = orc.box_dofmap(P, n, 0)
    assert V.ndofs == dm.max() + 1
    G, _ = orc.geometry(P, xg, xd)
    Gu, _ = orc.geometry(P, mesh.x, mesh.xdofmap)
    # coordinates of the lexicographic box dofs
    pts, _ = orc.gll(P + 1)
    Xu = V.tabulate_dof_coordinates()
    field = lambda X: np.sin(3 * X[:, 0]) * np.cos(2 * X[:, 1]) + X[:, 2] ** 2   # noqa: E731
    class _B:                                                 # box coordinates via the same map
        pass
    from fenicsx_fus_b200.unstructured import HexMesh as HM
    Vb = HexFunctionSpace(HM(xg, xd), P, renumber=False)
    Xb_all = Vb.tabulate_dof_coordinates()
    # evaluate K f on both numberings and compare at matching coordinates
    co_box = 1.0 + 0.1 * np.arange(xd.shape[0])
    yu = orc.stiffness_apply(P, V.dofmap, Gu, orc.dphi(P), co_box[perm], field(Xu), np.zeros(V.ndofs))
    yb = orc.stiffness_apply(P, Vb.dofmap, G, orc.dphi(P), co_box, field(Xb_all), np.zeros(V.ndofs))
    idx = _match(Xu, Xb_all)
    assert rel_l2(yu[idx], yb) < 1e-13
    # and the box generator's own dofmap agrees with it as well
    Xlex = np.zeros((V.ndofs, 3))
    N = P + 1
    X8 = xg[xd]
    xi = np.stack(np.meshgrid(pts, pts, pts, indexing="ij"), -1).reshape(-1, 3)
    acc = 0.0
    for v in range(8):
        a, b, c = v & 1, (v >> 1) & 1, (v >> 2) & 1
        w = ((xi[:, 0] if a else 1 - xi[:, 0]) * (xi[:, 1] if b else 1 - xi[:, 1])
             * (xi[:, 2] if c else 1 - xi[:, 2]))
        acc = acc + w[None, :, None] * X8[:, v, None, :]
    Xlex[dm.reshape(-1)] = acc.reshape(-1, 3)
    ylex = orc.stiffness_apply(P, dm, G, orc.dphi(P), co_box, field(Xlex), np.zeros(V.ndofs))
    assert rel_l2(yu[_match(Xu, Xlex)], ylex) < 1e-13


def test_reference_acceptance_setup_oracle(fus, orc, ref_mesh):
    """The reference's own operator test (tests/test_operators3d/main.cpp:59-131: P=4,
    u = sin(x) cos(pi y), c0 = 1.5e-3, rho0 = 1e-3) -- oracle vs the fixture values produced by
    the reference's kernels (oracle/_ref)."""
    from fenicsx_fus_b200.unstructured import HexFunctionSpace
    m, g = ref_mesh
    P = int(g["P"])
    V = HexFunctionSpace(m, P)
    assert V.ndofs == int(g["ndofs"])
    X = V.tabulate_dof_coordinates()
    u = np.sin(X[:, 0]) * np.cos(np.pi * X[:, 1])
    G, dJ = orc.geometry(P, m.x, m.xdofmap)
    ym = orc.mass_apply(P, V.dofmap, dJ, np.full(m.ncells, 1 / 1e-3 / 1.5e-3 ** 2), u, np.zeros(V.ndofs))
    ys = orc.stiffness_apply(P, V.dofmap, G, orc.dphi(P), np.full(m.ncells, -1 / 1e-3), u,
                             np.zeros(V.ndofs))
    idx = _match(X, g["sample_xyz"])
    assert rel_l2(ym[idx], g["mass_sample"]) < 1e-13 and rel_l2(ys[idx], g["stiff_sample"]) < 1e-12
    assert abs(np.linalg.norm(ym) - float(g["mass_l2"])) < 1e-12 * float(g["mass_l2"])
    assert abs(np.linalg.norm(ys) - float(g["stiff_l2"])) < 1e-12 * float(g["stiff_l2"])


@pytest.mark.gpu
@pytest.mark.parametrize("P", [2, 4, 5])
def test_gpu_operators_on_reference_mesh(fus, orc, gpu, ref_mesh, P):
    from fenicsx_fus_b200.unstructured import HexFunctionSpace
    m, g = ref_mesh
    V = HexFunctionSpace(m, P)
    X = V.tabulate_dof_coordinates()
    u = np.sin(X[:, 0]) * np.cos(np.pi * X[:, 1])
    G, dJ = orc.geometry(P, m.x, m.xdofmap)
    Gd, dJd = V.context().geometry()
    assert rel_l2(Gd, G) < 1e-13 and rel_l2(dJd, dJ) < 1e-13
    s_coeff = np.full(m.ncells, -1 / 1e-3)
    m_coeff = np.full(m.ncells, 1 / 1e-3 / 1.5e-3 ** 2)
    ys = fus.StiffnessSpectral3D(V)(u, s_coeff, np.zeros(V.ndofs))
    ym = fus.MassSpectral3D(V)(u, m_coeff, np.zeros(V.ndofs))
    yo = orc.stiffness_apply(P, V.dofmap, G, orc.dphi(P), s_coeff, u, np.zeros(V.ndofs))
    mo = orc.mass_apply(P, V.dofmap, dJ, m_coeff, u, np.zeros(V.ndofs))
    assert rel_l2(ys, yo) < 1e-12 and rel_l2(ym, mo) < 1e-12
    if P == int(g["P"]):
        idx = _match(X, g["sample_xyz"])
        assert rel_l2(ys[idx], g["stiff_sample"]) < 1e-12
        assert rel_l2(ym[idx], g["mass_sample"]) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["linear", "lossy", "westervelt"])
def test_gpu_models_on_reference_mesh(fus, orc, gpu, ref_mesh, kind):
    """RK4 on the unstructured mesh (all 2 124 exterior facets carry tag 1, i.e. source everywhere),
    15 steps from rest, against the oracle's literal loop."""
    from fenicsx_fus_b200.unstructured import HexFunctionSpace
    m, _ = ref_mesh
    P = 3
    V = HexFunctionSpace(m, P)
    nc, nd = m.ncells, V.ndofs
    cx = m.x[m.xdofmap].mean(axis=1)[:, 0]
    c0 = np.where(cx < 0.5, 1500.0, 2300.0)
    rho0 = np.where(cx < 0.5, 1000.0, 1700.0)
    f, p0, s0 = 2.0e3, 1.0e5, 1500.0                  # unit cube: wavelength 0.75 m
    # small attenuation: delta k_max^2 dt must stay inside RK4's stability interval
    delta0 = np.full(nc, fus.compute_diffusivity_of_sound(2 * np.pi * f, 1500.0, 0.01))
    beta0 = np.full(nc, 3.5)
    dt = 0.15 * m.h_min() / (2300.0 * P * P)
    G, dJ = orc.geometry(P, m.x, m.xdofmap)
    fn, fs = orc.facet_data(P, m.x, m.xdofmap, m.facets)
    om = orc.model(kind, P, nd, V.dofmap, G, dJ, orc.dphi(P), c0, rho0, delta0, beta0, m.facets,
                   fn, fs, f, p0, s0)
    if kind == "linear":
        mdl = fus.LinearSpectral3D(V, c0, rho0, f, p0, s0)
    elif kind == "lossy":
        mdl = fus.LossySpectral3D(V, c0, rho0, delta0, f, p0, s0)
    else:
        mdl = fus.WesterveltSpectral3D(V, c0, rho0, delta0, beta0, f, p0, s0)
    assert rel_l2(mdl.mass(), om.mass()) < 1e-12
    mdl.init()
    t0, tf = 1.0e-4, 1.0e-4 + 15 * dt
    steps = mdl.rk4(t0, tf, dt)
    u, v = np.zeros(nd), np.zeros(nd)
    assert om.rk4(t0, tf, dt, u, v) == steps
    assert np.isfinite(u).all() and 0 < np.abs(u).max() < 1e3 * p0     # a stable run
    assert rel_l2(mdl.u_sol(), u) < 1e-10 and rel_l2(mdl.v_sol(), v) < 1e-10


def test_reference_five_operator_checks_vs_dense(fus, orc, ref_mesh):
    """The five checks of cpp/fenicsx-sf-naive/tests/test_operators3d/main.cpp:104-339 on the
    reference's unstructured mesh, with its parameters (c0 = 1.5, rho0 = 1, delta0 = 10, beta0 = 10;
    u = 1, u_n = 2, w_n = u_n^2, v_n = cos x sin(pi y) cos(2 pi z)): the reference compares each
    spectral operator with an independently assembled FFCx form; here the independent side is the
    dense evaluation of the same GLL-quadrature forms (tests/dense_ref.py) on a subset of cells."""
    from dense_ref import element_matrices
    from fenicsx_fus_b200.unstructured import HexFunctionSpace
    m, _ = ref_mesh
    P = 2
    V = HexFunctionSpace(m, P)
    cells = np.arange(0, m.ncells, 97)                       # 66 general hexahedra
    dm = np.ascontiguousarray(V.dofmap[cells])
    xd = np.ascontiguousarray(m.xdofmap[cells])
    G, dJ = orc.geometry(P, m.x, xd)
    X = V.tabulate_dof_coordinates()
    nd, nc = V.ndofs, cells.size
    c0, rho0, delta0, beta0 = 1.5, 1.0, 10.0, 10.0
    u, un = np.ones(nd), np.full(nd, 2.0)
    wn = un * un
    vn = np.cos(X[:, 0]) * np.sin(np.pi * X[:, 1]) * np.cos(2 * np.pi * X[:, 2])
    checks = [("m1", "mass", 1.0 / rho0 / c0 ** 2, u),                          # main.cpp:104-160
              ("m2", "mass", -2.0 * beta0 / rho0 ** 2 / c0 ** 4, un),           # :162-207
              ("m3", "mass", 2.0 * beta0 / rho0 ** 2 / c0 ** 4, wn),            # :209-254
              ("b1", "stiff", -1.0 / rho0, vn),                                 # :256-297
              ("b2", "stiff", -delta0 / rho0 / c0 ** 2, vn)]                    # :299-339
    for name, op, coef, x in checks:
        cf = np.full(nc, coef)
        if op == "mass":
            y = orc.mass_apply(P, dm, dJ, cf, x, np.zeros(nd))
        else:
            y = orc.stiffness_apply(P, dm, G, orc.dphi(P), cf, x, np.zeros(nd))
        yd = np.zeros(nd)
        for k in range(nc):
            K, mdiag = element_matrices(P, m.x[xd[k]], coef)
            yd[dm[k]] += (coef * mdiag * x[dm[k]]) if op == "mass" else K @ x[dm[k]]
        assert rel_l2(y, yd) < 1e-12, name
