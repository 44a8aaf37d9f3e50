"""CPU tests of the product's host side: set-up arithmetic vs the oracle, the C-ABI surface, and
loud failure without a GPU.  No device compute here."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, rel_l2, warp_vertices


def test_capi_exports_every_declared_symbol(fus):
    """include/fus_b200.h and the built library agree symbol for symbol."""
    hdr = open(os.path.join(ROOT, "include", "fus_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(fus_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 35
    from fenicsx_fus_b200 import capi
    lib = capi.load()
    for name in declared:
        assert hasattr(lib, name), f"library lacks {name}"
    assert declared == set(capi.SIGNATURES), declared ^ set(capi.SIGNATURES)
    assert lib.fus_version() >= 100


@pytest.mark.parametrize("P", range(1, 8))
def test_tables_vs_oracle(fus, orc, P):
    p, w = fus.gll(P)
    po, wo = orc.gll(P + 1)
    assert np.allclose(p, po, rtol=0, atol=2e-16) and np.allclose(w, wo, rtol=0, atol=2e-16)
    d, do = fus.tabulate_dphi(P), orc.dphi(P)
    assert np.allclose(d, do, rtol=0, atol=5e-15 * np.abs(do).max())


@pytest.mark.parametrize("n,P,mode", [((1, 1, 1), 2, 0), ((3, 2, 4), 3, 0), ((3, 2, 4), 3, 1),
                                      ((2, 2, 2), 7, 1), ((5, 1, 2), 1, 1)])
def test_box_mesh_and_dofmap_bit_exact(fus, orc, n, P, mode):
    m = fus.BoxMesh(n, (0.1, 0.0, -1.0), (1.0, 2.0, 3.0))
    xg, xd = orc.box_mesh(n, (0.1, 0.0, -1.0), (1.0, 2.0, 3.0))
    assert np.array_equal(m.x, xg) and np.array_equal(m.xdofmap, xd)
    assert np.array_equal(m.facets, orc.box_facets(n))
    V = fus.FunctionSpace(m, P, numbering=mode)
    dm = orc.box_dofmap(P, n, mode)
    assert np.array_equal(V.dofmap, dm)                     # gather/scatter indices: bit-exact
    assert V.ndofs == dm.max() + 1 == len(np.unique(dm))
    # collocation: node i of a cell sits at quadrature point i
    X = V.tabulate_dof_coordinates()
    pts, _ = orc.gll(P + 1)
    h = (np.array([1.0, 2.0, 3.0]) - np.array([0.1, 0.0, -1.0])) / np.array(n)
    c = 0
    i0, i1, i2 = 1, min(2, P), 0
    node = dm[c, (i0 * (P + 1) + i1) * (P + 1) + i2]
    assert np.allclose(X[node], np.array([0.1, 0.0, -1.0]) + h * np.array([pts[i0], pts[i1], pts[i2]]))


@pytest.mark.parametrize("kind", ["linear", "lossy", "westervelt"])
def test_boundary_vectors_vs_oracle_facet_assembly(fus, orc, kind):
    """The lumped vectors reproduce the oracle's facet-by-facet assembly of L and of the facet
    part of a (forms.py)."""
    from fenicsx_fus_b200 import capi
    P, n = 3, (3, 2, 2)
    m = fus.BoxMesh(n, (0, 0, 0), (0.3, 0.2, 0.2), warp=lambda x: warp_vertices(x, 0.07, 5))
    V = fus.FunctionSpace(m, P, numbering=1)
    nc, nd = m.ncells, V.ndofs
    rng = np.random.default_rng(3)
    c0, rho0 = rng.uniform(1400, 2500, nc), rng.uniform(900, 1900, nc)
    delta0, beta0 = rng.uniform(1e-3, 5e-3, nc), rng.uniform(3, 5, nc)
    src, dsrc, absb, bmass = (np.zeros(nd) for _ in range(4))
    lib = capi.load()
    rc = lib.fus_boundary_vectors(capi.KINDS[kind], P, nc, nd, m.x, m.xdofmap, V.dofmap,
                                  m.facets.shape[0], m.facets, c0, rho0, capi.optional(delta0),
                                  capi.optional(src), capi.optional(dsrc), capi.optional(absb),
                                  capi.optional(bmass))
    assert rc == 0
    G, dJ = orc.geometry(P, m.x, m.xdofmap)
    fn, fs = orc.facet_data(P, m.x, m.xdofmap, m.facets)
    mdl = orc.model(kind, P, nd, V.dofmap, G, dJ, orc.dphi(P), c0, rho0, delta0, beta0, m.facets,
                    fn, fs, 0.5e6, 1e5, 1500.0)
    # mass: oracle's assembled m minus the volume part = facet mass
    vol = orc.mass_apply(P, V.dofmap, dJ, 1 / rho0 / c0 ** 2, np.ones(nd), np.zeros(nd))
    assert rel_l2(vol + bmass, mdl.mass()) < 1e-14
    if kind == "linear":
        assert not bmass.any() and not dsrc.any()
    # L: with K u removed (u = 0) f1*m = g*src + dg*dsrc - absb*v
    t = 0.7e-6
    v = rng.uniform(-1, 1, nd)
    b = mdl.f1(t, np.zeros(nd), v) * mdl.mass()
    if kind == "westervelt":
        pytest.skip("westervelt f1 has the solution-dependent mass; covered by linear/lossy here")
    f, p0, s0, w0 = 0.5e6, 1e5, 1500.0, 2 * np.pi * 0.5e6
    win = 0.5 * (1 - np.cos(f * np.pi * t / 4))
    dwin = 0.5 * np.pi * f / 4 * np.sin(f * np.pi * t / 4)
    if kind == "linear":
        g, dg = win * p0 * w0 / s0 * np.cos(w0 * t), 0.0
        kv = np.zeros(nd)
    else:
        g = win * 2 * p0 * w0 / s0 * np.cos(w0 * t)
        dg = dwin * 2 * p0 * w0 / s0 * np.cos(w0 * t) - win * 2 * p0 * w0 * w0 / s0 * np.sin(w0 * t)
        kv = orc.stiffness_apply(P, V.dofmap, G, orc.dphi(P), -delta0 / rho0 / c0 ** 2, v, np.zeros(nd))
    assert rel_l2(g * src + dg * dsrc - absb * v + kv, b) < 1e-13


def test_compute_paths_fail_loudly_without_gpu(fus):
    """No CPU fallback: creating a context without a usable device is an error."""
    if fus.device_count() > 0:
        pytest.skip("a GPU is present")
    m = fus.BoxMesh((1, 1, 1))
    V = fus.FunctionSpace(m, 2)
    with pytest.raises(fus.FusError, match="no usable CUDA device|CUDA"):
        V.context()


def test_bad_arguments_are_errors_not_crashes(fus):
    from fenicsx_fus_b200 import capi
    lib = capi.load()
    assert lib.fus_gll(0, np.zeros(4), np.zeros(4)) < 0
    assert lib.fus_box_dofmap(2, np.array([1, 0, 1], dtype=np.int32), 0, np.zeros(27, dtype=np.int32)) < 0
    h = C.c_void_p()
    # degree outside 1..7 -> FUS_ERR_UNSUPPORTED (the reference silently maps unknown P to Qdegree 0,
    # spectral_op.hpp:59)
    rc = lib.fus_ctx_create(9, 1, 1000, 1000, np.zeros(1000, dtype=np.int32), None, None,
                            np.zeros(100), 0, C.byref(h))
    assert rc < 0 and lib.fus_last_error()


def test_source_disc_tags(fus, orc):
    """Config 4 stand-in for the bowl transducer: a disc of source facets on x = 0; the lumped
    source vector then integrates to (1/rho) * tagged area and matches the oracle's facet assembly."""
    from fenicsx_fus_b200 import capi
    P, n, L = 2, (4, 8, 8), 0.08
    m = fus.BoxMesh(n, (0, 0, 0), (0.04, L, L))
    ntag = m.tag_source_disc((L / 2, L / 2), 0.025)
    h = L / 8
    cen = (np.arange(8) + 0.5) * h
    expect = sum(np.hypot(y - L / 2, z - L / 2) <= 0.025 for y in cen for z in cen)
    assert ntag == expect and 0 < ntag < 64
    assert (m.facets[:, 2] == 2).sum() == 64              # the absorbing face is untouched
    V = fus.FunctionSpace(m, P)
    nc, nd = m.ncells, V.ndofs
    c0, rho0 = np.full(nc, 1480.0), np.full(nc, 1000.0)
    src, dsrc, absb, bmass = (np.zeros(nd) for _ in range(4))
    rc = capi.load().fus_boundary_vectors(capi.KINDS["linear"], P, nc, nd, m.x, m.xdofmap, V.dofmap,
                                          m.facets.shape[0], m.facets, c0, rho0, None,
                                          capi.optional(src), capi.optional(dsrc),
                                          capi.optional(absb), capi.optional(bmass))
    assert rc == 0
    assert abs(src.sum() - ntag * h * h / 1000.0) < 1e-15
    fn, fs = orc.facet_data(P, m.x, m.xdofmap, m.facets)
    ref = np.zeros(nd)
    for k in np.flatnonzero(m.facets[:, 2] == 1):
        np.add.at(ref, V.dofmap[m.facets[k, 0], fn[k]], fs[k] / 1000.0)
    assert np.allclose(src, ref, rtol=1e-13, atol=1e-20)


@pytest.mark.parametrize("P", [1, 2, 3, 4, 5, 6, 7])
def test_trilinear_map_rebuilds_reference_geometry(fus, orc, P):
    """The arithmetic of the geometry_mode=2 stiffness kernel (fus_trilinear.hpp, compiled for the
    host behind fus_trilinear_geometry): |det J| w K K^T and |det J| w rebuilt from the 21 monomial
    coefficients of the cell map equal compute_scaled_geometrical_factor /
    compute_scaled_jacobian_determinant (precompute.hpp:33-213, the oracle) on warped cells placed
    away from the origin.  Both sides lose eps*|x|/h to cancellation, hence 1e-12."""
    from fenicsx_fus_b200 import capi
    lib = capi.load()
    m = fus.BoxMesh((3, 2, 2), (0.3, -0.2, 1.0), (0.33, -0.18, 1.02),
                    warp=lambda x: warp_vertices(x, 0.12, 11))
    G, dJ = orc.geometry(P, m.x, m.xdofmap)
    assert np.abs(G[:, :, [1, 2, 4]]).max() > 1e-2 * np.abs(G).max()      # all six entries matter
    co = np.zeros((m.ncells, 24))
    assert lib.fus_trilinear_coeffs(m.ncells, m.x, m.xdofmap, co) == 0
    assert (co[:, 21:] == 0).all()
    G2, dJ2 = np.zeros_like(G), np.zeros_like(dJ)
    assert lib.fus_trilinear_geometry(P, m.ncells, co, capi.optional(G2), capi.optional(dJ2)) == 0
    assert np.abs(G2 - G).max() <= 1e-12 * np.abs(G).max()
    assert np.abs(dJ2 - dJ).max() <= 1e-12 * np.abs(dJ).max()
    # an affine cell has no mixed terms; G and detJ alone can be requested
    mb = fus.BoxMesh((2, 1, 1), (0, 0, 0), (2.0, 0.5, 0.25))
    cb = np.zeros((mb.ncells, 24))
    assert lib.fus_trilinear_coeffs(mb.ncells, mb.x, mb.xdofmap, cb) == 0
    assert np.allclose(cb[:, :9].reshape(-1, 3, 3), np.diag([1.0, 0.5, 0.25]), atol=1e-16)
    assert (cb[:, 9:] == 0).all()
    dJb = np.zeros((mb.ncells, (P + 1) ** 3))
    assert lib.fus_trilinear_geometry(P, mb.ncells, cb, None, capi.optional(dJb)) == 0
    assert abs(dJb.sum() - 2.0 * 0.5 * 0.25) < 1e-14
    assert lib.fus_trilinear_geometry(0, 1, cb, None, None) < 0


def test_trilinear_map_on_the_reference_mesh(fus, orc):
    """Same check on the reference's own unstructured test mesh (6 312 general hexahedra,
    cpp/fenicsx-sf/tests/test_operators3d/mesh.h5, committed fixture)."""
    import os
    from fenicsx_fus_b200 import capi
    d = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_mesh_hex6312.npz"))
    x = np.ascontiguousarray(d["geometry"], dtype=np.float64)
    # VTK vertex order -> DOLFINx tensor order
    cells = np.ascontiguousarray(d["topology_vtk"][::13, (0, 1, 3, 2, 4, 5, 7, 6)], dtype=np.int32)
    P = 3
    G, dJ = orc.geometry(P, x, cells)
    co = np.zeros((cells.shape[0], 24))
    lib = capi.load()
    assert lib.fus_trilinear_coeffs(cells.shape[0], x, cells, co) == 0
    G2, dJ2 = np.zeros_like(G), np.zeros_like(dJ)
    assert lib.fus_trilinear_geometry(P, cells.shape[0], co, capi.optional(G2), capi.optional(dJ2)) == 0
    assert np.abs(G2 - G).max() <= 1e-12 * np.abs(G).max()
    assert np.abs(dJ2 - dJ).max() <= 1e-12 * np.abs(dJ).max()


def test_header_is_valid_c_and_c_example_links(fus, tmp_path):
    """include/fus_b200.h is a C header (no C++ in the signatures): gcc -std=c11 -pedantic accepts
    it and examples/c_abi_minimal.c links against the library."""
    import subprocess
    hdr = os.path.join(ROOT, "include", "fus_b200.h")
    res = subprocess.run(["/usr/bin/gcc", "-std=c11", "-Wall", "-Wextra", "-pedantic", "-Werror",
                          "-fsyntax-only", "-x", "c", hdr], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    exe = str(tmp_path / "c_abi_minimal")
    res = subprocess.run(["/usr/bin/gcc", "-std=c11", "-O1", "-Wall", "-Werror",
                          "-I" + os.path.join(ROOT, "include"),
                          os.path.join(ROOT, "examples", "c_abi_minimal.c"),
                          "-L" + os.path.join(ROOT, "fenicsx-fus_b200", "lib"), "-lfus_b200", "-lm",
                          "-Wl,-rpath," + os.path.join(ROOT, "fenicsx-fus_b200", "lib"), "-o", exe],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    if fus.device_count() == 0:                  # no GPU here: it must fail loudly, not fall back
        run = subprocess.run([exe], capture_output=True, text=True)
        assert run.returncode != 0 and "no CPU fallback" in run.stderr


def test_bad_arguments_of_the_newer_entry_points(fus):
    """Host-side entry points added for the trilinear, 2-D and partition paths return error codes
    (never crash) on malformed input; device entry points still fail loudly without a GPU."""
    from fenicsx_fus_b200 import capi
    lib = capi.load()
    i32, f64 = np.int32, np.float64
    assert lib.fus_trilinear_coeffs(-1, np.zeros(3), np.zeros(8, dtype=i32), np.zeros(24)) < 0
    assert lib.fus_trilinear_geometry(16, 1, np.zeros(24), None, None) < 0
    assert lib.fus_rect_mesh(np.array([0, 2], dtype=i32), np.zeros(2), np.ones(2), np.zeros(9),
                             np.zeros(4, dtype=i32)) < 0
    assert lib.fus_rect_dofmap(0, np.array([1, 1], dtype=i32), np.zeros(4, dtype=i32)) < 0
    assert lib.fus_rect_num_dofs(2, np.array([3, 2], dtype=i32)) == 7 * 5
    assert lib.fus_rect_facets(np.array([3, 2], dtype=i32), None) == 10
    h = C.c_void_p()
    ng, pg = np.array([4, 4, 4], dtype=i32), np.array([2, 2, 1], dtype=i32)
    assert lib.fus_box_partition_create(2, ng, pg, 4, 1, C.byref(h)) < 0        # rank outside the grid
    assert not h.value and b"rank" in lib.fus_last_error()
    assert lib.fus_box_partition_create(0, ng, pg, 0, 1, C.byref(h)) < 0        # degree
    assert lib.fus_box_partition_create(2, ng, pg, 0, 7, C.byref(h)) < 0        # numbering
    assert lib.fus_box_partition_create(2, ng, pg, 3, 1, C.byref(h)) == 0 and h.value
    sizes = np.zeros(9, dtype=np.int64)
    nl, lo = np.zeros(3, dtype=i32), np.zeros(3, dtype=i32)
    assert lib.fus_box_partition_info(h, sizes, nl, lo) == 0
    assert nl.tolist() == [2, 2, 4] and lo.tolist() == [2, 2, 0] and sizes[0] == 16
    assert sizes[8] == 9 ** 3 and sizes[2] < sizes[1]                           # it has ghosts
    assert lib.fus_box_partition_arrays(h, *([None] * 10)) == 0                 # all outputs optional
    assert lib.fus_box_partition_destroy(h) == 0
    if fus.device_count() == 0:
        m = fus.RectMesh((2, 2))
        V = fus.FunctionSpace(m, 2)
        rc = lib.fus_ctx_create_from_mesh_2d(2, m.ncells, V.ndofs, V.ndofs, V.dofmap, m.x.shape[0],
                                             m.x, m.xdofmap, 0, C.byref(h))
        assert rc < 0 and not h.value and b"no CPU fallback" in lib.fus_last_error()
        with pytest.raises(fus.FusError):
            fus.FunctionSpace(fus.BoxMesh((1, 1, 1)), 2).context(lean=True)
