"""Point evaluation after the time loop (the reference's
cpp/mwe/parallel_eval_line and python utils.compute_eval_params)."""
import os

import numpy as np
import pytest

from conftest import ROOT, warp_vertices


def test_eval_polynomial_on_box_is_exact(fus):
    from fenicsx_fus_b200 import sampling
    P = 4
    m = fus.BoxMesh((3, 2, 2), (0, 0, 0), (1.5, 1.0, 0.8))
    V = fus.FunctionSpace(m, P)
    X = V.tabulate_dof_coordinates()
    f = lambda x: 1 + x[:, 0] ** 4 - 2 * x[:, 1] ** 3 * x[:, 2] + x[:, 0] * x[:, 1] * x[:, 2] ** 2  # noqa: E731
    u = f(X)
    rng = np.random.default_rng(0)
    pts = rng.uniform([0, 0, 0], [1.5, 1.0, 0.8], (200, 3))
    pk, cells, xi, keep = sampling.compute_eval_params(m, pts.T)          # (3, n) like the reference
    assert len(keep) == 200 and np.array_equal(pk, pts)
    vals = sampling.eval_function(V, u, cells, xi)
    assert np.abs(vals - f(pts)).max() < 1e-12
    # points outside the mesh are dropped, like points not on this process in the reference
    out = np.array([[2.0, 0.5, 0.5], [0.5, 0.5, 0.5], [-0.1, 0.2, 0.2]])
    pk, cells, xi, keep = sampling.compute_eval_params(m, out)
    assert keep.tolist() == [1]
    # the line sampler of the reference example
    pl, vl = sampling.eval_line(V, u, (0.0, 0.5, 0.4), (1.5, 0.5, 0.4), 50)
    assert len(vl) == 50 and np.abs(vl - f(pl)).max() < 1e-12


def test_eval_linear_field_on_unstructured_mesh(fus):
    """A field linear in x lies in the isoparametric space, also on non-affine cells."""
    from fenicsx_fus_b200 import sampling
    from fenicsx_fus_b200.unstructured import HexFunctionSpace, HexMesh
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_mesh_hex6312.npz"))
    m = HexMesh(g["geometry"], g["topology_vtk"][:, (0, 1, 3, 2, 4, 5, 7, 6)], reorder="morton")
    V = HexFunctionSpace(m, 3)
    X = V.tabulate_dof_coordinates()
    f = lambda x: 0.3 + 2 * x[:, 0] - x[:, 1] + 0.5 * x[:, 2]   # noqa: E731
    rng = np.random.default_rng(1)
    pts = rng.uniform(0.01, 0.99, (100, 3))
    pk, cells, xi, keep = sampling.compute_eval_params(m, pts)
    assert len(keep) == 100
    # the located cell really contains the point
    xc, _ = sampling._trilinear(m.x[m.xdofmap[cells]], xi)
    assert np.abs(xc - pk).max() < 1e-12
    assert np.abs(sampling.eval_function(V, f(X), cells, xi) - f(pk)).max() < 1e-12


def test_eval_on_warped_box_reproduces_nodal_values(fus):
    from fenicsx_fus_b200 import sampling
    P = 3
    m = fus.BoxMesh((3, 3, 2), warp=lambda x: warp_vertices(x, 0.06, 2))
    V = fus.FunctionSpace(m, P, numbering=0)
    X = V.tabulate_dof_coordinates()
    rng = np.random.default_rng(3)
    u = rng.uniform(-1, 1, V.ndofs)
    sel = rng.choice(V.ndofs, 60, replace=False)
    pk, cells, xi, keep = sampling.compute_eval_params(m, X[sel])
    assert len(keep) == 60                                   # nodes (also boundary ones) are found
    assert np.abs(sampling.eval_function(V, u, cells, xi) - u[sel]).max() < 1e-10


def test_eval_on_the_reference_2d_example_mesh(fus):
    """Line sampling as in the reference's 2-D examples (python/src/fenicsxfus/utils.py:10-47 on the
    axis of the planar-wave runs): a degree-P polynomial is reproduced exactly on the example's mesh."""
    from fenicsx_fus_b200 import sampling
    from fenicsx_fus_b200.unstructured2d import QuadFunctionSpace, QuadMesh
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_mesh_quad8400.npz"))
    m = QuadMesh(g["geometry"], g["topology_vtk"][:, (0, 1, 3, 2)])
    P = 3
    V = QuadFunctionSpace(m, P)
    X = V.tabulate_dof_coordinates()
    f = lambda x: 1 + 50 * x[:, 0] ** 3 - 200 * x[:, 0] * x[:, 1] ** 2 + x[:, 1]  # noqa: E731
    u = f(X)
    pl, vl = sampling.eval_line(V, u, (0.0, 0.001), (0.12, 0.001), 60)
    assert pl.shape == (60, 2) and np.abs(vl - f(pl)).max() < 1e-12
    rng = np.random.default_rng(1)
    pts = rng.uniform([0, -0.035, 0], [0.12, 0.035, 0], (100, 3))         # (n,3) with z = 0 accepted
    pk, cells, xi, keep = sampling.compute_eval_params(m, pts)
    assert len(keep) == 100 and xi.shape == (100, 2)
    assert np.abs(sampling.eval_function(V, u, cells, xi) - f(pts)).max() < 1e-12
    pk, cells, xi, keep = sampling.compute_eval_params(m, np.array([[0.2, 0.0], [0.06, 0.0]]))
    assert keep.tolist() == [1]


def test_eval_surface_like_the_reference_example(fus, tmp_path):
    """cpp/mwe/parallel_eval_surface: u = sin(2 pi x) cos(2 pi y) on [-1,1]^2 (32 x 32 cells), sampled
    on a 100 x 100 grid and appended to a text file; and a tilted plane through a 3-D box."""
    from fenicsx_fus_b200 import sampling
    m = fus.RectMesh((32, 32), (-1.0, -1.0), (1.0, 1.0))
    V = fus.FunctionSpace(m, 4)
    X = V.tabulate_dof_coordinates()
    f2 = lambda x: np.sin(2 * np.pi * x[:, 0]) * np.cos(2 * np.pi * x[:, 1])  # noqa: E731
    out = tmp_path / "surface_data.txt"
    pk, vals = sampling.eval_surface(V, f2(X), (-1.0, -1.0), (2.0, 0.0), (0.0, 2.0), 100, path=str(out))
    assert len(vals) == 100 * 100 and np.abs(vals - f2(pk)).max() < 2e-5       # interpolation error
    rows = np.loadtxt(out, delimiter=",")
    assert rows.shape == (10000, 3) and np.array_equal(rows[:, 2], vals)
    assert np.allclose(rows[:100, 1], -1.0) and np.allclose(rows[:100, 0], np.linspace(-1, 1, 100))
    # 3-D: a polynomial of the space is reproduced exactly on a tilted plane; corners outside the box drop
    m3 = fus.BoxMesh((3, 2, 2), (0, 0, 0), (1.5, 1.0, 0.8))
    V3 = fus.FunctionSpace(m3, 3)
    f3 = lambda x: 1 + x[:, 0] ** 3 - x[:, 1] * x[:, 2] ** 2                   # noqa: E731
    u3 = f3(V3.tabulate_dof_coordinates())
    pk, vals = sampling.eval_surface(V3, u3, (0.0, 0.0, 0.1), (1.5, 0.0, 0.6), (0.0, 1.0, 0.0), (20, 10))
    assert len(vals) == 200 and np.abs(vals - f3(pk)).max() < 1e-12
    pk, vals = sampling.eval_surface(V3, u3, (0.0, 0.0, 0.5), (3.0, 0.0, 0.0), (0.0, 1.0, 0.0), (21, 5))
    assert 0 < len(vals) < 105 and pk[:, 0].max() <= 1.5 + 1e-12
    assert np.abs(vals - f3(pk)).max() < 1e-12

