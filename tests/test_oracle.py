"""CPU tests that pin the oracle (oracle/fus_oracle.c) before it is trusted as the checker.

1. the reference's own known-answer test for contract/transpose (cpp/mwe/sum_factorisation);
2. the unmodified reference header (oracle/_ref) vs the plain-C restatement, operator level;
3. closed-form known answers for the Basix/DOLFINx-supplied tables;
4. an independent dense O(N^6) evaluation of the same weak form (tests/dense_ref.py);
5. exactness identities (K 1 = 0, sum(M 1) = volume*coef, symmetry);
6. the committed golden fixtures in tests/golden/.
"""
import json
import os

import numpy as np
import pytest

from conftest import rel_l2, warp_vertices
import dense_ref

GOLD = os.path.join(os.path.dirname(__file__), "golden")
KAT = [4, 5, 6, 7, 12, 17, 22, 27, 20, 29, 38, 47]   # cpp/mwe/sum_factorisation


def test_kat_contract_transpose_restatement(orc):
    x, d, c = np.arange(8.0), np.arange(6.0), np.zeros(12)
    orc.lib.fo_contract(2, 3, 2, 2, 1, d, x, c)
    assert c.tolist() == KAT
    # equals numpy.tensordot as in cpp/mwe/sum_factorisation/main.py
    T = np.tensordot(d.reshape(3, 2), x.reshape(2, 2, 2), axes=[1, 0])
    assert np.array_equal(c.reshape(3, 2, 2), T)
    ct = np.zeros(12)
    orc.lib.fo_transpose(3, 2, 2, 2, 1, 6, c, ct)        # transpose<double,3,2,2, 2,1,6>
    assert np.array_equal(ct.reshape(2, 3, 2), np.transpose(T, [2, 0, 1]).reshape(2, 3, 2)) or \
        np.array_equal(ct, np.transpose(T, [2, 0, 1]).reshape(-1))


def test_kat_reference_header(orc_ref):
    out, out_t = np.zeros(12), np.zeros(12)
    orc_ref.lib.fr_kat(out, out_t)
    assert out.tolist() == KAT
    with open(os.path.join(GOLD, "kat_sum_factorisation.json")) as f:
        g = json.load(f)
    assert out.tolist() == g["out"] and out_t.tolist() == g["out_transposed"]


@pytest.mark.parametrize("N", [2, 3, 4, 5, 6, 7, 8])
def test_contract_transpose_vs_reference_header(orc, orc_ref, N):
    rng = np.random.default_rng(N)
    A, B = rng.standard_normal(N * N), rng.standard_normal(N ** 3)
    for tr in (0, 1):
        c0, c1 = rng.standard_normal(N ** 3), None
        c1 = c0.copy()
        orc.lib.fo_contract(N, N, N, N, tr, A, B, c0)
        orc_ref.lib.fr_contract_cube(N, tr, A, B, c1)
        assert rel_l2(c0, c1) < 1e-15        # -Ofast may fuse multiply-add
    t0, t1 = np.zeros(N ** 3), np.zeros(N ** 3)
    orc.lib.fo_transpose(N, N, N, N, N * N, 1, B, t0)
    orc_ref.lib.fr_transpose_cube(N, 0, B.copy(), t1)
    assert np.array_equal(t0, t1)
    assert np.array_equal(t0.reshape(N, N, N), B.reshape(N, N, N).transpose(1, 0, 2))
    orc.lib.fo_transpose(N, N, N, 1, N, N * N, B, t0)
    orc_ref.lib.fr_transpose_cube(N, 1, B.copy(), t1)
    assert np.array_equal(t0, t1)
    assert np.array_equal(t0.reshape(N, N, N), B.reshape(N, N, N).transpose(2, 1, 0))


def test_gll_known_answers(orc):
    p, w = orc.gll(3)                                        # P=2
    assert np.allclose(p, [0, 1, 0.5], atol=1e-16) and np.allclose(w, [1 / 6, 1 / 6, 2 / 3], atol=1e-16)
    p, w = orc.gll(5)                                        # P=4: {+-1, +-sqrt(3/7), 0} on [-1,1]
    s = np.sqrt(3 / 7)
    assert np.allclose(p, [0, 1, 0.5 - 0.5 * s, 0.5, 0.5 + 0.5 * s], atol=1e-15)
    assert np.allclose(w, np.array([1 / 10, 1 / 10, 49 / 90, 32 / 45, 49 / 90]) / 2, atol=1e-15)
    for m in range(2, 12):
        p, w = orc.gll(m)
        assert abs(w.sum() - 1) < 1e-14 and p[0] == 0 and p[1] == 1
        assert np.all(np.diff(p[2:]) > 0)
        pd, wd = dense_ref.gll_basix_order(m)
        assert np.allclose(p, pd, atol=1e-14) and np.allclose(w, wd, atol=1e-14)
        # exact for polynomials up to degree 2m-3
        for k in range(2 * m - 2):
            assert abs((w * p ** k).sum() - 1 / (k + 1)) < 1e-14


@pytest.mark.parametrize("P", range(1, 8))
def test_dphi_properties(orc, P):
    N = P + 1
    D = orc.dphi(P).reshape(N, N)
    p, _ = orc.gll(N)
    assert np.abs(D.sum(1)).max() < 1e-12                    # derivative of a constant
    assert np.allclose(D @ p, 1.0, atol=1e-12)               # derivative of x
    if P >= 2:
        assert np.allclose(D @ p ** 2, 2 * p, atol=1e-11)
    _, Dd = dense_ref.lagrange_tables(p)
    assert np.allclose(D, Dd, rtol=0, atol=1e-10 * np.abs(D).max())


def test_affine_cell_geometry(orc):
    P, h = 4, 0.25
    xg, xd = orc.box_mesh((2, 2, 2), (0, 0, 0), (2 * h, 2 * h, 2 * h))
    G, dJ = orc.geometry(P, xg, xd)
    _, w = orc.gll(P + 1)
    w3 = np.einsum("a,b,c->abc", w, w, w).reshape(-1)
    assert np.allclose(dJ, h ** 3 * w3[None, :], rtol=1e-14)
    # G = h w_q I  
    for k, val in enumerate([1, 0, 0, 1, 0, 1]):
        assert np.allclose(G[:, :, k], val * h * w3[None, :], rtol=1e-13, atol=1e-16)


@pytest.mark.parametrize("P", [1, 2, 3, 4])
def test_stiffness_vs_dense(orc, P):
    n = (2, 1, 2)
    xg, xd = orc.box_mesh(n)
    xg = warp_vertices(xg, amp=0.1, seed=P)
    dm = orc.box_dofmap(P, n, 0)
    nd = dm.max() + 1
    G, _ = orc.geometry(P, xg, xd)
    rng = np.random.default_rng(P)
    coeffs, x = rng.uniform(0.5, 2, dm.shape[0]), rng.uniform(-1, 1, nd)
    y = orc.stiffness_apply(P, dm, G, orc.dphi(P), coeffs, x, np.zeros(nd))
    yd = dense_ref.dense_stiffness_apply(P, xg, xd, dm, coeffs, x)
    assert rel_l2(y, yd) < 1e-12


@pytest.mark.parametrize("P,mode", [(2, 0), (3, 1), (5, 1), (7, 0)])
def test_operator_identities_and_ref_kernels(orc, orc_ref, P, mode):
    n = (3, 2, 2)
    xg, xd = orc.box_mesh(n, (0, 0, 0), (1.5, 1.0, 1.0))
    xg = warp_vertices(xg, amp=0.08, seed=3)
    dm = orc.box_dofmap(P, n, mode)
    nd = dm.max() + 1
    assert len(np.unique(dm)) == nd == (3 * P + 1) * (2 * P + 1) ** 2
    G, dJ = orc.geometry(P, xg, xd)
    dphi = orc.dphi(P)
    rng = np.random.default_rng(5)
    coeffs = rng.uniform(0.5, 2, dm.shape[0])
    x, z = rng.uniform(-1, 1, nd), rng.uniform(-1, 1, nd)
    # K 1 = 0
    y1 = orc.stiffness_apply(P, dm, G, dphi, coeffs, np.ones(nd), np.zeros(nd))
    assert np.abs(y1).max() < 1e-12 * np.abs(G).max() * P ** 2 * 50
    # symmetry  z.Kx = x.Kz
    kx = orc.stiffness_apply(P, dm, G, dphi, coeffs, x, np.zeros(nd))
    kz = orc.stiffness_apply(P, dm, G, dphi, coeffs, z, np.zeros(nd))
    assert abs(z @ kx - x @ kz) < 1e-12 * abs(z @ kx)
    # accumulate semantics (spectral_op.hpp:240-241): y += Kx
    y = orc.stiffness_apply(P, dm, G, dphi, coeffs, x, z.copy())
    assert rel_l2(y, z + kx) < 1e-15
    # mass: sum(M 1) = sum_c coeff_c vol_c
    m1 = orc.mass_apply(P, dm, dJ, coeffs, np.ones(nd), np.zeros(nd))
    assert abs(m1.sum() - (coeffs * dJ.sum(1)).sum()) < 1e-13 * m1.sum()
    # the same cell loops on the reference's own kernels (threads 1 and 3: "ranks")
    for nt in (1, 3):
        orc_ref.lib.fr_set_threads(nt)
        yr = orc_ref.stiffness_apply(P, dm, G, dphi, coeffs, x, np.zeros(nd), use_ref_kernels=True)
        assert rel_l2(yr, kx) < 1e-14
        mr = orc_ref.mass_apply(P, dm, dJ, coeffs, x, np.zeros(nd), use_ref_kernels=True)
        assert rel_l2(mr, orc.mass_apply(P, dm, dJ, coeffs, x, np.zeros(nd))) < 1e-15
    orc_ref.lib.fr_set_threads(1)


def test_facet_data_on_box(orc):
    P, n = 3, (2, 3, 2)
    lo, hi = (0, 0, 0), (1.0, 0.6, 0.8)
    xg, xd = orc.box_mesh(n, lo, hi)
    facets = orc.box_facets(n)
    assert facets.shape[0] == 2 * (3 * 2 + 2 * 2 + 2 * 3)
    fn, fs = orc.facet_data(P, xg, xd, facets)
    # total surface area
    assert abs(fs.sum() - 2 * (1.0 * 0.6 + 0.6 * 0.8 + 1.0 * 0.8)) < 1e-13
    tag1 = facets[:, 2] == 1
    assert abs(fs[tag1].sum() - 0.6 * 0.8) < 1e-14
    # facet nodes lie on the right plane
    dm = orc.box_dofmap(P, n, 0)
    My, Mz = 3 * P + 1, 2 * P + 1
    for k in np.where(tag1)[0]:
        gx = dm[facets[k, 0], fn[k]] // (My * Mz)
        assert np.all(gx == 0)


@pytest.mark.parametrize("P", [2, 3, 4, 5])
def test_golden_operators(orc, P):
    g = np.load(os.path.join(GOLD, f"stiffness_P{P}.npz"))
    y = orc.stiffness_apply(P, g["dofmap"], g["G"], g["dphi"], g["coeffs"], g["x"], g["y0"].copy())
    assert rel_l2(y, g["y"]) < 1e-14
    # the stored tables are what the oracle regenerates
    G, dJ = orc.geometry(P, g["xg"], g["xd"])
    assert rel_l2(G, g["G"]) < 1e-15 and rel_l2(dJ, g["detJ"]) < 1e-15
    assert np.allclose(orc.dphi(P), g["dphi"], rtol=0, atol=1e-14)   # _ref is built -Ofast
    m = np.load(os.path.join(GOLD, f"mass_P{P}.npz"))
    ym = orc.mass_apply(P, m["dofmap"], m["detJ"], m["coeffs"], m["x"], m["y0"].copy())
    assert rel_l2(ym, m["y"]) < 1e-15


@pytest.mark.parametrize("kind", ["linear", "lossy", "westervelt"])
def test_golden_rk4(orc, kind):
    g = np.load(os.path.join(GOLD, f"rk4_{kind}.npz"))
    P = int(g["P"])
    nd = g["u"].shape[0]
    mdl = orc.model(kind, P, nd, g["dofmap"], g["G"], g["detJ"], g["dphi"], g["c0"], g["rho0"],
                    g["delta0"], g["beta0"], g["facets"], g["fnodes"], g["fscale"],
                    float(g["freq"]), float(g["p0"]), float(g["s0"]))
    assert rel_l2(mdl.mass(), g["mass"]) < 1e-15
    kv = mdl.f1(float(g["f1_t"]), g["u_init"].copy(), g["v_init"].copy())
    assert rel_l2(kv, g["f1"]) < 1e-13
    u, v = g["u_init"].copy(), g["v_init"].copy()
    steps = mdl.rk4(float(g["t0"]), float(g["tf"]), float(g["dt"]), u, v)
    assert steps == int(g["steps"]) == 12
    assert rel_l2(u, g["u"]) < 1e-12 and rel_l2(v, g["v"]) < 1e-12


def test_rk4_plane_wave_physics(orc):
    """Linear model on a thin column reproduces p0 sin(w(t - x/c)) H(t - x/c) (the analytic
    solution used by python/tests/test_linearspectral_1d.py:72-90) once the window has opened."""
    P, nx = 4, 24
    c, rho, f, p0 = 1500.0, 1000.0, 0.5e6, 1.0
    lam = c / f
    Lx = 6 * lam
    n = (nx, 1, 1)
    h = Lx / nx
    xg, xd = orc.box_mesh(n, (0, 0, 0), (Lx, h, h))
    dm = orc.box_dofmap(P, n, 0)
    nd = dm.max() + 1
    G, dJ = orc.geometry(P, xg, xd)
    facets = orc.box_facets(n)
    fn, fs = orc.facet_data(P, xg, xd, facets)
    nc = dm.shape[0]
    mdl = orc.model("linear", P, nd, dm, G, dJ, orc.dphi(P), np.full(nc, c), np.full(nc, rho),
                    None, None, facets, fn, fs, f, p0, c)
    dt = 0.3 * h / (c * P * P)
    tf = 5.0 / f            # window (4 periods) fully open, front at 5 wavelengths < Lx
    u, v = np.zeros(nd), np.zeros(nd)
    mdl.rk4(0.0, tf, dt, u, v)
    # dof x-coordinates for the lexicographic numbering
    pts, _ = orc.gll(P + 1)
    pos = np.concatenate([[0, 1], np.arange(2, P + 1)])
    xs = np.zeros(nd)
    My = Mz = P + 1
    for cx in range(nx):
        for i0 in range(P + 1):
            xs[dm[cx].reshape(P + 1, P + 1, P + 1)[i0]] = (cx + pts[i0]) * h
    # compare in the region the fully-windowed wave has reached: x < c (tf - 4/f)
    sel = xs < 0.9 * c * (tf - 4.0 / f)
    exact = p0 * np.sin(2 * np.pi * f * (tf - xs / c))
    err = np.sqrt(((u - exact)[sel] ** 2).sum() / (exact[sel] ** 2).sum())
    assert err < 2e-2


def _plane_wave_error(orc, P, nx):
    """Relative L2 error of LinearSpectral3D on a column of nx cells (6 wavelengths) against
    p0 sin(w (t - x/c)) behind the front, small time step (spatial error dominates)."""
    c, rho, f, p0 = 1500.0, 1000.0, 0.5e6, 1.0
    Lx = 6 * c / f
    n, h = (nx, 1, 1), Lx / nx
    xg, xd = orc.box_mesh(n, (0, 0, 0), (Lx, h, h))
    dm = orc.box_dofmap(P, n, 0)
    nd = dm.max() + 1
    G, dJ = orc.geometry(P, xg, xd)
    facets = orc.box_facets(n)
    fn, fs = orc.facet_data(P, xg, xd, facets)
    nc = dm.shape[0]
    mdl = orc.model("linear", P, nd, dm, G, dJ, orc.dphi(P), np.full(nc, c), np.full(nc, rho),
                    None, None, facets, fn, fs, f, p0, c)
    dt, tf = 0.1 * h / (c * P * P), 5.0 / f
    u, v = np.zeros(nd), np.zeros(nd)
    mdl.rk4(0.0, tf, dt, u, v)
    pts, _ = orc.gll(P + 1)
    xs = np.zeros(nd)
    for cx in range(nx):
        for i0 in range(P + 1):
            xs[dm[cx].reshape(P + 1, P + 1, P + 1)[i0]] = (cx + pts[i0]) * h
    sel = xs < 0.9 * c * (tf - 4.0 / f)
    exact = p0 * np.sin(2 * np.pi * f * (tf - xs / c))
    return np.sqrt(((u - exact)[sel] ** 2).sum() / (exact[sel] ** 2).sum())


def test_plane_wave_spectral_convergence(orc):
    """An anchor outside this repository for the numbers Basix/DOLFINx/FFCx would supply (GLL
    rule, derivative table, geometry, the lumped boundary terms): halving h must reduce the error
    against the analytic plane wave (python/tests/test_linearspectral_1d.py:72-90) at the
    dispersion rate h^(2P) of a GLL spectral element method -- 2^4 for P = 2, 2^8 for P = 4.
    Wrong nodes, weights or derivative entries destroy that rate."""
    e2 = [_plane_wave_error(orc, 2, nx) for nx in (24, 48)]
    assert e2[1] < 1e-3 and 10.0 < e2[0] / e2[1] < 40.0             # measured 18.5
    e4 = [_plane_wave_error(orc, 4, nx) for nx in (12, 24)]
    assert e4[1] < 5e-5 and e4[0] / e4[1] > 150.0                   # measured 412


def _column(orc, P, nx, L):
    """A column of nx hexahedra along x whose only boundary facets are the two end faces (tag 1 at
    x = 0, tag 2 at x = L): the lateral walls stay natural, which makes the 3-D models one-dimensional,
    the setting of the reference's python/tests/*_1d.py."""
    n, h = (nx, 1, 1), L / nx
    xg, xd = orc.box_mesh(n, (0, 0, 0), (L, h, h))
    dm = orc.box_dofmap(P, n, 0)
    nd = dm.max() + 1
    G, dJ = orc.geometry(P, xg, xd)
    facets = orc.box_facets(n)
    facets = np.ascontiguousarray(facets[facets[:, 2] > 0])
    fn, fs = orc.facet_data(P, xg, xd, facets)
    pts, _ = orc.gll(P + 1)
    xs = np.zeros(nd)
    for cx in range(nx):
        for i0 in range(P + 1):
            xs[dm[cx].reshape(P + 1, P + 1, P + 1)[i0]] = (cx + pts[i0]) * h
    return dm, nd, G, dJ, facets, fn, fs, xs, h


def test_reference_analytic_test_lossy(orc):
    """python/tests/test_lossyspectral_1d.py:13-118 (degree 4, 4 elements per wavelength, f0 = 10,
    c0 = 1, rho0 = 4, 5 dB/m, CFL 0.5, t_end = L/c0 + 16/f0) with the oracle's LossySpectral3D:
    the attenuated plane wave p0 exp(-alpha x) sin(w t - w x / c0) within the reference's own
    threshold of 1e-2.  Covers what the linear test cannot: the second stiffness term, the
    d g/dt source term, the boundary mass term and the source factor 2 of Lossy.hpp:216."""
    P, epw = 4, 4
    f0, c0, rho0, alphadB, L = 10.0, 1.0, 4.0, 5.0, 1.0
    w0 = 2 * np.pi * f0
    aNp = alphadB / 20 * np.log(10)
    delta0 = 2 * aNp * c0 ** 3 / w0 / w0            # compute_diffusivity_of_sound, Lossy.hpp:376-380
    p0 = rho0 * c0 * 1.0
    nx = int(epw * L / (c0 / f0) + 1)
    dm, nd, G, dJ, facets, fn, fs, xs, h = _column(orc, P, nx, L)
    nc = dm.shape[0]
    mdl = orc.model("lossy", P, nd, dm, G, dJ, orc.dphi(P), np.full(nc, c0), np.full(nc, rho0),
                    np.full(nc, delta0), None, facets, fn, fs, f0, p0, c0)
    dt, tend = 0.5 * h / (c0 * P * P), L / c0 + 16 / f0
    u, v = np.zeros(nd), np.zeros(nd)
    mdl.rk4(0.0, tend, dt, u, v)
    exact = p0 * np.exp(-aNp * xs) * np.sin(w0 * tend - w0 / c0 * xs)
    assert np.linalg.norm(u - exact) / np.linalg.norm(exact) < 1e-2      # measured 5.6e-3


def test_reference_analytic_test_westervelt(orc):
    """python/tests/test_westerveltspectral_1d.py:13-128 (degree 4, 8 elements per wavelength,
    beta = 0.01, delta = 0, CFL 0.9, t_end = L/c0 + 8/f0) with the oracle's
    WesterveltSpectral3D: the Fubini solution.  The reference accepts 1e-1; the linear solution
    alone is 0.18 away, so the two nonlinear mass terms of Westervelt.hpp:249-265 are what is
    being checked."""
    from scipy.special import jv
    P, epw = 4, 8
    f0, c0, rho0, beta0, L = 10.0, 1.0, 1.0, 0.01, 1.0
    w0, p0 = 2 * np.pi * f0, 1.0
    nx = int(epw * L / (c0 / f0) + 1)
    dm, nd, G, dJ, facets, fn, fs, xs, h = _column(orc, P, nx, L)
    nc = dm.shape[0]
    mdl = orc.model("westervelt", P, nd, dm, G, dJ, orc.dphi(P), np.full(nc, c0), np.full(nc, rho0),
                    np.zeros(nc), np.full(nc, beta0), facets, fn, fs, f0, p0, c0)
    dt, tend = 0.9 * h / (c0 * P * P), L / c0 + 8 / f0
    u, v = np.zeros(nd), np.zeros(nd)
    mdl.rk4(0.0, tend, dt, u, v)
    sigma = (xs + 1e-7) / (c0 ** 2 / w0 / beta0 / (p0 / rho0 / c0))
    exact = np.zeros(nd)
    for term in range(1, 50):
        exact += 2 / term / sigma * jv(term, term * sigma) * np.sin(term * w0 * (tend - xs / c0))
    exact *= p0
    linear = p0 * np.sin(w0 * (tend - xs / c0))
    err = np.linalg.norm(u - exact) / np.linalg.norm(exact)
    assert err < 1e-2 and np.linalg.norm(linear - exact) / np.linalg.norm(exact) > 0.1   # 3.8e-3


@pytest.mark.parametrize("P", [2, 3, 4, 5, 6, 7])
def test_gll_element_mass_matrix_is_diagonal(orc, P):
    """python/tests/test_element_mass_matrix.py:13-72: a GLL-variant Lagrange element integrated
    with the GLL rule of matching degree has a diagonal mass matrix (dimensions 1, 2, 3) -- the
    property that lets MassSpectral3D multiply nodal values by detJ[c][q] (spectral_op.hpp:75-85).
    Basis functions are built independently of the oracle (monomial coefficients)."""
    from dense_ref import lagrange_tables
    pts, wts = orc.gll(P + 1)
    phi, _ = lagrange_tables(pts)
    M1 = phi.T @ np.diag(wts) @ phi
    assert np.abs(M1 - np.diag(wts)).max() < 1e-12 and abs(wts.sum() - 1.0) < 1e-14
    if P <= 4:
        M2 = np.kron(M1, M1)
        assert np.abs(M2 - np.diag(np.diag(M2))).max() < 1e-12
    if P <= 3:
        M3 = np.kron(M2, M1)
        assert np.abs(M3 - np.diag(np.diag(M3))).max() < 1e-12
        assert np.allclose(np.diag(M3), np.einsum("a,b,c->abc", wts, wts, wts).reshape(-1))
