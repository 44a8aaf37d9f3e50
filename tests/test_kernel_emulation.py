"""The CUDA kernels of fenicsx-fus_b200/csrc/fus_kernels.cuh executed on the CPU by a SIMT emulator
(tests/emu/simt_emu.hpp: one OS thread per CUDA thread, real barriers, real atomics) and compared with
the oracle.  The build container has no GPU; this keeps the kernels' indexing, shared-memory staging,
software pipelining, barrier placement and tail handling under test in the CPU suite.  It says
nothing about the device's memory model or speed -- the `-m gpu` tests remain the parity tests."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT, rel_l2, warp_vertices

EMU_DIR = os.path.join(ROOT, "tests", "emu")
_f64 = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i32 = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_p, _ll, _int, _dbl = C.c_void_p, C.c_int64, C.c_int, C.c_double


def _opt(a):
    return None if a is None else a.ctypes.data_as(_p)


@pytest.fixture(scope="module")
def emu():
    lib = os.path.join(EMU_DIR, "libfus_emu.so")
    deps = [os.path.join(EMU_DIR, f) for f in ("emu_kernels.cpp", "simt_emu.hpp")] + [
        os.path.join(ROOT, "fenicsx-fus_b200", "csrc", f)
        for f in ("fus_kernels.cuh", "fus_trilinear.hpp", "fus_halo_kernels.cuh")]
    if not os.path.exists(lib) or os.path.getmtime(lib) < max(os.path.getmtime(d) for d in deps):
        subprocess.run(["/usr/bin/g++", "-std=c++20", "-O1", "-pthread", "-shared", "-fPIC",
                        "-I" + os.path.join(ROOT, "include"),
                        "-I" + os.path.join(ROOT, "fenicsx-fus_b200", "csrc"), "-I" + EMU_DIR,
                        os.path.join(EMU_DIR, "emu_kernels.cpp"), "-o", lib], check=True)
    L = C.CDLL(lib)
    L.emu_stiffness.argtypes = [_int, _int, _int, _f64, _p, _f64, _i32, _p, _p, _f64, _p, _ll, _f64,
                                _f64, _f64, _int, _ll, _ll]
    L.emu_stiffness_quad.argtypes = [_int, _f64, _p, _f64, _i32, _f64, _f64, _p, _ll, _f64, _f64,
                                     _f64, _int]
    L.emu_tri_coeffs_and_mass.argtypes = [_int, _f64, _i32, _ll, _f64, _f64, _f64, _i32, _f64, _f64,
                                          _f64]
    L.emu_mass.argtypes = [_f64, _f64, _i32, _f64, _f64, _ll, _int]
    f32 = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
    L.emu_operators_f32.argtypes = [_int, f32, f32, f32, _i32, _f64, _f64, f32, _ll, _f64, _f64, _f64,
                                    _int]
    u64 = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
    i64 = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
    L.emu_halo_pack.argtypes = [_f64, _p, _i32, i64, _int, _f64, _ll, _int]
    L.emu_halo_unpack.argtypes = [_int, _f64, _p, _i32, i64, _int, _f64, _ll, _int]
    L.emu_fused_create.argtypes = [_ll, _int, i64, _i32, i64, _i32, _p, i64, _dbl]
    L.emu_fused_create.restype = _p
    L.emu_fused_connect.argtypes = [_p, _int, _p, i64]
    L.emu_fused_nshared.argtypes = [_p]
    L.emu_fused_nshared.restype = _ll
    L.emu_fused_state.argtypes = [_p, u64, np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")]
    L.emu_fused_destroy.argtypes = [_p]
    L.emu_fused_destroy.restype = None
    L.emu_fused_ready.argtypes = [_p, _int]
    L.emu_fused_entry_put.argtypes = [_p, _f64, _f64]
    L.emu_fused_exit.argtypes = [_p, _f64, _f64]
    L.emu_stiffness_fused.argtypes = [_int, _int, _f64, _p, _f64, _i32, _f64, _f64, _p, _ll, _f64,
                                      _f64, _f64, _int, _p, _ll]
    L.emu_fused_forward_landed.argtypes = [_p]
    L.emu_geometry.argtypes = [_int, _f64, _i32, _ll, _f64, _f64, _f64, _f64, _f64, C.POINTER(_int)]
    L.emu_geometry_quad.argtypes = [_int, _f64, _i32, _ll, _f64, _f64, _f64, _f64]
    L.emu_rk4_stage.argtypes = [_int, _int, _f64, _f64, _p, _f64, _f64, _f64, _f64, _f64, _f64, _ll,
                                _ll, _dbl, _ll, _p, _p, _p, _p, _p, _dbl, _dbl, _int, _int, _p]
    L.emu_boundary.argtypes = [_f64, _f64, _i32, _f64, _f64, _f64, _ll, _dbl, _dbl]
    return L


def _case(fus, orc, P, n=(5, 3, 2), warp=True, numbering=1):
    m = fus.BoxMesh(n, (0.2, -0.1, 0.3), (1.2, 0.5, 0.7),
                    warp=(lambda x: warp_vertices(x, 0.08, 3)) if warp else None)
    V = fus.FunctionSpace(m, P, numbering=numbering)
    G, dJ = orc.geometry(P, m.x, m.xdofmap)
    pts, wts = orc.gll(P + 1)
    return m, V, G, dJ, pts, wts


@pytest.mark.parametrize("variant", [0, 1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("P", [1, 2, 3, 4, 5, 6, 7])
def test_emulated_stiffness_kernels_vs_oracle(fus, orc, emu, P, variant):
    """stiffness_col_kernel / stiffness_point_kernel / stiffness_line_kernel with streamed G, plain
    and fused two-vector gather, 30 warped cells on at most 2 blocks (grid-stride loop, ragged tail,
    the look-ahead pipeline across cells).  Variants 3-6 are the line kernel's experimental software
    pipelines (coefficient folded into x; dofmap prefetched / loaded with the G refills; G through
    the TMA-fed shared-memory ring, emulated with synchronous copies and phase counters)."""
    m, V, G, dJ, pts, wts = _case(fus, orc, P, numbering=P % 2)
    nd, nc = V.ndofs, m.ncells
    rng = np.random.default_rng(100 * P + variant)
    x, x2 = rng.uniform(-1, 1, nd), rng.uniform(-1, 1, nd)
    c1, c2 = rng.uniform(0.5, 2, nc), rng.uniform(-1, 1, nc)
    dphi = orc.dphi(P)
    y = np.zeros(nd)
    assert emu.emu_stiffness(P + 1, variant, 0, x, None, y, V.dofmap, _opt(G), None, c1, None, nc,
                             dphi, pts, wts, 2, 0, nc) == 0
    yo = orc.stiffness_apply(P, V.dofmap, G, dphi, c1, x, np.zeros(nd))
    assert rel_l2(y, yo) < 1e-13
    yf = np.zeros(nd)
    assert emu.emu_stiffness(P + 1, variant, 0, x, _opt(x2), yf, V.dofmap, _opt(G), None, c1,
                             _opt(c2), nc, dphi, pts, wts, 2, 0, nc) == 0
    yfo = orc.stiffness_apply(P, V.dofmap, G, dphi, c1, x, np.zeros(nd))
    yfo = orc.stiffness_apply(P, V.dofmap, G, dphi, c2, x2, yfo)
    assert rel_l2(yf, yfo) < 1e-13
    # integer data: gather/scatter indexing bit-exact (sums of small integers are exact)
    xi = rng.integers(-4, 5, nd).astype(np.float64)
    Gi = np.zeros_like(G)
    Gi[:, :, [0, 3, 5]] = 1.0
    di = np.rint(4 * dphi) / 4                                   # exactly representable table
    yi = np.zeros(nd)
    emu.emu_stiffness(P + 1, variant, 0, xi, None, yi, V.dofmap, _opt(Gi), None, np.ones(nc), None,
                      nc, di, pts, wts, 2, 0, nc)
    assert np.array_equal(yi, orc.stiffness_apply(P, V.dofmap, Gi, di, np.ones(nc), xi, np.zeros(nd)))


@pytest.mark.parametrize("variant,P", [(0, 2), (0, 3), (2, 4), (2, 5), (2, 6), (2, 7), (1, 3),
                                       (3, 4), (4, 5), (5, 2), (5, 4), (5, 6), (5, 7),
                                       (6, 1), (6, 3), (6, 4), (6, 5), (6, 6)])
def test_emulated_split_launches_of_a_partitioned_stage(fus, orc, emu, variant, P):
    """A partitioned stage applies the cells in three launches -- interior A [ni, mid), interface
    [0, ni), interior B [mid, nc) (assemble_rhs in csrc/fus_capi.cu) -- with sub-ranges that are
    not multiples of the cells-per-block packing; together they must equal one full application."""
    m, V, G, dJ, pts, wts = _case(fus, orc, P)
    nd, nc = V.ndofs, m.ncells
    rng = np.random.default_rng(7 * P + variant)
    x, c1 = rng.uniform(-1, 1, nd), rng.uniform(0.5, 2, nc)
    dphi = orc.dphi(P)
    ni, mid = 7, 7 + (nc - 7) // 2
    y = np.zeros(nd)
    for cb, ce in ((ni, mid), (0, ni), (mid, nc)):
        assert emu.emu_stiffness(P + 1, variant, 0, x, None, y, V.dofmap, _opt(G), None, c1, None, nc,
                                 dphi, pts, wts, 3, cb, ce) == 0
    assert rel_l2(y, orc.stiffness_apply(P, V.dofmap, G, dphi, c1, x, np.zeros(nd))) < 1e-13
    # an empty range launches nothing harmful
    y2 = y.copy()
    emu.emu_stiffness(P + 1, variant, 0, x, None, y2, V.dofmap, _opt(G), None, c1, None, nc, dphi, pts,
                      wts, 3, 5, 6)
    one = orc.stiffness_apply(P, V.dofmap[5:6], G[5:6], dphi, c1[5:6], x, np.zeros(nd))
    assert rel_l2(y2 - y, one) < 1e-12


@pytest.mark.parametrize("P", [1, 2, 4, 5, 7])
def test_emulated_compressed_geometry_kernels(fus, orc, emu, P):
    """stiffness_line_kernel<N,FUSE2,1> (one Ghat per affine cell) and <N,FUSE2,2> (G rebuilt from
    the trilinear cell map), tri_coeff_kernel and mass_tri_kernel."""
    from fenicsx_fus_b200 import capi
    rng = np.random.default_rng(P)
    dphi = orc.dphi(P)
    # mode 2 on warped cells
    m, V, G, dJ, pts, wts = _case(fus, orc, P)
    nd, nc = V.ndofs, m.ncells
    x, x2 = rng.uniform(-1, 1, nd), rng.uniform(-1, 1, nd)
    c1, c2 = rng.uniform(0.5, 2, nc), rng.uniform(-1, 1, nc)
    co_ref = np.zeros((nc, 24))
    assert capi.load().fus_trilinear_coeffs(nc, m.x, m.xdofmap, co_ref) == 0
    co, ym = np.zeros((nc, 24)), np.zeros(nd)
    assert emu.emu_tri_coeffs_and_mass(P + 1, m.x, m.xdofmap, nc, co, x, ym, V.dofmap, c1, pts,
                                       wts) == 0
    assert np.array_equal(co, co_ref)                             # same helper, same arithmetic
    assert rel_l2(ym, orc.mass_apply(P, V.dofmap, dJ, c1, x, np.zeros(nd))) < 1e-12
    for fuse, mode in ((False, 2), (True, 2), (False, 3)):   # 3: the same code under a register cap
        y = np.zeros(nd)
        emu.emu_stiffness(P + 1, 2, mode, x, _opt(x2) if fuse else None, y, V.dofmap, None, _opt(co),
                          c1, _opt(c2) if fuse else None, nc, dphi, pts, wts, 2, 0, nc)
        yo = orc.stiffness_apply(P, V.dofmap, G, dphi, c1, x, np.zeros(nd))
        if fuse:
            yo = orc.stiffness_apply(P, V.dofmap, G, dphi, c2, x2, yo)
        assert rel_l2(y, yo) < 1e-12
    # mode 1 on a sheared box: Ghat = G / w at any point of the cell
    A = np.array([[1.0, 0.3, 0.1], [0.0, 0.8, 0.25], [0.05, 0.0, 1.2]])
    ma = fus.BoxMesh((4, 3, 2), (0, 0, 0), (1.0, 0.6, 0.5), warp=lambda z: z @ A.T)
    Va = fus.FunctionSpace(ma, P, numbering=1)
    Ga, _ = orc.geometry(P, ma.x, ma.xdofmap)
    ghat = np.ascontiguousarray(Ga[:, 0, :] / (wts[0] ** 3))
    xa, ca = rng.uniform(-1, 1, Va.ndofs), rng.uniform(0.5, 2, ma.ncells)
    ya = np.zeros(Va.ndofs)
    emu.emu_stiffness(P + 1, 2, 1, xa, None, ya, Va.dofmap, None, _opt(ghat), ca, None, ma.ncells,
                      dphi, pts, wts, 2, 0, ma.ncells)
    assert rel_l2(ya, orc.stiffness_apply(P, Va.dofmap, Ga, dphi, ca, xa, np.zeros(Va.ndofs))) < 1e-12
    xb, cb2 = rng.uniform(-1, 1, Va.ndofs), rng.uniform(-1, 1, ma.ncells)
    yb = np.zeros(Va.ndofs)                                       # affine + fused two-vector gather
    emu.emu_stiffness(P + 1, 2, 1, xa, _opt(xb), yb, Va.dofmap, None, _opt(ghat), ca, _opt(cb2),
                      ma.ncells, dphi, pts, wts, 2, 0, ma.ncells)
    yref = orc.stiffness_apply(P, Va.dofmap, Ga, dphi, ca, xa, np.zeros(Va.ndofs))
    yref = orc.stiffness_apply(P, Va.dofmap, Ga, dphi, cb2, xb, yref)
    assert rel_l2(yb, yref) < 1e-12


@pytest.mark.parametrize("P", [1, 2, 3, 4, 5, 6, 7])
def test_emulated_quad_kernel_vs_oracle(fus, orc, emu, P):
    """stiffness_quad_kernel<N,FUSE2> (2-D variant): 35 warped cells on 2 blocks, so the block loop
    with its trailing barrier iterates several times."""
    rng = np.random.default_rng(P)
    h = np.array([1.2, 0.9]) / np.array([7, 5])

    def warp(z):
        w = z.copy()
        w[:, :2] += 0.1 * h * rng.uniform(-1, 1, (z.shape[0], 2))
        return w
    m = fus.RectMesh((7, 5), (0.1, -0.2), (1.3, 0.7), warp=warp)
    V = fus.FunctionSpace(m, P)
    nd, nc = V.ndofs, m.ncells
    G, dJ = orc.geometry_2d(P, m.x, m.xdofmap)
    Gq = np.ascontiguousarray(G.transpose(0, 2, 1))              # device layout Gq[cell][p][q]
    pts, wts = orc.gll(P + 1)
    x, x2 = rng.uniform(-1, 1, nd), rng.uniform(-1, 1, nd)
    c1, c2 = rng.uniform(0.5, 2, nc), rng.uniform(-1, 1, nc)
    y = np.zeros(nd)
    assert emu.emu_stiffness_quad(P + 1, x, None, y, V.dofmap, Gq, c1, None, nc, orc.dphi(P), pts,
                                  wts, 2) == 0
    yo = orc.stiffness_apply_2d(P, V.dofmap, G, orc.dphi(P), c1, x, np.zeros(nd))
    assert rel_l2(y, yo) < 1e-13
    yf = np.zeros(nd)
    emu.emu_stiffness_quad(P + 1, x, _opt(x2), yf, V.dofmap, Gq, c1, _opt(c2), nc, orc.dphi(P), pts,
                           wts, 2)
    yfo = orc.stiffness_apply_2d(P, V.dofmap, G, orc.dphi(P), c2, x2, yo.copy())
    assert rel_l2(yf, yfo) < 1e-13


def test_emulated_mass_boundary_and_rk4_stage_kernels(fus, orc, emu):
    """mass_kernel, boundary_kernel and the four fused RK4 epilogues (linear and Westervelt) against
    the formulas they fuse (Linear.hpp:203-221,279-294; Westervelt.hpp:249-265)."""
    P = 3
    m, V, G, dJ, pts, wts = _case(fus, orc, P)
    nd, nc = V.ndofs, m.ncells
    rng = np.random.default_rng(9)
    x, c1 = rng.uniform(-1, 1, nd), rng.uniform(0.5, 2, nc)
    y = np.zeros(nd)
    emu.emu_mass(x, y, np.ascontiguousarray(V.dofmap.reshape(-1)), np.ascontiguousarray(dJ.reshape(-1)),
                 c1, nc * (P + 1) ** 3, (P + 1) ** 3)
    assert rel_l2(y, orc.mass_apply(P, V.dofmap, dJ, c1, x, np.zeros(nd))) < 1e-14
    # boundary terms over a compacted list
    nb = 300
    bidx = np.sort(rng.choice(nd, nb, replace=False)).astype(np.int32)
    bs, bd, ba = rng.uniform(0, 1, nb), rng.uniform(0, 1, nb), rng.uniform(0, 1, nb)
    b, v = rng.uniform(-1, 1, nd), rng.uniform(-1, 1, nd)
    want = b.copy()
    want[bidx] += 0.7 * bs + (-0.3) * bd - ba * v[bidx]
    emu.emu_boundary(b, v, bidx, bs, bd, ba, nb, 0.7, -0.3)
    assert np.allclose(b, want, rtol=1e-15, atol=0)
    # RK4 epilogues against the reference's running sums (Linear.hpp:278-295): stage inputs after
    # every stage, the new state after stage 3; b re-seeded with the next stage's boundary terms on
    # the owned boundary dofs and zero elsewhere (ghost entries included); ghosts of the state
    # vectors untouched.  Odd owned count and a grid smaller than the chunk count: the straddling
    # chunk and the grid-stride loop are exercised.  hints: the L2 evict-first build of the kernel.
    nowned = nd - 37
    dt = 1e-3
    a_r, b_r = (0.0, 0.5, 0.5, 1.0), (1 / 6, 1 / 3, 1 / 3, 1 / 6)
    chunk = emu.emu_rk4_stage(0, 0, np.zeros(8), np.ones(8), None, np.zeros(8), np.zeros(8),
                              np.zeros(8), np.zeros(8), np.zeros(8), np.zeros(8), 0, 8, dt, 0, None,
                              None, None, None, None, 0.0, 0.0, 0, 1, None)
    assert chunk > 0
    bown = bidx[bidx < nowned]
    nbo = bown.size
    bchunk = np.searchsorted(bown, np.arange(0, (nd + chunk - 1) // chunk + 1) * chunk).astype(np.int64)
    for west, hints in ((0, 0), (1, 0), (0, 1), (1, 1)):
        mvec, dnl = rng.uniform(1, 2, nd), rng.uniform(0.01, 0.02, nd)
        u0, v0 = rng.uniform(-1, 1, nd), rng.uniform(-1, 1, nd)
        st = dict(u0=u0.copy(), v0=v0.copy(), ua=np.zeros(nd), va=np.zeros(nd), un=np.zeros(nd),
                  vn=np.zeros(nd))
        ref = {k: a.copy() for k, a in st.items()}
        for i in range(4):
            b = rng.uniform(-1, 1, nd)
            o = slice(0, nowned)
            un_in = ref["u0"][o] if i == 0 else ref["un"][o]
            vn_in = ref["v0"][o] if i == 0 else ref["vn"][o]
            mm, bb = mvec[o].copy(), b[o].copy()
            if west:
                mm = mm - dnl[o] * un_in
                bb = bb + dnl[o] * vn_in * vn_in
            kv, ku = bb / mm, vn_in.copy()
            ua = (ref["u0"][o] if i == 0 else ref["ua"][o]) + b_r[i] * dt * ku
            va = (ref["v0"][o] if i == 0 else ref["va"][o]) + b_r[i] * dt * kv
            if i < 3:
                ref["un"][o] = ref["u0"][o] + a_r[i + 1] * dt * ku
                ref["vn"][o] = ref["v0"][o] + a_r[i + 1] * dt * kv
                ref["ua"][o], ref["va"][o] = ua, va
            else:
                ref["u0"][o], ref["v0"][o] = ua, va
            gn, dgn = 0.3 + i, -0.2 * (i + 1)
            rc = emu.emu_rk4_stage(i, west, b, mvec, _opt(dnl) if west else None, st["u0"], st["v0"],
                                   st["ua"], st["va"], st["un"], st["vn"], nowned, nd, dt, nbo,
                                   _opt(bown), _opt(bs), _opt(bd), _opt(ba), _opt(bchunk), gn, dgn,
                                   hints, 3, None)
            assert rc == chunk
            vnext = st["vn"] if i < 3 else st["v0"]
            want_b = np.zeros(nd)
            want_b[bown] = gn * bs[:nbo] + dgn * bd[:nbo] - ba[:nbo] * vnext[bown]
            assert np.allclose(b, want_b, rtol=1e-15, atol=0), (west, i)
            # the accumulators are internal now (written in stage 1, read in stage 3): compare the
            # stage inputs and the state
            for k in ("u0", "v0", "un", "vn"):
                assert np.allclose(st[k], ref[k], rtol=1e-14, atol=1e-15), (west, i, k)
            for k in st:                    # ghost entries are never written
                assert np.array_equal(st[k][nowned:], (u0 if k == "u0" else v0 if k == "v0"
                                                       else np.zeros(nd))[nowned:]), (west, i, k)
        # without a boundary list b is zero-filled
        b = rng.uniform(-1, 1, nd)
        emu.emu_rk4_stage(2, west, b, mvec, _opt(dnl) if west else None, st["u0"], st["v0"],
                          st["ua"], st["va"], st["un"], st["vn"], nowned, nd, dt, 0, None, None, None,
                          None, None, 0.0, 0.0, hints, 2, None)
        assert not b.any()


@pytest.mark.parametrize("P", [1, 2, 3, 4, 5, 6, 7])
def test_emulated_setup_kernels(fus, orc, emu, P):
    """geometry_kernel (G in the device layout, detJ), g_from_device_layout_kernel (back to the
    reference layout), affine_detect_kernel and geometry_quad_kernel against precompute.hpp as the
    oracle restates it."""
    m, V, G, dJ, pts, wts = _case(fus, orc, P, n=(3, 2, 2))
    nc = m.ncells
    Gk, dJk, ghat, flag = np.zeros_like(G), np.zeros_like(dJ), np.zeros((nc, 6)), C.c_int(-1)
    assert emu.emu_geometry(P + 1, m.x, m.xdofmap, nc, Gk, dJk, pts, wts, ghat, C.byref(flag)) == 0
    assert rel_l2(Gk, G) < 1e-14 and rel_l2(dJk, dJ) < 1e-14 and flag.value == 0     # warped: not affine
    A = np.array([[1.0, 0.3, 0.1], [0.0, 0.8, 0.25], [0.05, 0.0, 1.2]])
    ma = fus.BoxMesh((3, 2, 2), (0, 0, 0), (1.0, 0.6, 0.5), warp=lambda z: z @ A.T)
    Ga, dJa = orc.geometry(P, ma.x, ma.xdofmap)
    Gk, dJk = np.zeros_like(Ga), np.zeros_like(dJa)
    emu.emu_geometry(P + 1, ma.x, ma.xdofmap, ma.ncells, Gk, dJk, pts, wts, ghat, C.byref(flag))
    assert flag.value == 1 and rel_l2(Gk, Ga) < 1e-14
    assert np.allclose(ghat, Ga[:, 0, :] / wts[0] ** 3, rtol=1e-13)
    mq = fus.RectMesh((4, 3), (0.1, -0.2), (1.3, 0.7),
                      warp=lambda z: z + 0.03 * np.sin(7 * z[:, ::-1]) * np.array([1, 1, 0]))
    Gq_ref, dJq_ref = orc.geometry_2d(P, mq.x, mq.xdofmap)
    Gq, dJq = np.zeros((mq.ncells, 3, (P + 1) ** 2)), np.zeros_like(dJq_ref)
    assert emu.emu_geometry_quad(P + 1, mq.x, mq.xdofmap, mq.ncells, Gq, dJq, pts, wts) == 0
    assert rel_l2(Gq.transpose(0, 2, 1), Gq_ref) < 1e-14 and rel_l2(dJq, dJq_ref) < 1e-14


@pytest.mark.parametrize("P", [1, 2, 3, 4, 5, 6, 7])
def test_emulated_fp32_operators(fus, orc, emu, P):
    """stiffness_line_kernel<N,false,0,float> and mass_kernel_f32 on float copies of the cell data
    (SURVEY.md section 8f-4): the FP64 oracle on the same float-rounded inputs, to float accuracy."""
    m, V, G, dJ, pts, wts = _case(fus, orc, P)
    nd, nc = V.ndofs, m.ncells
    rng = np.random.default_rng(50 + P)
    x = rng.uniform(-1, 1, nd).astype(np.float32)
    c = rng.uniform(0.5, 2, nc).astype(np.float32)
    y, ym = np.zeros(nd, dtype=np.float32), np.zeros(nd, dtype=np.float32)
    assert emu.emu_operators_f32(P + 1, x, y, ym, np.ascontiguousarray(V.dofmap), G, dJ, c, nc,
                                 orc.dphi(P), pts, wts, 2) == 0
    x64, c64 = x.astype(np.float64), c.astype(np.float64)
    yo = orc.stiffness_apply(P, V.dofmap, G, orc.dphi(P), c64, x64, np.zeros(nd))
    mo = orc.mass_apply(P, V.dofmap, dJ, c64, x64, np.zeros(nd))
    assert rel_l2(y, yo) < 2e-6 and rel_l2(ym, mo) < 1e-6
    assert y.dtype == np.float32 and np.isfinite(y).all()


def test_emulation_with_scheduling_jitter(fus, orc, emu):
    """The stiffness kernels once more with every emulated thread sleeping a pseudo-random time after
    each barrier (FUS_EMU_JITTER): threads of a warp and of a block then run far apart, which makes
    a missing __syncwarp / __syncthreads / named barrier all but certain to corrupt the result."""
    import sys
    code = (
        "import os, sys, ctypes as C, numpy as np\n"
        f"sys.path.insert(0, {ROOT!r}); sys.path.insert(0, {os.path.join(ROOT, 'tests')!r})\n"
        "import fenicsx_fus_b200 as fus\n"
        "from conftest import warp_vertices, rel_l2\n"
        "from oracle.oracle import Oracle\n"
        "orc = Oracle()\n"
        f"L = C.CDLL({os.path.join(EMU_DIR, 'libfus_emu.so')!r})\n"
        "f64 = np.ctypeslib.ndpointer(dtype=np.float64, flags='C_CONTIGUOUS')\n"
        "i32 = np.ctypeslib.ndpointer(dtype=np.int32, flags='C_CONTIGUOUS')\n"
        "L.emu_stiffness.argtypes = [C.c_int] * 3 + [f64, C.c_void_p, f64, i32, C.c_void_p, C.c_void_p,"
        " f64, C.c_void_p, C.c_int64, f64, f64, f64, C.c_int, C.c_int64, C.c_int64]\n"
        "for P, variant in ((2, 0), (4, 2), (5, 2), (6, 2), (3, 1)):\n"
        "    m = fus.BoxMesh((3, 2, 2), warp=lambda x: warp_vertices(x, 0.08, 3))\n"
        "    V = fus.FunctionSpace(m, P, numbering=1)\n"
        "    G, dJ = orc.geometry(P, m.x, m.xdofmap)\n"
        "    pts, wts = orc.gll(P + 1)\n"
        "    rng = np.random.default_rng(P)\n"
        "    x, c = rng.uniform(-1, 1, V.ndofs), rng.uniform(0.5, 2, m.ncells)\n"
        "    y = np.zeros(V.ndofs)\n"
        "    L.emu_stiffness(P + 1, variant, 0, x, None, y, V.dofmap, G.ctypes.data_as(C.c_void_p), None,"
        " c, None, m.ncells, orc.dphi(P), pts, wts, 1, 0, m.ncells)\n"
        "    e = rel_l2(y, orc.stiffness_apply(P, V.dofmap, G, orc.dphi(P), c, x, np.zeros(V.ndofs)))\n"
        "    assert e < 1e-13, (P, variant, e)\n"
        "print('jitter ok')\n")
    env = dict(os.environ, FUS_EMU_JITTER="200")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env,
                         timeout=600)
    assert res.returncode == 0 and "jitter ok" in res.stdout, res.stdout + res.stderr


@pytest.mark.parametrize("kind", ["linear", "westervelt"])
def test_emulated_fused_rk4_step_vs_reference_flow(fus, orc, emu, kind):
    """The fused stage flow of fus_model_rk4 (csrc/fus_capi.cu: operator with the fused two-vector
    gather, rk4_stage_kernel<STAGE,WESTERVELT> seeding the next stage's boundary terms; 33 vector
    passes per step) executed
    with the emulated kernels, against the oracle's literal restatement of the reference loop
    (Linear.hpp:228-314, Westervelt.hpp:216-373; ~30 passes per stage).  Same fields after 3 steps."""
    from fenicsx_fus_b200 import capi
    P, n, h = 3, (3, 2, 2), 0.002
    m = fus.BoxMesh(n, (0, 0, 0), tuple(h * k for k in n), warp=lambda x: warp_vertices(x, 0.05, 5))
    V = fus.FunctionSpace(m, P, numbering=1)
    nd, nc = V.ndofs, m.ncells
    G, dJ = orc.geometry(P, m.x, m.xdofmap)
    pts, wts = orc.gll(P + 1)
    dphi = orc.dphi(P)
    rng = np.random.default_rng(3)
    c0, rho0 = rng.uniform(1400, 1600, nc), rng.uniform(900, 1100, nc)
    west = kind == "westervelt"
    delta = rng.uniform(1e-3, 3e-3, nc) if west else None
    beta = rng.uniform(3, 4, nc) if west else None
    f0, p0, s0 = 0.5e6, 6.0e4, 1500.0
    fn, fs = orc.facet_data(P, m.x, m.xdofmap, m.facets)
    om = orc.model(kind, P, nd, V.dofmap, G, dJ, dphi, c0, rho0, delta, beta, m.facets, fn, fs, f0, p0,
                   s0)
    dt = 0.2 * np.sqrt(3) * h / (1600.0 * P * P)
    u0, v0 = 1e3 * rng.uniform(-1, 1, nd), 1e9 * rng.uniform(-1, 1, nd)
    u_ref, v_ref = u0.copy(), v0.copy()
    steps, tf = 3, 2.5 * dt                # the reference loop shortens the last step to land on tf
    assert om.rk4(0.0, tf, dt, u_ref, v_ref) == steps
    # ---- model set-up as fus_model_create does it
    src, dsrc, absb, bmass = (np.zeros(nd) for _ in range(4))
    assert capi.load().fus_boundary_vectors(
        capi.KINDS[kind], P, nc, nd, m.x, m.xdofmap, V.dofmap, m.facets.shape[0], m.facets, c0, rho0,
        _opt(delta), capi.optional(src), capi.optional(dsrc), capi.optional(absb),
        capi.optional(bmass)) == 0
    Nd = (P + 1) ** 3
    dmf, dJf = np.ascontiguousarray(V.dofmap.reshape(-1)), np.ascontiguousarray(dJ.reshape(-1))
    mvec = np.zeros(nd)
    emu.emu_mass(np.ones(nd), mvec, dmf, dJf, 1.0 / rho0 / c0 ** 2, nc * Nd, Nd)
    mvec += bmass
    dnl = None
    if west:
        dnl = np.zeros(nd)
        emu.emu_mass(np.ones(nd), dnl, dmf, dJf, 2.0 * beta / rho0 ** 2 / c0 ** 4, nc * Nd, Nd)
    lin = -1.0 / rho0
    att = (-delta / rho0 / c0 ** 2) if west else None
    bidx = np.flatnonzero((src != 0) | (dsrc != 0) | (absb != 0)).astype(np.int32)
    bs, bd, ba = src[bidx].copy(), dsrc[bidx].copy(), absb[bidx].copy()
    chunk = emu.emu_rk4_stage(0, 0, np.zeros(8), np.ones(8), None, np.zeros(8), np.zeros(8),
                              np.zeros(8), np.zeros(8), np.zeros(8), np.zeros(8), 0, 8, 1.0, 0, None,
                              None, None, None, None, 0.0, 0.0, 0, 1, None)
    bchunk = np.searchsorted(bidx, np.arange(0, (nd + chunk - 1) // chunk + 1) * chunk).astype(np.int64)
    # ---- the time loop as fus_model_rk4 / issue_step issue it
    st = dict(u0=u0.copy(), v0=v0.copy(), ua=np.zeros(nd), va=np.zeros(nd), un=np.zeros(nd),
              vn=np.zeros(nd))
    a_r = (0.0, 0.5, 0.5, 1.0)
    w0, kappa = 2 * np.pi * f0, (2.0 if west else 1.0)

    def scalars(tn):                       # Linear.hpp:185-192, Lossy.hpp:199-220
        win = 0.5 * (1 - np.cos(f0 * np.pi * tn / 4.0)) if tn < 4.0 / f0 else 1.0
        dwin = 0.5 * np.pi * f0 / 4.0 * np.sin(f0 * np.pi * tn / 4.0) if tn < 4.0 / f0 else 0.0
        g = kappa * win * p0 * w0 / s0 * np.cos(w0 * tn)
        dg = (kappa * (dwin * p0 * w0 / s0 * np.cos(w0 * tn) - win * p0 * w0 * w0 / s0 * np.sin(w0 * tn))
              if west else 0.0)
        return g, dg

    # the boundary terms of the first stage come from boundary_kernel; every later stage's are
    # seeded into b by the epilogue before it, with the scalars of that (step, stage)
    b = np.zeros(nd)
    g, dg = scalars(0.0)
    emu.emu_boundary(b, st["v0"], bidx, bs, bd, ba, bidx.size, g, dg)
    t, dt_full = 0.0, dt
    while t < tf:                          # Linear.hpp:270-298, as fus_model_rk4 tabulates it
        dt = min(dt_full, tf - t)
        dt_next = min(dt_full, tf - (t + dt))
        for i in range(4):
            u_in = st["u0"] if i == 0 else st["un"]
            v_in = st["v0"] if i == 0 else st["vn"]
            emu.emu_stiffness(P + 1, 0, 0, u_in, _opt(v_in) if west else None, b, V.dofmap, _opt(G),
                              None, lin, _opt(att), nc, dphi, pts, wts, 2, 0, nc)
            gn, dgn = scalars(t + a_r[i + 1] * dt) if i < 3 else scalars(t + dt)
            rc = emu.emu_rk4_stage(i, int(west), b, mvec, _opt(dnl), st["u0"], st["v0"], st["ua"],
                                   st["va"], st["un"], st["vn"], nd, nd, dt, bidx.size, _opt(bidx),
                                   _opt(bs), _opt(bd), _opt(ba), _opt(bchunk), gn, dgn, 0, 2, None)
            assert rc == chunk
        t += dt
        del dt_next
    assert rel_l2(st["u0"], u_ref) < 1e-12 and rel_l2(st["v0"], v_ref) < 1e-12


class _FusedRank:
    """One rank of an emulated partitioned run on the fused peer transport: its partition, its
    mailbox (laid out by fus_halo_mailbox_layout, as halo_peer_export does), its transport state
    (tests/emu: emu_fused_*, the host mirror of halo_peer_connect) and its model vectors."""

    def __init__(self, emu, part):
        from fenicsx_fus_b200 import capi
        self.emu, self.p = emu, part
        self.neigh, self.soff, self.sidx, self.roff, self.ridx = part.halo_arrays()
        self.layout = np.zeros(6, dtype=np.int64)
        assert capi.load().fus_halo_mailbox_layout(self.sidx.size, self.ridx.size, len(self.neigh),
                                                   self.layout) == 0
        self.mbox = np.zeros(int(self.layout[5]) // 8 + 1, dtype=np.uint64)      # 8-byte aligned
        z = np.zeros(1, np.int32)
        self.h = emu.emu_fused_create(part.nowned, len(self.neigh), self.soff,
                                      self.sidx if self.sidx.size else z, self.roff,
                                      self.ridx if self.ridx.size else z,
                                      self.mbox.ctypes.data_as(_p), self.layout, 5.0)
        assert self.h, "the partitioner's numbering must have the shape the fused kernels need"
        assert emu.emu_fused_nshared(self.h) == part.nshared

    def connect(self, ranks):
        from fenicsx_fus_b200.partition import peer_byte_offsets
        for k, q in enumerate(self.neigh):
            o = ranks[int(q)]
            offs = np.array(peer_byte_offsets(o.neigh, o.soff, o.roff, o.layout, self.p.rank),
                            dtype=np.int64)
            assert all(0 <= b < o.layout[5] for b in offs)
            self.emu.emu_fused_connect(self.h, k, o.mbox.ctypes.data_as(_p), offs)

    def state(self):
        seq, ctr = np.zeros(8, dtype=np.uint64), np.zeros(8, dtype=np.uint32)
        err = self.emu.emu_fused_state(self.h, seq, ctr)
        return err, seq, ctr

    def close(self):
        self.emu.emu_fused_destroy(self.h)


def _reduce_to_owners(ranks, vecs):
    """Host stand-in for scatter_rev at set-up: the sum over all ranks that hold a dof."""
    total = {}
    for rk, v in zip(ranks, vecs):
        for k, val in zip(rk.p.global_key, v):
            total[int(k)] = total.get(int(k), 0.0) + val
    return [np.array([total[int(k)] for k in rk.p.global_key]) for rk in ranks]


@pytest.mark.parametrize("P,n,pg,geom,kind", [
    (2, (4, 3, 2), (2, 1, 1), 0, "linear"), (2, (8, 2, 2), (3, 1, 1), 0, "westervelt"),
    (1, (4, 4, 2), (2, 2, 1), 6, "linear"), (1, (2, 2, 2), (2, 2, 2), 4, "westervelt"),
    (3, (2, 6, 2), (1, 3, 1), 6, "westervelt"), (4, (6, 2, 1), (2, 1, 1), 0, "linear"),
    (3, (10, 5, 4), (2, 1, 1), 0, "linear")])
def test_emulated_fused_halo_rk4(fus, orc, emu, P, n, pg, geom, kind):
    """The partitioned RK4 flow on the fused peer transport, every rank of a process grid emulated:
    handshake and owner -> ghost update at entry (halo_ready_kernel, halo_entry_put_kernel), then per
    stage the HALO line kernel over the interface cells (ghost values gathered from the mailbox, ghost
    partial sums of b shipped to the owners after the last cell), the plain kernel over the rest, and
    the HALO epilogue (neighbours' partial sums added in send-list order, next stage input stored
    into the neighbours' mailboxes, flags raised by the block that finishes the last shared chunk,
    closing wait for the neighbours' forward data), and
    halo_exit_unpack_kernel -- with the native partitioner's lists, fus_halo_mailbox_layout and
    fus_halo_peer_offsets, i.e. what the multi-GPU runs connect over CUDA IPC.  Ranks advance in
    lockstep, kernel by kernel, so no wait may ever block (5 s time-out = failure).  Every rank's
    whole state, ghosts included, equals the single-domain oracle; the device-side sequence numbers
    end where the protocol says; slabs (a middle rank with interior cells and two neighbours), 2x2,
    2x2x2 and 1x3 grids; all three HALO pipelines; the fused two-vector gather (Westervelt)."""
    from fenicsx_fus_b200 import capi
    from fenicsx_fus_b200.partition import BoxPartition
    lib = capi.load()
    west = kind == "westervelt"
    h = 0.002
    hi = tuple(h * k for k in n)
    R = int(np.prod(pg))
    parts = [BoxPartition(P, n, pg, r, lo=(0, 0, 0), hi=hi) for r in range(R)]
    ranks = [_FusedRank(emu, p) for p in parts]
    for rk in ranks:
        rk.connect(ranks)
    pts, wts = orc.gll(P + 1)
    dphi = orc.dphi(P)
    Nd = (P + 1) ** 3
    ncg = int(np.prod(n))
    gid = np.arange(ncg)
    c0g, rho0g = 1500.0 + 60.0 * np.sin(gid), 1000.0 + 40.0 * np.cos(2.0 * gid)
    deltag = (2e-3 + 1e-3 * np.sin(3.0 * gid)) if west else None
    betag = (3.5 + 0.3 * np.cos(gid)) if west else None
    f0, p0, s0 = 0.5e6, 6.0e4, 1500.0
    # ---- single-domain oracle
    xg, xd = orc.box_mesh(n, (0, 0, 0), hi)
    dmg = orc.box_dofmap(P, n, 0)
    ndg = int(dmg.max()) + 1
    Gg, dJg = orc.geometry(P, xg, xd)
    fg = orc.box_facets(n)
    fng, fsg = orc.facet_data(P, xg, xd, fg)
    om = orc.model(kind, P, ndg, dmg, Gg, dJg, dphi, c0g, rho0g, deltag, betag, fg, fng, fsg, f0, p0, s0)
    rng = np.random.default_rng(11)
    u0g, v0g = 1e3 * rng.uniform(-1, 1, ndg), 1e9 * rng.uniform(-1, 1, ndg)
    dt = 0.2 * np.sqrt(3) * h / (1600.0 * P * P)
    # ---- per-rank set-up as fus_model_create does it (partial sums reduced to the owners once)
    S = []
    for rk in ranks:
        p = rk.p
        cg = p.cell_global
        d = dict(c0=c0g[cg].copy(), rho0=rho0g[cg].copy())
        d["G"], d["dJ"] = orc.geometry(P, p.x, p.xdofmap)
        src, dsrc, absb, bmass = (np.zeros(p.ndofs) for _ in range(4))
        dl = deltag[cg].copy() if west else None
        assert lib.fus_boundary_vectors(capi.KINDS[kind], P, p.ncells, p.ndofs, p.x, p.xdofmap,
                                        p.dofmap, p.facets.shape[0], p.facets, d["c0"], d["rho0"],
                                        _opt(dl), capi.optional(src), capi.optional(dsrc),
                                        capi.optional(absb), capi.optional(bmass)) == 0
        dmf, dJf = np.ascontiguousarray(p.dofmap.reshape(-1)), np.ascontiguousarray(d["dJ"].reshape(-1))
        mvec = np.zeros(p.ndofs)
        emu.emu_mass(np.ones(p.ndofs), mvec, dmf, dJf, 1.0 / d["rho0"] / d["c0"] ** 2, p.ncells * Nd, Nd)
        d["m"] = mvec + bmass
        d["dnl"] = np.zeros(p.ndofs)
        if west:
            emu.emu_mass(np.ones(p.ndofs), d["dnl"], dmf, dJf,
                         2.0 * betag[cg] / d["rho0"] ** 2 / d["c0"] ** 4, p.ncells * Nd, Nd)
        d.update(src=src, dsrc=dsrc, absb=absb, lin=-1.0 / d["rho0"],
                 att=(-dl / d["rho0"] / d["c0"] ** 2) if west else None)
        S.append(d)
    for key in ("m", "dnl", "src", "dsrc", "absb"):
        for d, red in zip(S, _reduce_to_owners(ranks, [d[key] for d in S])):
            d[key] = red
    chunk = emu.emu_rk4_stage(0, 0, np.zeros(8), np.ones(8), None, np.zeros(8), np.zeros(8),
                              np.zeros(8), np.zeros(8), np.zeros(8), np.zeros(8), 0, 8, 1.0, 0, None,
                              None, None, None, None, 0.0, 0.0, 0, 1, None)
    for rk, d in zip(ranks, S):
        p = rk.p
        own = np.arange(p.nowned)
        sel = own[(d["src"][own] != 0) | (d["dsrc"][own] != 0) | (d["absb"][own] != 0)].astype(np.int32)
        d["bidx"], d["bs"], d["bd"], d["ba"] = sel, d["src"][sel].copy(), d["dsrc"][sel].copy(), d["absb"][sel].copy()
        d["bchunk"] = np.searchsorted(sel, np.arange(0, (p.ndofs + chunk - 1) // chunk + 1) * chunk).astype(np.int64)
        d["st"] = dict(u0=u0g[p.global_key].copy(), v0=v0g[p.global_key].copy(), ua=np.zeros(p.ndofs),
                       va=np.zeros(p.ndofs), un=np.zeros(p.ndofs), vn=np.zeros(p.ndofs))
        d["st"]["u0"][p.nowned:] = -5.0          # stale ghosts: the entry exchange must not need them
        d["st"]["v0"][p.nowned:] = 7.0
        d["b"] = np.zeros(p.ndofs)
    w0, kappa = 2 * np.pi * f0, (2.0 if west else 1.0)

    def scalars(tn):
        win = 0.5 * (1 - np.cos(f0 * np.pi * tn / 4.0)) if tn < 4.0 / f0 else 1.0
        dwin = 0.5 * np.pi * f0 / 4.0 * np.sin(f0 * np.pi * tn / 4.0) if tn < 4.0 / f0 else 0.0
        g = kappa * win * p0 * w0 / s0 * np.cos(w0 * tn)
        dg = (kappa * (dwin * p0 * w0 / s0 * np.cos(w0 * tn) - win * p0 * w0 * w0 / s0 * np.sin(w0 * tn))
              if west else 0.0)
        return g, dg

    a_r = (0.0, 0.5, 0.5, 1.0)
    # Epilogue grid.  The last case has ~3 000 dofs per rank = 3-4 chunks on ONE block, so that the
    # block takes chunk after chunk from the counter (prefetched claims, alternating shared words);
    # the others have fewer chunks than blocks.
    stage_grid = 1 if n == (10, 5, 4) else 3
    for call in range(2):                      # two rk4 calls: the handshake and the exit/entry pairing
        for rk in ranks:
            emu.emu_fused_ready(rk.h, 1)
        for rk in ranks:
            emu.emu_fused_ready(rk.h, 2)
        for rk, d in zip(ranks, S):
            emu.emu_fused_entry_put(rk.h, d["st"]["u0"], d["st"]["v0"])
        for rk in ranks:                       # the put's closing wait (in-kernel on the device)
            emu.emu_fused_forward_landed(rk.h)
        t = call * dt
        g, dg = scalars(t)
        for rk, d in zip(ranks, S):
            d["b"][:] = 0.0
            emu.emu_boundary(d["b"], d["st"]["v0"], d["bidx"], d["bs"], d["bd"], d["ba"],
                             d["bidx"].size, g, dg)
        for i in range(4):
            for rk, d in zip(ranks, S):
                p, st = rk.p, d["st"]
                u_in = st["u0"] if i == 0 else st["un"]
                v_in = st["v0"] if i == 0 else st["vn"]
                assert emu.emu_stiffness_fused(P + 1, geom, u_in, _opt(v_in) if west else None, d["b"],
                                               p.dofmap, d["G"], d["lin"], _opt(d["att"]), p.ncells,
                                               dphi, pts, wts, 2, rk.h, p.ninterface_cells) == 0
            gn, dgn = scalars(t + a_r[i + 1] * dt) if i < 3 else scalars(t + dt)
            for rk, d in zip(ranks, S):
                p, st = rk.p, d["st"]
                rc = emu.emu_rk4_stage(i, int(west), d["b"], d["m"], _opt(d["dnl"]), st["u0"], st["v0"],
                                       st["ua"], st["va"], st["un"], st["vn"], p.nowned, p.ndofs, dt,
                                       d["bidx"].size, _opt(d["bidx"]), _opt(d["bs"]), _opt(d["bd"]),
                                       _opt(d["ba"]), _opt(d["bchunk"]), gn, dgn, 0, stage_grid, rk.h)
                assert rc == chunk
            for rk in ranks:                   # the epilogue's closing wait (in-kernel on the device)
                emu.emu_fused_forward_landed(rk.h)
        for rk, d in zip(ranks, S):
            emu.emu_fused_exit(rk.h, d["st"]["u0"], d["st"]["v0"])
        for rk in ranks:
            err, seq, ctr = rk.state()
            assert err == 0
            # entry + 4 epilogues per call sent forward; 4 operators per call sent reverse
            assert (int(seq[0]), int(seq[1]), int(seq[2]), int(seq[3]), int(seq[4])) == (
                5 * (call + 1), 5 * (call + 1), 4 * (call + 1), 4 * (call + 1), call + 1)
    u_ref2, v_ref2 = u0g.copy(), v0g.copy()
    # two calls of one step each == the reference loop over two (full) steps
    assert om.rk4(0.0, 2.0 * dt, dt, u_ref2, v_ref2) == 2
    for rk, d in zip(ranks, S):
        key = rk.p.global_key
        assert rel_l2(d["st"]["u0"], u_ref2[key]) < 1e-12, (rk.p.rank, "u")      # ghosts included
        assert rel_l2(d["st"]["v0"], v_ref2[key]) < 1e-12, (rk.p.rank, "v")
        rk.close()


def test_fused_halo_refuses_other_numberings(fus, emu):
    """The fused kernels rely on (P1) shared owned dofs numbered first and (P2) ghosts numbered
    neighbour by neighbour: lists that break either are refused (the context then stays on NCCL)."""
    from fenicsx_fus_b200.partition import BoxPartition
    p = BoxPartition(2, (4, 2, 2), (2, 1, 1), 0)
    neigh, soff, sidx, roff, ridx = p.halo_arrays()
    assert sidx.size and np.array_equal(np.unique(sidx), np.arange(p.nshared))
    mbox = np.zeros(1 << 16, dtype=np.uint64)
    lay = np.zeros(6, dtype=np.int64)
    z = np.zeros(1, np.int32)
    ok = emu.emu_fused_create(p.nowned, len(neigh), soff, sidx, roff, z, mbox.ctypes.data_as(_p), lay,
                              1.0)
    assert ok
    emu.emu_fused_destroy(ok)
    bad = sidx.copy()
    bad[0] = p.nowned - 1                      # an owned dof that is not among the first nshared
    assert not emu.emu_fused_create(p.nowned, len(neigh), soff, bad, roff, z, mbox.ctypes.data_as(_p),
                                    lay, 1.0)
    p1 = BoxPartition(2, (4, 2, 2), (2, 1, 1), 1)
    neigh, soff, sidx, roff, ridx = p1.halo_arrays()
    assert np.array_equal(ridx, p1.nowned + np.arange(ridx.size))
    assert not emu.emu_fused_create(p1.nowned, len(neigh), soff, z, roff, ridx[::-1].copy(),
                                    mbox.ctypes.data_as(_p), lay, 1.0)


def test_emulated_nccl_pack_unpack(fus, emu):
    """halo_pack_kernel / halo_unpack_kernel (NCCL transport): pack on the sender, a host copy in
    place of ncclSend/ncclRecv, unpack (insert forward, add reverse) on the receiver."""
    from fenicsx_fus_b200.partition import BoxPartition
    P, n, pg = 2, (4, 3, 3), (2, 1, 1)
    parts = [BoxPartition(P, n, pg, r) for r in range(2)]
    arr = [p.halo_arrays() for p in parts]
    u = [np.where(np.arange(p.ndofs) < p.nowned, np.cos(p.global_key.astype(float)), -5.0) for p in parts]
    v = [np.where(np.arange(p.ndofs) < p.nowned, p.global_key.astype(float), -6.0) for p in parts]
    bufs = []
    for r in range(2):
        neigh, soff, sidx, roff, ridx = arr[r]
        buf = np.zeros(2 * max(1, sidx.size))
        if sidx.size:
            emu.emu_halo_pack(u[r], _opt(v[r]), sidx, soff, len(neigh), buf, sidx.size, 2)
        bufs.append(buf)
    for r in range(2):
        neigh, soff, sidx, roff, ridx = arr[r]
        if ridx.size:                                             # two ranks: one neighbour each
            emu.emu_halo_unpack(0, u[r], _opt(v[r]), ridx, roff, len(neigh), bufs[1 - r], ridx.size, 2)
        key = parts[r].global_key.astype(float)
        assert np.array_equal(u[r], np.cos(key)) and np.array_equal(v[r], key)
