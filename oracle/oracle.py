"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  See the header of oracle/fus_oracle.c for scope and parity status.

`Oracle()`            -> plain-C restatement (libfus_oracle.so, strict IEEE)
`Oracle(ref=True)`    -> oracle/_ref: the same entry points compiled with the reference's flags,
                         plus `fr_*` cell loops built on the reference's own
                         sum_factorisation.hpp (threaded; the timed CPU baseline)
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i64 = C.c_int64
_int = C.c_int
_dbl = C.c_double


def _opt(arr):
    """Pointer or NULL for optional float64 arrays."""
    if arr is None:
        return None
    assert arr.dtype == np.float64 and arr.flags.c_contiguous
    return arr.ctypes.data_as(C.c_void_p)


def build(ref=True):
    """(Re)build the oracle libraries with oracle/Makefile.  `make ref` needs /root/reference;
    when it is absent (GPU box) the prebuilt oracle/_ref/*.so that travelled are used."""
    subprocess.run(["make", "-s", "-C", _HERE], check=True)
    if ref and os.path.isdir("/root/reference/cpp/fenicsx-sf/common"):
        subprocess.run(["make", "-s", "-C", _HERE, "ref"], check=True)


def _cpu_has_avx512():
    try:
        with open("/proc/cpuinfo") as f:
            txt = f.read()
        return " avx512f" in txt and " avx512dq" in txt and " avx512vl" in txt and " avx512bw" in txt
    except OSError:
        return False


def ref_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libfus_ref_v3.so"))


class Oracle:
    def __init__(self, ref=False):
        self.is_ref = ref
        if ref:
            name = "libfus_ref_v4.so" if _cpu_has_avx512() else "libfus_ref_v3.so"
            path = os.path.join(_HERE, "_ref", name)
            if not os.path.exists(path):
                build(ref=True)
        else:
            path = os.path.join(_HERE, "libfus_oracle.so")
            src = os.path.join(_HERE, "fus_oracle.c")
            if not os.path.exists(path) or os.path.getmtime(path) < os.path.getmtime(src):
                build(ref=False)
        self.path = path
        L = self.lib = C.CDLL(path)
        L.fo_gll.argtypes = [_int, _f64p, _f64p]
        L.fo_gll.restype = _int
        L.fo_dphi.argtypes = [_int, _f64p, _f64p]
        L.fo_contract.argtypes = [_int, _int, _int, _int, _int, _f64p, _f64p, _f64p]
        L.fo_transpose.argtypes = [_int, _int, _int, _int, _int, _int, _f64p, _f64p]
        L.fo_box_mesh.argtypes = [_int, _int, _int, _f64p, _f64p, _f64p, _i32p]
        L.fo_box_dofmap.argtypes = [_int, _int, _int, _int, _int, _i32p]
        L.fo_geometry.argtypes = [_i64, _f64p, _i32p, _int, _f64p, _f64p, C.c_void_p, C.c_void_p]
        L.fo_mass_apply.argtypes = [_int, _i64, _i32p, _f64p, _f64p, _f64p, _f64p]
        L.fo_stiffness_apply.argtypes = [_int, _i64, _i32p, _f64p, _f64p, _f64p, _f64p, _f64p]
        L.fo_facet_data.argtypes = [_int, _f64p, _i32p, _f64p, _f64p, _i64, _int, _i32p, _f64p]
        L.fo_box_facets.argtypes = [_int, _int, _int, C.c_void_p]
        L.fo_box_facets.restype = _i64
        L.fo_model_create.argtypes = [_int, _int, _i64, _i64, _i64, _i32p, _f64p, _f64p, _f64p,
                                      _f64p, _f64p, _f64p, _f64p, _i64, _i32p, _i32p, _f64p,
                                      _dbl, _dbl, _dbl]
        L.fo_model_create.restype = C.c_void_p
        L.fo_model_destroy.argtypes = [C.c_void_p]
        L.fo_model_mass.argtypes = [C.c_void_p]
        L.fo_model_mass.restype = C.POINTER(C.c_double)
        L.fo_model_set_ops.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.fo_model_f1.argtypes = [C.c_void_p, _dbl, _f64p, _f64p, _f64p]
        L.fo_model_rk4.argtypes = [C.c_void_p, _dbl, _dbl, _dbl, _f64p, _f64p]
        L.fo_model_rk4.restype = _int
        # 2-D quadrilateral variant (cpp/fenicsx-sf-naive)
        L.fo_rect_mesh.argtypes = [_int, _int, _f64p, _f64p, _f64p, _i32p]
        L.fo_rect_dofmap.argtypes = [_int, _int, _int, _i32p]
        L.fo_geometry_2d.argtypes = L.fo_geometry.argtypes
        L.fo_mass_apply_2d.argtypes = L.fo_mass_apply.argtypes
        L.fo_stiffness_apply_2d.argtypes = L.fo_stiffness_apply.argtypes
        L.fo_facet_data_2d.argtypes = L.fo_facet_data.argtypes
        L.fo_rect_facets.argtypes = [_int, _int, C.c_void_p]
        L.fo_rect_facets.restype = _i64
        L.fo_model_create_2d.argtypes = L.fo_model_create.argtypes
        L.fo_model_create_2d.restype = C.c_void_p
        L.fo_contract_2d.argtypes = [_int, _int, _int, _f64p, _f64p, _f64p]
        L.fo_transpose_2d.argtypes = [_int, _int, _int, _int, _f64p, _f64p]
        if ref:
            if hasattr(L, "fr_stiffness_apply_2d"):      # absent from a prebuilt older _ref
                L.fr_stiffness_apply_2d.argtypes = L.fo_stiffness_apply.argtypes
            L.fr_set_threads.argtypes = [_int]
            L.fr_max_threads.restype = _int
            L.fr_stiffness_apply.argtypes = L.fo_stiffness_apply.argtypes
            L.fr_mass_apply.argtypes = L.fo_mass_apply.argtypes
            L.fr_kat.argtypes = [_f64p, _f64p]
            L.fr_contract_cube.argtypes = [_int, _int, _f64p, _f64p, _f64p]
            L.fr_transpose_cube.argtypes = [_int, _int, _f64p, _f64p]

    # ---- 1-D tables -----------------------------------------------------------------
    def gll(self, m):
        pts, wts = np.zeros(m), np.zeros(m)
        assert self.lib.fo_gll(m, pts, wts) == 0
        return pts, wts

    def dphi(self, P):
        N = P + 1
        pts, _ = self.gll(N)
        d = np.zeros(N * N)
        self.lib.fo_dphi(N, pts, d)
        return d

    # ---- mesh -------------------------------------------------------------------------
    def box_mesh(self, n, lo=(0.0, 0.0, 0.0), hi=(1.0, 1.0, 1.0)):
        nx, ny, nz = n
        xg = np.zeros(((nx + 1) * (ny + 1) * (nz + 1), 3))
        xd = np.zeros((nx * ny * nz, 8), dtype=np.int32)
        self.lib.fo_box_mesh(nx, ny, nz, np.array(lo, dtype=np.float64),
                             np.array(hi, dtype=np.float64), xg, xd)
        return xg, xd

    def box_dofmap(self, P, n, mode=0):
        nx, ny, nz = n
        dm = np.zeros((nx * ny * nz, (P + 1) ** 3), dtype=np.int32)
        self.lib.fo_box_dofmap(P, nx, ny, nz, mode, dm)
        return dm

    def box_facets(self, n):
        nx, ny, nz = n
        k = self.lib.fo_box_facets(nx, ny, nz, None)
        f = np.zeros((k, 3), dtype=np.int32)
        self.lib.fo_box_facets(nx, ny, nz, f.ctypes.data_as(C.c_void_p))
        return f

    def geometry(self, P, xg, xd, want_G=True, want_detJ=True):
        N = P + 1
        pts, wts = self.gll(N)
        nc = xd.shape[0]
        G = np.zeros((nc, N ** 3, 6)) if want_G else None
        dJ = np.zeros((nc, N ** 3)) if want_detJ else None
        self.lib.fo_geometry(nc, xg, xd, N, pts, wts, _opt(G), _opt(dJ))
        return G, dJ

    def facet_data(self, P, xg, xd, facets):
        N = P + 1
        pts, wts = self.gll(N)
        nf = facets.shape[0]
        fn = np.zeros((nf, N * N), dtype=np.int32)
        fs = np.zeros((nf, N * N))
        for k in range(nf):
            self.lib.fo_facet_data(N, xg, xd, pts, wts, int(facets[k, 0]), int(facets[k, 1]),
                                   fn[k], fs[k])
        return fn, fs

    # ---- operators ----------------------------------------------------------------------
    def stiffness_apply(self, P, dofmap, G, dphi, coeffs, x, y, use_ref_kernels=False):
        fn = self.lib.fr_stiffness_apply if use_ref_kernels else self.lib.fo_stiffness_apply
        fn(P, dofmap.shape[0], dofmap, G, dphi, coeffs, x, y)
        return y

    def mass_apply(self, P, dofmap, detJ, coeffs, x, y, use_ref_kernels=False):
        fn = self.lib.fr_mass_apply if use_ref_kernels else self.lib.fo_mass_apply
        fn(P, dofmap.shape[0], dofmap, detJ, coeffs, x, y)
        return y

    # ---- 2-D quadrilateral variant ------------------------------------------------------
    def rect_mesh(self, n, lo=(0.0, 0.0), hi=(1.0, 1.0)):
        nx, ny = n
        xg = np.zeros(((nx + 1) * (ny + 1), 3))
        xd = np.zeros((nx * ny, 4), dtype=np.int32)
        self.lib.fo_rect_mesh(nx, ny, np.array(lo, dtype=np.float64),
                              np.array(hi, dtype=np.float64), xg, xd)
        return xg, xd

    def rect_dofmap(self, P, n):
        dm = np.zeros((n[0] * n[1], (P + 1) ** 2), dtype=np.int32)
        self.lib.fo_rect_dofmap(P, n[0], n[1], dm)
        return dm

    def rect_facets(self, n):
        k = self.lib.fo_rect_facets(n[0], n[1], None)
        f = np.zeros((k, 3), dtype=np.int32)
        self.lib.fo_rect_facets(n[0], n[1], f.ctypes.data_as(C.c_void_p))
        return f

    def geometry_2d(self, P, xg, xd):
        N = P + 1
        pts, wts = self.gll(N)
        nc = xd.shape[0]
        G, dJ = np.zeros((nc, N * N, 3)), np.zeros((nc, N * N))
        self.lib.fo_geometry_2d(nc, xg, xd, N, pts, wts, _opt(G), _opt(dJ))
        return G, dJ

    def facet_data_2d(self, P, xg, xd, facets):
        N = P + 1
        pts, wts = self.gll(N)
        nf = facets.shape[0]
        fn, fs = np.zeros((nf, N), dtype=np.int32), np.zeros((nf, N))
        for k in range(nf):
            self.lib.fo_facet_data_2d(N, xg, xd, pts, wts, int(facets[k, 0]), int(facets[k, 1]),
                                      fn[k], fs[k])
        return fn, fs

    def stiffness_apply_2d(self, P, dofmap, G, dphi, coeffs, x, y, use_ref_kernels=False):
        fn = self.lib.fr_stiffness_apply_2d if use_ref_kernels else self.lib.fo_stiffness_apply_2d
        fn(P, dofmap.shape[0], dofmap, G, dphi, coeffs, x, y)
        return y

    def mass_apply_2d(self, P, dofmap, detJ, coeffs, x, y):
        self.lib.fo_mass_apply_2d(P, dofmap.shape[0], dofmap, detJ, coeffs, x, y)
        return y

    def model_2d(self, kind, P, ndofs, dofmap, G, detJ, dphi, c0, rho0, delta0, beta0, facets,
                 fnodes, fscale, freq, p0, s0):
        return OracleModel(self, kind, P, ndofs, dofmap, G, detJ, dphi, c0, rho0, delta0, beta0,
                           facets, fnodes, fscale, freq, p0, s0, None, False, dim=2)

    # ---- models -------------------------------------------------------------------------
    def model(self, kind, P, ndofs, dofmap, G, detJ, dphi, c0, rho0, delta0, beta0, facets,
              fnodes, fscale, freq, p0, s0, nowned=None, use_ref_kernels=False):
        return OracleModel(self, kind, P, ndofs, dofmap, G, detJ, dphi, c0, rho0, delta0, beta0,
                           facets, fnodes, fscale, freq, p0, s0, nowned, use_ref_kernels)


class OracleModel:
    KINDS = {"linear": 0, "lossy": 1, "westervelt": 2}

    def __init__(self, orc, kind, P, ndofs, dofmap, G, detJ, dphi, c0, rho0, delta0, beta0,
                 facets, fnodes, fscale, freq, p0, s0, nowned, use_ref_kernels, dim=3):
        self.orc = orc
        self.ndofs = ndofs
        kind = self.KINDS.get(kind, kind)
        nc = dofmap.shape[0]
        z = np.zeros(nc)
        delta0 = z if delta0 is None else delta0
        beta0 = z if beta0 is None else beta0
        # keep every borrowed array alive for the lifetime of the C object
        self._keep = [np.ascontiguousarray(a) for a in
                      (dofmap, G, detJ, dphi, c0, rho0, delta0, beta0, facets, fnodes, fscale)]
        (dofmap, G, detJ, dphi, c0, rho0, delta0, beta0, facets, fnodes, fscale) = self._keep
        create = orc.lib.fo_model_create if dim == 3 else orc.lib.fo_model_create_2d
        self.h = create(kind, P, nc, ndofs, ndofs if nowned is None else nowned, dofmap, G, detJ,
                        dphi, c0, rho0, delta0, beta0, facets.shape[0], facets, fnodes, fscale,
                        freq, p0, s0)
        if use_ref_kernels:
            L = orc.lib
            orc.lib.fo_model_set_ops(self.h, C.cast(L.fr_stiffness_apply, C.c_void_p),
                                     C.cast(L.fr_mass_apply, C.c_void_p))

    def mass(self):
        p = self.orc.lib.fo_model_mass(self.h)
        return np.ctypeslib.as_array(p, shape=(self.ndofs,)).copy()

    def f1(self, t, u, v):
        out = np.zeros(self.ndofs)
        self.orc.lib.fo_model_f1(self.h, t, u, v, out)
        return out

    def rk4(self, t0, tf, dt, u, v):
        """Advances u, v in place; returns the number of steps taken."""
        return self.orc.lib.fo_model_rk4(self.h, t0, tf, dt, u, v)

    def __del__(self):
        try:
            self.orc.lib.fo_model_destroy(self.h)
        except Exception:
            pass
