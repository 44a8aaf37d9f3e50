"""fenicsx-fus_b200: B200-native sum-factorised acoustic operator + RK4 (host-side Python mirror).

Thin Python mirror of the reference's operator/solver interface for the hot path
(cpp/fenicsx-sf/common/{spectral_op,Linear,Lossy,Westervelt}.hpp), on top of the C ABI in
include/fus_b200.h.  All arithmetic runs in the CUDA library; nothing here computes on the CPU
besides mesh/dofmap set-up, and nothing falls back when the library or a GPU is missing.

The C++ mirror with the reference's exact template signatures lives in include/fus/*.hpp.
"""
import ctypes as C

import numpy as np

from . import capi
from .capi import KINDS, FusError, check  # noqa: F401

__all__ = ["BoxMesh", "FunctionSpace", "HexMesh", "HexFunctionSpace", "StiffnessSpectral3D", "MassSpectral3D",
           "LinearSpectral3D", "LossySpectral3D", "WesterveltSpectral3D", "gll",
           "tabulate_dphi", "compute_diffusivity_of_sound", "launch_count", "device_count"]


def __getattr__(name):
    # unstructured-mesh ingestion lives in its own module (needs numpy only); import lazily
    if name in ("HexMesh", "HexFunctionSpace"):
        from . import unstructured
        return getattr(unstructured, name)
    raise AttributeError(name)


def gll(P):
    """GLL points/weights on [0,1], Basix order (spectral_op.hpp:57-59)."""
    pts, wts = np.zeros(P + 1), np.zeros(P + 1)
    check(capi.load().fus_gll(P, pts, wts), "fus_gll")
    return pts, wts


def tabulate_dphi(P):
    """dphi[q*N+i] = phi_i'(xi_q) (precompute.hpp:217-234, spectral_op.hpp:168-170)."""
    d = np.zeros((P + 1) * (P + 1))
    check(capi.load().fus_tabulate_dphi(P, d), "fus_tabulate_dphi")
    return d


def compute_diffusivity_of_sound(w0, c0, alpha):
    """Westervelt.hpp:408-413."""
    return 2 * alpha * c0 * c0 * c0 / w0 / w0


def launch_count():
    return int(capi.load().fus_launch_count())


def device_count():
    return int(capi.load().fus_device_count())


class BoxMesh:
    """Structured hexahedral box (dolfinx::mesh::create_box stand-in).

    x: (nverts,3) vertex coordinates, xdofmap: (ncells,8) DOLFINx tensor vertex order,
    facets: (nfacets,3) exterior facets {cell, local facet, tag} with tag 1 on x=lo, 2 on x=hi.
    `warp(x) -> x'` optionally displaces the vertices (non-affine trilinear cells).
    """

    def __init__(self, n, lo=(0.0, 0.0, 0.0), hi=(1.0, 1.0, 1.0), warp=None):
        lib = capi.load()
        self.n = np.asarray(n, dtype=np.int32)
        assert self.n.shape == (3,)
        nx, ny, nz = (int(v) for v in self.n)
        self.lo = np.asarray(lo, dtype=np.float64)
        self.hi = np.asarray(hi, dtype=np.float64)
        self.x = np.zeros(((nx + 1) * (ny + 1) * (nz + 1), 3))
        self.xdofmap = np.zeros((nx * ny * nz, 8), dtype=np.int32)
        check(lib.fus_box_mesh(self.n, self.lo, self.hi, self.x, self.xdofmap), "fus_box_mesh")
        if warp is not None:
            self.x = np.ascontiguousarray(warp(self.x), dtype=np.float64)
        nf = lib.fus_box_facets(self.n, None)
        self.facets = np.zeros((nf, 3), dtype=np.int32)
        lib.fus_box_facets(self.n, self.facets.ctypes.data_as(C.c_void_p))
        self.ncells = nx * ny * nz

    def tag_source_disc(self, centre_yz, radius):
        """Re-tag the x = lo face: only facets whose centroid lies within `radius` of
        (y, z) = centre_yz keep tag 1 (source), the rest of that face becomes 0.  Stands in for the
        transducer surface of the reference's bowl meshes, which are not available
        (SURVEY.md section 8d config 4)."""
        corners = {2: (0, 2, 4, 6)}                       # local facet 2 is x = 0
        f = self.facets
        on = f[:, 1] == 2
        cen = self.x[self.xdofmap[f[on, 0]][:, corners[2]]].mean(axis=1)
        d = np.hypot(cen[:, 1] - centre_yz[0], cen[:, 2] - centre_yz[1])
        tags = f[:, 2].copy()
        tags[np.flatnonzero(on)] = np.where(d <= radius, 1, 0)
        self.facets = np.ascontiguousarray(np.stack([f[:, 0], f[:, 1], tags], axis=1), dtype=np.int32)
        return int((tags == 1).sum())

    def h_min(self):
        """Cell diameter as dolfinx::mesh::h reports it for an undeformed box (sqrt(3) h)."""
        h = (self.hi - self.lo) / self.n
        return float(np.sqrt((h * h).sum()))


class FunctionSpace:
    """Degree-P GLL Lagrange space on a BoxMesh with the tensor-product dofmap
    (create_functionspace + reorder_dofmap, permute.hpp:15-42)."""

    def __init__(self, mesh, P, numbering=1):
        lib = capi.load()
        self.mesh, self.P = mesh, int(P)
        self.N = self.P + 1
        self.ndofs = int(lib.fus_box_num_dofs(self.P, mesh.n))
        self.nowned = self.ndofs
        self.dofmap = np.zeros((mesh.ncells, self.N ** 3), dtype=np.int32)
        check(lib.fus_box_dofmap(self.P, mesh.n, numbering, self.dofmap), "fus_box_dofmap")
        self._ctx = None

    def context(self, device=0, lean=False):
        """The device context of this space (created on first use).  lean=True keeps neither G
        nor detJ on the device: geometric factors are rebuilt from the trilinear cell map."""
        if self._ctx is None:
            self._ctx = Context.from_mesh(self, device, lean=lean)
        return self._ctx

    def tabulate_dof_coordinates(self):
        """Physical coordinates of every dof (trilinear map of the GLL nodes)."""
        pts, _ = gll(self.P)
        N, m = self.N, self.mesh
        X = m.x[m.xdofmap]                      # (nc, 8, 3)
        xi = np.stack(np.meshgrid(pts, pts, pts, indexing="ij"), -1).reshape(-1, 3)  # (Nd,3)
        out = np.zeros((self.ndofs, 3))
        for v in range(8):
            a, b, c = v & 1, (v >> 1) & 1, (v >> 2) & 1
            w = ((xi[:, 0] if a else 1 - xi[:, 0]) * (xi[:, 1] if b else 1 - xi[:, 1])
                 * (xi[:, 2] if c else 1 - xi[:, 2]))            # (Nd,)
            contrib = w[None, :, None] * X[:, v, None, :]         # (nc, Nd, 3)
            if v == 0:
                acc = contrib
            else:
                acc = acc + contrib
        out[self.dofmap.reshape(-1)] = acc.reshape(-1, 3)
        return out


class Context:
    """Owner of the device cell data (fus_ctx)."""

    def __init__(self, handle, P, ncells, ndofs, nowned, device):
        self.h, self.P, self.ncells, self.ndofs = handle, P, ncells, ndofs
        self.nowned, self.device = nowned, device
        self.lib = capi.load()

    @classmethod
    def from_mesh(cls, V, device=0, dofmap=None, ndofs=None, nowned=None, lean=False):
        lib = capi.load()
        m = V.mesh
        dm = V.dofmap if dofmap is None else dofmap
        ndofs = V.ndofs if ndofs is None else ndofs
        nowned = ndofs if nowned is None else nowned
        h = C.c_void_p()
        create = lib.fus_ctx_create_from_mesh_lean if lean else lib.fus_ctx_create_from_mesh
        check(create(V.P, dm.shape[0], ndofs, nowned, np.ascontiguousarray(dm), m.x.shape[0], m.x,
                     m.xdofmap, device, C.byref(h)), "fus_ctx_create_from_mesh")
        return cls(h, V.P, dm.shape[0], ndofs, nowned, device)

    @classmethod
    def from_arrays(cls, P, dofmap, ndofs, G, detJ, dphi, device=0, nowned=None):
        lib = capi.load()
        h = C.c_void_p()
        nowned = ndofs if nowned is None else nowned
        check(lib.fus_ctx_create(P, dofmap.shape[0], ndofs, nowned, np.ascontiguousarray(dofmap),
                                 capi.optional(G), capi.optional(detJ), dphi, device,
                                 C.byref(h)), "fus_ctx_create")
        return cls(h, P, dofmap.shape[0], ndofs, nowned, device)

    def set_stream(self, cuda_stream):
        check(self.lib.fus_ctx_set_stream(self.h, C.c_void_p(cuda_stream)), "fus_ctx_set_stream")

    def set_option(self, name, value):
        check(self.lib.fus_ctx_set_option(self.h, name.encode(), int(value)), "fus_ctx_set_option")

    def get_option(self, name):
        v = C.c_int(0)
        check(self.lib.fus_ctx_get_option(self.h, name.encode(), C.byref(v)), "fus_ctx_get_option")
        return v.value

    def sync(self):
        check(self.lib.fus_ctx_sync(self.h), "fus_ctx_sync")

    def profile(self, kernel):
        """(launches, total device ms) of one kernel family since option profile_kernels=1."""
        n, ms = C.c_int64(0), C.c_double(0.0)
        check(self.lib.fus_ctx_profile(self.h, kernel.encode(), C.byref(n), C.byref(ms)),
              "fus_ctx_profile")
        return n.value, ms.value

    def geometry(self, want_G=True, want_detJ=True):
        Nd = (self.P + 1) ** 3
        G = np.zeros((self.ncells, Nd, 6)) if want_G else None
        dJ = np.zeros((self.ncells, Nd)) if want_detJ else None
        check(self.lib.fus_ctx_get_geometry(self.h, capi.optional(G), capi.optional(dJ)),
              "fus_ctx_get_geometry")
        return G, dJ

    # -- device buffers for callers without torch --
    def alloc(self, nbytes):
        p = C.c_void_p()
        check(self.lib.fus_dev_alloc(self.h, nbytes, C.byref(p)), "fus_dev_alloc")
        return p

    def free(self, p):
        check(self.lib.fus_dev_free(self.h, p), "fus_dev_free")

    def upload(self, dptr, arr):
        arr = np.ascontiguousarray(arr)
        check(self.lib.fus_dev_upload(self.h, dptr, arr.ctypes.data_as(C.c_void_p), arr.nbytes),
              "fus_dev_upload")

    def download(self, arr, dptr):
        check(self.lib.fus_dev_download(self.h, arr.ctypes.data_as(C.c_void_p), dptr, arr.nbytes),
              "fus_dev_download")

    def destroy(self):
        if self.h:
            self.lib.fus_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


def _is_device_tensor(a):
    return hasattr(a, "data_ptr") and hasattr(a, "is_cuda") and a.is_cuda


class _Operator:
    _dev_fn = _host_fn = None

    def __init__(self, V, device=0):
        self.V = V
        self.ctx = V if isinstance(V, Context) else V.context(device)

    def __call__(self, x, coeffs, y):
        """y += A(coeffs) x.  numpy arrays go through the host entry point (copies included);
        CUDA tensors (float64, contiguous) through the device entry point on the context stream."""
        lib = self.ctx.lib
        if _is_device_tensor(x):
            check(getattr(lib, self._dev_fn)(self.ctx.h, C.c_void_p(x.data_ptr()),
                                             C.c_void_p(coeffs.data_ptr()),
                                             C.c_void_p(y.data_ptr())), self._dev_fn)
        else:
            check(getattr(lib, self._host_fn)(self.ctx.h, x, np.ascontiguousarray(coeffs), y),
                  self._host_fn)
        return y


class StiffnessSpectral3D(_Operator):
    """StiffnessSpectral3D<T,P>::operator() (spectral_op.hpp:132-243)."""
    _dev_fn, _host_fn = "fus_stiffness_apply_dev", "fus_stiffness_apply_host"


class MassSpectral3D(_Operator):
    """MassSpectral3D<T,P>::operator() (spectral_op.hpp:29-107)."""
    _dev_fn, _host_fn = "fus_mass_apply_dev", "fus_mass_apply_host"


class _Model:
    """Common part of {Linear,Lossy,Westervelt}Spectral3D (Linear.hpp:52-347)."""
    kind = None

    def __init__(self, V, c0, rho0, delta0, beta0, sourceFrequency, sourceAmplitude, sourceSpeed,
                 facets=None, device=0):
        lib = capi.load()
        self.V, self.ctx = V, V.context(device)
        m = V.mesh
        nc, nd = m.ncells, V.ndofs
        c0 = np.ascontiguousarray(np.broadcast_to(np.asarray(c0, dtype=np.float64), (nc,)))
        rho0 = np.ascontiguousarray(np.broadcast_to(np.asarray(rho0, dtype=np.float64), (nc,)))
        if delta0 is not None:
            delta0 = np.ascontiguousarray(np.broadcast_to(np.asarray(delta0, np.float64), (nc,)))
        if beta0 is not None:
            beta0 = np.ascontiguousarray(np.broadcast_to(np.asarray(beta0, np.float64), (nc,)))
        facets = m.facets if facets is None else np.ascontiguousarray(facets, dtype=np.int32)
        src, dsrc, absb, bmass = (np.zeros(nd) for _ in range(4))
        k = KINDS[self.kind]
        check(lib.fus_boundary_vectors(k, V.P, nc, nd, m.x, m.xdofmap, V.dofmap,
                                       facets.shape[0], facets, c0, rho0, capi.optional(delta0),
                                       capi.optional(src), capi.optional(dsrc),
                                       capi.optional(absb), capi.optional(bmass)),
              "fus_boundary_vectors")
        self.boundary = dict(src=src, dsrc=dsrc, absb=absb, bmass=bmass)
        h = C.c_void_p()
        check(lib.fus_model_create(self.ctx.h, k, c0, rho0, capi.optional(delta0),
                                   capi.optional(beta0), capi.optional(src), capi.optional(dsrc),
                                   capi.optional(absb), capi.optional(bmass),
                                   float(sourceFrequency), float(sourceAmplitude),
                                   float(sourceSpeed), C.byref(h)), "fus_model_create")
        self.h = h
        self.lib = lib

    def init(self, u=None, v=None):
        """init() (Linear.hpp:161-164): u_n = v_n = 0, or set them from host arrays."""
        check(self.lib.fus_model_set_state(self.h, capi.optional(u), capi.optional(v)),
              "fus_model_set_state")

    def f1(self, t, u, v):
        out = np.zeros(self.V.ndofs)
        check(self.lib.fus_model_f1(self.h, float(t), u, v, out), "fus_model_f1")
        return out

    def rk4(self, startTime, finalTime, timeStep):
        """rk4() (Linear.hpp:228-314).  Returns the number of steps taken."""
        n = C.c_int(0)
        check(self.lib.fus_model_rk4(self.h, float(startTime), float(finalTime), float(timeStep),
                                     C.byref(n)), "fus_model_rk4")
        return n.value

    def u_sol(self):
        u = np.zeros(self.V.ndofs)
        check(self.lib.fus_model_get_state(self.h, capi.optional(u), None), "fus_model_get_state")
        return u

    def v_sol(self):
        v = np.zeros(self.V.ndofs)
        check(self.lib.fus_model_get_state(self.h, None, capi.optional(v)), "fus_model_get_state")
        return v

    def mass(self):
        mvec = np.zeros(self.V.ndofs)
        check(self.lib.fus_model_get_mass(self.h, mvec), "fus_model_get_mass")
        return mvec

    def number_of_dofs(self):
        return self.V.ndofs

    def destroy(self):
        if getattr(self, "h", None):
            self.lib.fus_model_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass


class LinearSpectral3D(_Model):
    """LinearSpectral3D<T,P> (Linear.hpp:52-347)."""
    kind = "linear"

    def __init__(self, V, speedOfSound, density, sourceFrequency, sourceAmplitude, sourceSpeed,
                 **kw):
        super().__init__(V, speedOfSound, density, None, None, sourceFrequency, sourceAmplitude,
                         sourceSpeed, **kw)


class LossySpectral3D(_Model):
    """LossySpectral3D<T,P> (Lossy.hpp:54-373)."""
    kind = "lossy"

    def __init__(self, V, speedOfSound, density, diffusivityOfSound, sourceFrequency,
                 sourceAmplitude, sourceSpeed, **kw):
        super().__init__(V, speedOfSound, density, diffusivityOfSound, None, sourceFrequency,
                         sourceAmplitude, sourceSpeed, **kw)


class WesterveltSpectral3D(_Model):
    """WesterveltSpectral3D<T,P> (Westervelt.hpp:56-406)."""
    kind = "westervelt"

    def __init__(self, V, speedOfSound, density, diffusivityOfSound, coefficientOfNonlinearity,
                 sourceFrequency, sourceAmplitude, sourceSpeed, **kw):
        super().__init__(V, speedOfSound, density, diffusivityOfSound, coefficientOfNonlinearity,
                         sourceFrequency, sourceAmplitude, sourceSpeed, **kw)
