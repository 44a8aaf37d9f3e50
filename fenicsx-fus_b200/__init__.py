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

__all__ = ["BoxMesh", "RectMesh", "FunctionSpace", "HexMesh", "HexFunctionSpace",
           "StiffnessSpectral3D", "MassSpectral3D", "LinearSpectral3D", "LossySpectral3D",
           "WesterveltSpectral3D", "StiffnessSpectral2D", "MassSpectral2D", "LinearSpectral2D",
           "LossySpectral2D", "WesterveltSpectral2D", "gll",
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


class RectMesh:
    """Structured quadrilateral rectangle (dolfinx::mesh::create_rectangle stand-in) for the 2-D
    operators of cpp/fenicsx-sf-naive.

    x: (nverts,3) vertex coordinates padded with z = 0 as DOLFINx stores them, xdofmap: (ncells,4)
    DOLFINx vertex order v = a + 2b, facets: (nfacets,3) exterior edges {cell, local facet, tag}
    with tag 1 on x=lo, 2 on x=hi.  `warp(x) -> x'` displaces the vertices (bilinear cells)."""
    dim = 2

    def __init__(self, n, lo=(0.0, 0.0), hi=(1.0, 1.0), warp=None):
        lib = capi.load()
        self.n = np.asarray(n, dtype=np.int32)
        assert self.n.shape == (2,)
        nx, ny = (int(v) for v in self.n)
        self.lo = np.asarray(lo, dtype=np.float64)
        self.hi = np.asarray(hi, dtype=np.float64)
        self.x = np.zeros(((nx + 1) * (ny + 1), 3))
        self.xdofmap = np.zeros((nx * ny, 4), dtype=np.int32)
        check(lib.fus_rect_mesh(self.n, self.lo, self.hi, self.x, self.xdofmap), "fus_rect_mesh")
        if warp is not None:
            self.x = np.ascontiguousarray(warp(self.x), dtype=np.float64)
            self.x[:, 2] = 0.0
        nf = lib.fus_rect_facets(self.n, None)
        self.facets = np.zeros((nf, 3), dtype=np.int32)
        lib.fus_rect_facets(self.n, self.facets.ctypes.data_as(C.c_void_p))
        self.ncells = nx * ny

    def h_min(self):
        h = (self.hi - self.lo) / self.n
        return float(np.sqrt((h * h).sum()))


class FunctionSpace:
    """Degree-P GLL Lagrange space on a BoxMesh (hexahedra) or a RectMesh (quadrilaterals) with
    the tensor-product dofmap (create_functionspace + reorder_dofmap, permute.hpp:15-42)."""

    def __init__(self, mesh, P, numbering=1):
        lib = capi.load()
        self.mesh, self.P = mesh, int(P)
        self.N = self.P + 1
        self.dim = getattr(mesh, "dim", 3)
        if self.dim == 2:
            self.ndofs = int(lib.fus_rect_num_dofs(self.P, mesh.n))
            self.dofmap = np.zeros((mesh.ncells, self.N ** 2), dtype=np.int32)
            check(lib.fus_rect_dofmap(self.P, mesh.n, self.dofmap), "fus_rect_dofmap")
        else:
            self.ndofs = int(lib.fus_box_num_dofs(self.P, mesh.n))
            self.dofmap = np.zeros((mesh.ncells, self.N ** 3), dtype=np.int32)
            check(lib.fus_box_dofmap(self.P, mesh.n, numbering, self.dofmap), "fus_box_dofmap")
        self.nowned = self.ndofs
        self._ctx = None

    def context(self, device=0, lean=False):
        """The device context of this space (created on first use).  lean=True keeps neither G
        nor detJ on the device: geometric factors are rebuilt from the trilinear cell map."""
        if self._ctx is None:
            self._ctx = Context.from_mesh(self, device, lean=lean)
        return self._ctx

    def tabulate_dof_coordinates(self):
        """Physical coordinates of every dof (tri-/bilinear map of the GLL nodes)."""
        pts, _ = gll(self.P)
        N, m = self.N, self.mesh
        if self.dim == 2:
            X = m.x[m.xdofmap]                  # (nc, 4, 3)
            xi = np.stack(np.meshgrid(pts, pts, indexing="ij"), -1).reshape(-1, 2)
            acc = 0.0
            for v in range(4):
                w = (xi[:, 0] if v & 1 else 1 - xi[:, 0]) * (xi[:, 1] if v >> 1 else 1 - xi[:, 1])
                acc = acc + w[None, :, None] * X[:, v, None, :]
            out = np.zeros((self.ndofs, 3))
            out[self.dofmap.reshape(-1)] = acc.reshape(-1, 3)
            return out
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

    def __init__(self, handle, P, ncells, ndofs, nowned, device, dim=3):
        self.h, self.P, self.ncells, self.ndofs = handle, P, ncells, ndofs
        self.nowned, self.device, self.dim = nowned, device, dim
        self.lib = capi.load()

    @classmethod
    def from_mesh(cls, V, device=0, dofmap=None, ndofs=None, nowned=None, lean=False):
        lib = capi.load()
        m = V.mesh
        dm = V.dofmap if dofmap is None else dofmap
        ndofs = V.ndofs if ndofs is None else ndofs
        nowned = ndofs if nowned is None else nowned
        h = C.c_void_p()
        dim = getattr(m, "dim", 3)
        if dim == 2:
            if lean:
                raise FusError("lean contexts are a hexahedral feature")
            create = lib.fus_ctx_create_from_mesh_2d
        else:
            create = lib.fus_ctx_create_from_mesh_lean if lean else lib.fus_ctx_create_from_mesh
        check(create(V.P, dm.shape[0], ndofs, nowned, np.ascontiguousarray(dm), m.x.shape[0], m.x,
                     m.xdofmap, device, C.byref(h)), "fus_ctx_create_from_mesh")
        return cls(h, V.P, dm.shape[0], ndofs, nowned, device, dim)

    @classmethod
    def from_arrays(cls, P, dofmap, ndofs, G, detJ, dphi, device=0, nowned=None, dim=3):
        lib = capi.load()
        h = C.c_void_p()
        nowned = ndofs if nowned is None else nowned
        create = lib.fus_ctx_create if dim == 3 else lib.fus_ctx_create_2d
        check(create(P, dofmap.shape[0], ndofs, nowned, np.ascontiguousarray(dofmap),
                     capi.optional(G), capi.optional(detJ), dphi, device, C.byref(h)),
              "fus_ctx_create")
        return cls(h, P, dofmap.shape[0], ndofs, nowned, device, dim)

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
        Nd = (self.P + 1) ** self.dim
        G = np.zeros((self.ncells, Nd, 6 if self.dim == 3 else 3)) if want_G else None
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
        """Release the device data.  Refused (the handle stays valid) while models built on this
        context are alive: destroy them first."""
        if self.h and self.lib.fus_ctx_destroy(self.h) == 0:
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
        # float32 data selects the FP32 instantiation (the reference's T = float operators)
        f32 = "float32" in str(x.dtype)
        if _is_device_tensor(x):
            fn = self._dev_fn.replace("_dev", "_f32_dev") if f32 else self._dev_fn
            check(getattr(lib, fn)(self.ctx.h, C.c_void_p(x.data_ptr()), C.c_void_p(coeffs.data_ptr()),
                                   C.c_void_p(y.data_ptr())), fn)
        else:
            fn = self._host_fn.replace("_host", "_f32_host") if f32 else self._host_fn
            coeffs = np.ascontiguousarray(coeffs, dtype=np.float32 if f32 else np.float64)
            check(getattr(lib, fn)(self.ctx.h, x, coeffs, y), fn)
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
        bvec = lib.fus_boundary_vectors_2d if getattr(m, "dim", 3) == 2 else lib.fus_boundary_vectors
        check(bvec(k, V.P, nc, nd, m.x, m.xdofmap, V.dofmap, facets.shape[0], facets, c0, rho0,
                   capi.optional(delta0), capi.optional(src), capi.optional(dsrc),
                   capi.optional(absb), capi.optional(bmass)), "fus_boundary_vectors")
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


# 2-D quadrilateral variant (cpp/fenicsx-sf-naive/common): the same classes on a RectMesh space --
# the library dispatches on the dimension of the context.
class StiffnessSpectral2D(StiffnessSpectral3D):
    """StiffnessSpectral2D<T,P>::operator() (fenicsx-sf-naive spectral_op.hpp:226-359)."""


class MassSpectral2D(MassSpectral3D):
    """MassSpectral2D<T,P>::operator() (fenicsx-sf-naive spectral_op.hpp:28-107)."""


class LinearSpectral2D(LinearSpectral3D):
    """LinearSpectral2D<T,P> (fenicsx-sf-naive Linear.hpp:52-350)."""


class LossySpectral2D(LossySpectral3D):
    """LossySpectral2D<T,P> (fenicsx-sf-naive Lossy.hpp)."""


class WesterveltSpectral2D(WesterveltSpectral3D):
    """WesterveltSpectral2D<T,P> (fenicsx-sf-naive Westervelt.hpp)."""
