"""Unstructured quadrilateral meshes for the 2-D variant (SURVEY.md sections 8f-1 and 8f-4).

The reference ships the meshes of its 2-D examples (`cpp/fenicsx-sf-naive/examples/*/mesh.h5`,
`python/examples/*/mesh.h5`: quadrilaterals in VTK vertex order, cell tags, PolyLine facet tags);
this module stands in for what DOLFINx does when it reads them (`io::XDMFFile::read_mesh`,
`read_meshtags`, `fem::create_functionspace`, `reorder_dofmap`, permute.hpp:15-42):
* `QuadMesh.from_xdmf_h5` parses the HDF5 companion (fenicsx-fus_b200/hdf5min.py), permutes VTK
  order (0,1,2,3 counter-clockwise) to the DOLFINx tensor order v = a + 2b, finds the exterior edges
  and attaches facet and cell tags;
* `QuadFunctionSpace` numbers vertex / edge-interior / cell-interior GLL nodes conformingly (edges
  oriented from the lower to the higher global vertex) and emits the dofmap in tensor order
  (index i0*N + i1, 1-D node order [0, 1, interior ascending]).
The kernels only see x, xdofmap and the dofmap: `fus_ctx_create_from_mesh_2d`.
"""
import numpy as np

from . import capi, hdf5min

_VTK_TO_TENSOR_2D = (0, 1, 3, 2)
# DOLFINx quadrilateral facets: 0: xi1=0 (v0,v1), 1: xi0=0 (v0,v2), 2: xi0=1 (v1,v3), 3: xi1=1 (v2,v3)
_FACET_VERTS = ((0, 1), (0, 2), (1, 3), (2, 3))


class QuadMesh:
    """x: (nverts,3) padded with z = 0; xdofmap: (ncells,4) in DOLFINx tensor vertex order;
    facets: (nfacets,3) exterior edges {cell, local facet, tag} (tag 0 when untagged);
    cell_tags: (ncells,) (0 when untagged)."""
    dim = 2

    def __init__(self, x, cells_tensor, facet_lines=None, facet_values=None, cell_values=None):
        x = np.asarray(x, dtype=np.float64)
        self.x = np.zeros((x.shape[0], 3))
        self.x[:, :2] = x[:, :2]
        self.xdofmap = np.ascontiguousarray(cells_tensor, dtype=np.int32)
        self.ncells = self.xdofmap.shape[0]
        self.n = None
        c = self.xdofmap.astype(np.int64)
        nv = self.x.shape[0]
        ends = np.stack([np.stack([c[:, a], c[:, b]], axis=-1) for a, b in _FACET_VERTS], axis=1)
        key = np.minimum(ends[..., 0], ends[..., 1]) * nv + np.maximum(ends[..., 0], ends[..., 1])
        uniq, inv, counts = np.unique(key.reshape(-1), return_inverse=True, return_counts=True)
        if np.any(counts > 2):
            raise ValueError("non-manifold mesh: an edge is shared by more than two cells")
        self.nedges = uniq.size
        self.cell_edges = inv.reshape(self.ncells, 4)
        self.edge_rev = ends[..., 0] > ends[..., 1]           # local direction vs low->high global
        idx = np.flatnonzero(counts[inv] == 1)
        tags = np.zeros(idx.size, dtype=np.int32)
        if facet_lines is not None and len(facet_lines):
            fl = np.asarray(facet_lines, dtype=np.int64)
            fkey = np.minimum(fl[:, 0], fl[:, 1]) * nv + np.maximum(fl[:, 0], fl[:, 1])
            lut = dict(zip(fkey.tolist(), np.ravel(facet_values).astype(int).tolist()))
            tags = np.array([lut.get(int(k), 0) for k in key.reshape(-1)[idx]], dtype=np.int32)
        self.facets = np.ascontiguousarray(np.stack([idx // 4, idx % 4, tags], axis=1), dtype=np.int32)
        self.cell_tags = (np.zeros(self.ncells, dtype=np.int32) if cell_values is None
                          else np.ravel(cell_values).astype(np.int32))

    @classmethod
    def from_xdmf_h5(cls, h5_path, name):
        f = hdf5min.File(h5_path)
        topo = f.read(f"/Mesh/{name}/topology")
        geom = f.read(f"/Mesh/{name}/geometry")
        lines = vals = cvals = None
        try:
            lines = f.read(f"/MeshTags/{name}_facets/topology")
            vals = f.read(f"/MeshTags/{name}_facets/Values")
        except KeyError:
            pass
        try:
            ctopo = f.read(f"/MeshTags/{name}_cells/topology")
            cv = f.read(f"/MeshTags/{name}_cells/Values")
            if np.array_equal(ctopo, topo):                   # DOLFINx writes them in cell order
                cvals = cv
            else:
                lut = {tuple(sorted(r)): int(v) for r, v in zip(ctopo.tolist(), np.ravel(cv).tolist())}
                cvals = np.array([lut.get(tuple(sorted(r)), 0) for r in topo.tolist()])
        except KeyError:
            pass
        return cls(geom, topo[:, _VTK_TO_TENSOR_2D], lines, vals, cvals)

    def h_min(self):
        """Smallest cell diameter (largest vertex-to-vertex distance per cell), like mesh::h."""
        X = self.x[self.xdofmap]
        d = np.linalg.norm(X[:, :, None, :] - X[:, None, :, :], axis=-1)
        return float(d.reshape(self.ncells, -1).max(axis=1).min())


class QuadFunctionSpace:
    """Degree-P GLL Lagrange space on a QuadMesh with the tensor-product dofmap."""
    dim = 2

    def __init__(self, mesh, P, renumber=True):
        self.mesh, self.P, self.N = mesh, int(P), int(P) + 1
        P, N = self.P, self.N
        c = mesh.xdofmap.astype(np.int64)
        nc, nv, ne = mesh.ncells, mesh.x.shape[0], mesh.nedges
        pos = [0, P] + list(range(1, P))                       # Basix node -> grid offset
        m = P - 1
        off_e, off_c = nv, nv + ne * m
        self.ndofs = int(off_c + nc * m * m)
        dm = np.empty((nc, N * N), dtype=np.int64)
        cells = np.arange(nc, dtype=np.int64)
        for i0 in range(N):
            for i1 in range(N):
                p0, p1 = pos[i0], pos[i1]
                on0, on1 = p0 in (0, P), p1 in (0, P)
                col = i0 * N + i1
                if on0 and on1:
                    dm[:, col] = c[:, int(p0 == P) + 2 * int(p1 == P)]
                elif on1:                                       # edge along xi0: facet 0 or 3
                    lf = 0 if p1 == 0 else 3
                    idx = np.where(mesh.edge_rev[:, lf], m - p0, p0 - 1)
                    dm[:, col] = off_e + mesh.cell_edges[:, lf] * m + idx
                elif on0:                                       # edge along xi1: facet 1 or 2
                    lf = 1 if p0 == 0 else 2
                    idx = np.where(mesh.edge_rev[:, lf], m - p1, p1 - 1)
                    dm[:, col] = off_e + mesh.cell_edges[:, lf] * m + idx
                else:
                    dm[:, col] = off_c + cells * (m * m) + (p0 - 1) * m + (p1 - 1)
        if renumber:                                            # first appearance in the cell loop
            flat = dm.reshape(-1)
            uniq, first = np.unique(flat, return_index=True)
            new_id = np.empty(self.ndofs, dtype=np.int64)
            new_id[uniq[np.argsort(first, kind="stable")]] = np.arange(self.ndofs)
            dm = new_id[dm]
        self.dofmap = np.ascontiguousarray(dm, dtype=np.int32)
        self.nowned = self.ndofs
        self.counts = dict(vertices=nv, edges=ne, cells=nc)
        self._ctx = None

    def context(self, device=0, lean=False):
        if self._ctx is None:
            from . import Context
            self._ctx = Context.from_mesh(self, device, lean=lean)
        return self._ctx

    def tabulate_dof_coordinates(self, return_spread=False):
        """Physical coordinates of every dof (bilinear map of the GLL nodes); with return_spread
        also the largest disagreement between cells sharing a dof (zero up to rounding iff the
        numbering is conforming)."""
        lib = capi.load()
        pts, wts = np.zeros(self.N), np.zeros(self.N)
        capi.check(lib.fus_gll(self.P, pts, wts), "fus_gll")
        m = self.mesh
        X = m.x[m.xdofmap]                                      # (nc, 4, 3)
        xi = np.stack(np.meshgrid(pts, pts, indexing="ij"), -1).reshape(-1, 2)
        acc = 0.0
        for v in range(4):
            w = (xi[:, 0] if v & 1 else 1 - xi[:, 0]) * (xi[:, 1] if v >> 1 else 1 - xi[:, 1])
            acc = acc + w[None, :, None] * X[:, v, None, :]
        flat, pc = self.dofmap.reshape(-1), acc.reshape(-1, 3)
        out = np.zeros((self.ndofs, 3))
        out[flat] = pc
        if not return_spread:
            return out
        return out, float(np.abs(out[flat] - pc).max())
