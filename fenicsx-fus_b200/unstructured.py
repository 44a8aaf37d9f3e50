"""Unstructured hexahedral meshes: ingestion and degree-P GLL dof numbering.

Stands in for what the reference gets from DOLFINx when it reads a mesh
(`io::XDMFFile::read_mesh`, `fem::create_functionspace`, then `reorder_dofmap`, permute.hpp:15-42):
* `HexMesh.from_xdmf_h5` reads the HDF5 companion of an XDMF file (topology in VTK vertex order is
  permuted to the DOLFINx tensor order (0,1,3,2,4,5,7,6)), finds the exterior
  facets and attaches the facet tags;
* `HexFunctionSpace` numbers the GLL nodes of all cells conformingly: vertex, edge-interior,
  face-interior and cell-interior dofs, with edges oriented from the lower to the higher global vertex
  and faces framed at their lowest vertex, and emits the dofmap directly in tensor-product order
  (1-D node order [0, 1, interior ascending]).  Dofs are finally renumbered by first appearance in
  the cell loop, which keeps a cell's dofs close together in memory.
The kernels are mesh-agnostic; only this host-side set-up differs from the box path.
"""
import numpy as np

from . import capi, hdf5min

_VTK_TO_TENSOR = (0, 1, 3, 2, 4, 5, 7, 6)
# DOLFINx hexahedron facet numbering: (fixed direction, side) -> local facet
_FACET_ID = {(2, 0): 0, (1, 0): 1, (0, 0): 2, (0, 1): 3, (1, 1): 4, (2, 1): 5}


def _vertex(bits):
    """local vertex number of reference coordinates bits = (a, b, c) in {0,1}^3 (x fastest)."""
    return bits[0] + 2 * bits[1] + 4 * bits[2]


def morton_order(centroids, bits=10):
    """Permutation that sorts points along a Z-order curve: neighbouring cells end up close in
    memory, which is what the gather/scatter of the operator wants from a mesh in arbitrary order
    (measured: 0.272 ms -> see profiles/ for a randomly ordered 54^3 box at P=4)."""
    c = np.asarray(centroids, dtype=np.float64)
    lo, hi = c.min(axis=0), c.max(axis=0)
    q = ((c - lo) / np.where(hi > lo, hi - lo, 1.0) * ((1 << bits) - 1)).astype(np.uint64)
    key = np.zeros(c.shape[0], dtype=np.uint64)
    for b in range(bits):
        for d in range(3):
            key |= ((q[:, d] >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b + (2 - d))
    return np.argsort(key, kind="stable")


class HexMesh:
    """x: (nverts,3); xdofmap: (ncells,8) in DOLFINx tensor vertex order;
    facets: (nfacets,3) exterior facets {cell, local facet, tag} (tag 0 when untagged)."""

    def __init__(self, x, cells_tensor, facet_quads=None, facet_values=None, reorder=None):
        self.x = np.ascontiguousarray(x, dtype=np.float64)
        cells_tensor = np.asarray(cells_tensor)
        # cell_perm[k] = index in the input of the cell stored at position k (per-cell data given
        # in input order must be permuted with it)
        self.cell_perm = np.arange(cells_tensor.shape[0])
        if reorder == "morton":
            self.cell_perm = morton_order(self.x[cells_tensor].mean(axis=1))
            cells_tensor = cells_tensor[self.cell_perm]
        elif reorder is not None:
            raise ValueError("reorder must be None or 'morton'")
        self.xdofmap = np.ascontiguousarray(cells_tensor, dtype=np.int32)
        self.ncells = self.xdofmap.shape[0]
        self.n = None
        self._build_faces()
        tags = np.zeros(self.ext_cell.size, dtype=np.int32)
        if facet_quads is not None and len(facet_quads):
            key = {tuple(sorted(int(v) for v in q)): int(val)
                   for q, val in zip(facet_quads, np.ravel(facet_values))}
            for k, fk in enumerate(self.ext_key):
                tags[k] = key.get(tuple(int(v) for v in fk), 0)
        self.facets = np.ascontiguousarray(
            np.stack([self.ext_cell, self.ext_lf, tags], axis=1), dtype=np.int32)

    @classmethod
    def from_xdmf_h5(cls, h5_path, name="hex", reorder="morton"):
        f = hdf5min.File(h5_path)
        topo = f.read(f"/Mesh/{name}/topology")
        geom = f.read(f"/Mesh/{name}/geometry")
        quads = vals = None
        try:
            quads = f.read(f"/MeshTags/{name}_facets/topology")
            vals = f.read(f"/MeshTags/{name}_facets/Values")
        except KeyError:
            pass
        return cls(geom, topo[:, _VTK_TO_TENSOR], quads, vals, reorder=reorder)

    def _build_faces(self):
        """Global face ids of the 6 faces of every cell and the exterior ones."""
        c = self.xdofmap.astype(np.int64)
        faces = []
        self._face_defs = []
        for d in range(3):
            e, f = [k for k in range(3) if k != d]
            for side in (0, 1):
                corners = []
                for s in (0, 1):
                    for t in (0, 1):
                        bits = [0, 0, 0]
                        bits[d], bits[e], bits[f] = side, s, t
                        corners.append(_vertex(bits))
                faces.append(c[:, corners])                 # (nc, 4) in (s,t) = 00,01,10,11 order
                self._face_defs.append((d, side, e, f))
        self.cell_face_corners = np.stack(faces, axis=1)     # (nc, 6, 4)
        keys = np.sort(self.cell_face_corners.reshape(-1, 4), axis=1)
        uniq, inv, counts = np.unique(keys, axis=0, return_inverse=True, return_counts=True)
        self.nfaces = uniq.shape[0]
        self.cell_faces = inv.reshape(self.ncells, 6)
        ext = counts[inv] == 1
        idx = np.flatnonzero(ext)
        self.ext_cell = (idx // 6).astype(np.int32)
        self.ext_lf = np.array([_FACET_ID[(self._face_defs[k][0], self._face_defs[k][1])]
                                for k in idx % 6], dtype=np.int32)
        self.ext_key = keys[idx]
        if np.any(counts > 2):
            raise ValueError("non-manifold mesh: a face is shared by more than two cells")

    def h_min(self):
        """Smallest cell diameter (largest vertex-to-vertex distance per cell), like mesh::h."""
        X = self.x[self.xdofmap]
        d = np.linalg.norm(X[:, :, None, :] - X[:, None, :, :], axis=-1)
        return float(d.reshape(self.ncells, -1).max(axis=1).min())


class HexFunctionSpace:
    """Degree-P GLL Lagrange space on a HexMesh with the tensor-product dofmap."""

    def __init__(self, mesh, P, renumber=True):
        self.mesh, self.P, self.N = mesh, int(P), int(P) + 1
        P, N = self.P, self.N
        c = mesh.xdofmap.astype(np.int64)
        nc, nv = mesh.ncells, mesh.x.shape[0]
        pos = np.array([0, P] + list(range(1, P)), dtype=np.int64)    # Basix node -> grid offset
        m = P - 1                                                     # interior nodes per edge

        # ---- global edges (12 per cell: free direction d, the two other coordinates fixed) ----
        edge_def, pairs = {}, []
        for d in range(3):
            e, f = [k for k in range(3) if k != d]
            for se in (0, 1):
                for sf in (0, 1):
                    bits = [0, 0, 0]
                    bits[e], bits[f] = se, sf
                    a = _vertex(bits)
                    bits[d] = 1
                    b = _vertex(bits)
                    edge_def[(d, se, sf)] = len(pairs)
                    pairs.append((a, b))
        ga = np.stack([c[:, a] for a, _ in pairs], axis=1)            # (nc, 12)
        gb = np.stack([c[:, b] for _, b in pairs], axis=1)
        ekeys = np.stack([np.minimum(ga, gb), np.maximum(ga, gb)], axis=-1).reshape(-1, 2)
        ekey1 = ekeys[:, 0] * nv + ekeys[:, 1]
        ue, einv = np.unique(ekey1, return_inverse=True)
        ne = ue.size
        cell_edges = einv.reshape(nc, 12)
        edge_rev = ga > gb                                            # local direction vs global

        # ---- face frames ------------------------------------------------------------------------
        fc = mesh.cell_face_corners                                   # (nc, 6, 4): (s,t)=00,01,10,11
        origin = np.argmin(fc, axis=2)                                # position of the lowest vertex
        os_, ot_ = origin // 2, origin % 2
        take = lambda s, t: np.take_along_axis(fc, (2 * s + t)[..., None], axis=2)[..., 0]  # noqa: E731
        n_s = take(1 - os_, ot_)                                      # neighbour along local s
        n_t = take(os_, 1 - ot_)                                      # neighbour along local t
        swap = n_t < n_s                                              # canonical first axis is t
        nf = mesh.nfaces

        off_e = nv
        off_f = off_e + ne * m
        off_c = off_f + nf * m * m
        self.ndofs = int(off_c + nc * m ** 3)
        if self.ndofs > np.iinfo(np.int32).max:
            raise ValueError("more than 2^31 dofs on one rank")
        dm = np.empty((nc, N ** 3), dtype=np.int64)
        cells = np.arange(nc, dtype=np.int64)
        for i0 in range(N):
            for i1 in range(N):
                for i2 in range(N):
                    p = (int(pos[i0]), int(pos[i1]), int(pos[i2]))
                    on = [q in (0, P) for q in p]
                    col = (i0 * N + i1) * N + i2
                    if all(on):
                        dm[:, col] = c[:, _vertex([q == P for q in p])]
                    elif sum(on) == 2:
                        d = on.index(False)
                        e, f = [k for k in range(3) if k != d]
                        le = edge_def[(d, int(p[e] == P), int(p[f] == P))]
                        j = p[d]                                     # 1..P-1 from the local start
                        idx = np.where(edge_rev[:, le], m - j, j - 1)
                        dm[:, col] = off_e + cell_edges[:, le] * m + idx
                    elif sum(on) == 1:
                        d = on.index(True)
                        e, f = [k for k in range(3) if k != d]
                        lf = 2 * d + int(p[d] == P)                  # order of mesh._face_defs
                        de = np.where(os_[:, lf] == 0, p[e], P - p[e])
                        df = np.where(ot_[:, lf] == 0, p[f], P - p[f])
                        s1 = np.where(swap[:, lf], df, de)
                        t1 = np.where(swap[:, lf], de, df)
                        dm[:, col] = off_f + mesh.cell_faces[:, lf] * (m * m) + (s1 - 1) * m + (t1 - 1)
                    else:
                        loc = ((p[0] - 1) * m + (p[1] - 1)) * m + (p[2] - 1)
                        dm[:, col] = off_c + cells * m ** 3 + loc
        if renumber:                                                  # first appearance in the cell loop
            flat = dm.reshape(-1)
            _, first = np.unique(flat, return_index=True)
            order = np.argsort(first, kind="stable")
            new_id = np.empty(self.ndofs, dtype=np.int64)
            new_id[np.unique(flat)[order]] = np.arange(self.ndofs)
            dm = new_id[dm]
        self.dofmap = np.ascontiguousarray(dm, dtype=np.int32)
        self.nowned = self.ndofs
        self.counts = dict(vertices=nv, edges=ne, faces=nf, cells=nc)
        self._ctx = None

    def context(self, device=0, lean=False):
        if self._ctx is None:
            from . import Context
            self._ctx = Context.from_mesh(self, device, lean=lean)
        return self._ctx

    def tabulate_dof_coordinates(self, return_spread=False):
        """Physical coordinates of every dof (trilinear map of the GLL nodes).  With
        return_spread also the largest disagreement between cells that share a dof -- zero up to
        rounding iff the numbering is conforming."""
        lib = capi.load()
        pts, wts = np.zeros(self.N), np.zeros(self.N)
        capi.check(lib.fus_gll(self.P, pts, wts), "fus_gll")
        m = self.mesh
        X = m.x[m.xdofmap]                                            # (nc, 8, 3)
        xi = np.stack(np.meshgrid(pts, pts, pts, indexing="ij"), -1).reshape(-1, 3)
        acc = 0.0
        for v in range(8):
            a, b, c = v & 1, (v >> 1) & 1, (v >> 2) & 1
            w = ((xi[:, 0] if a else 1 - xi[:, 0]) * (xi[:, 1] if b else 1 - xi[:, 1])
                 * (xi[:, 2] if c else 1 - xi[:, 2]))
            acc = acc + w[None, :, None] * X[:, v, None, :]
        flat = self.dofmap.reshape(-1)
        pc = acc.reshape(-1, 3)
        out = np.zeros((self.ndofs, 3))
        out[flat] = pc
        if not return_spread:
            return out
        spread = float(np.abs(out[flat] - pc).max())
        return out, spread
