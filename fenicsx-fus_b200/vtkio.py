"""Field output after (or during) the time loop.

The reference's examples dump `u_n` with `dolfinx::io::VTXWriter` (ADIOS2 BP, e.g.
`cpp/fenicsx-sf-naive/examples/linear_planewave2d_1/main.cpp:144-158`) for ParaView.  ADIOS2 is a
third-party library that is not in this image, so the same step is offered as VTK XML
unstructured-grid files (`.vtu`, plus a `.pvd` collection for time series), which ParaView and
VisIt read natively.  A degree-P cell is written as P^3 linear hexahedra (P^2 quadrilaterals for the
2-D variant, where the reference's examples actually write their output) on its GLL lattice: every
nodal value is kept exactly and no high-order cell support is needed in the reader.

Host-side post-processing only: fields come from `model.u_sol()` / `model.v_sol()`.
"""
import base64
import os
import struct

import numpy as np

from . import gll

# VTK_HEXAHEDRON (type 12) vertex order in terms of (dx, dy, dz) offsets
_VTK_HEX = ((0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 1, 0), (0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 1, 1))


def gll_lattice_order(P):
    """Positions of the 1-D nodes sorted by coordinate: Basix order is [0, 1, interior...]."""
    pts, _ = gll(P)
    return np.argsort(pts, kind="stable")


_VTK_QUAD = ((0, 0), (1, 0), (1, 1), (0, 1))


def subcell_connectivity(dofmap, P, dim=3):
    """(ncells * P^3, 8) linear hexahedra -- or (ncells * P^2, 4) quadrilaterals for dim = 2 -- in
    VTK vertex order from the tensor dofmap `dofmap[c, i0*N*N + i1*N + i2]` / `dofmap[c, i0*N + i1]`
    (permute.hpp:15-42 ordering)."""
    N = P + 1
    order = gll_lattice_order(P)                       # lattice position -> 1-D node index
    if dim == 2:
        dm2 = np.asarray(dofmap).reshape(-1, N, N)[:, order][:, :, order]
        conn2 = np.empty((dm2.shape[0], P, P, 4), dtype=np.int64)
        for k, (dx, dy) in enumerate(_VTK_QUAD):
            conn2[..., k] = dm2[:, dx:dx + P, dy:dy + P]
        return conn2.reshape(-1, 4)
    dm = np.asarray(dofmap).reshape(-1, N, N, N)[:, order][:, :, order][:, :, :, order]
    conn = np.empty((dm.shape[0], P, P, P, 8), dtype=np.int64)
    for k, (dx, dy, dz) in enumerate(_VTK_HEX):
        conn[..., k] = dm[:, dx:dx + P, dy:dy + P, dz:dz + P]
    return conn.reshape(-1, 8)


def _data_array(name, arr, ncomp=None, binary=True):
    arr = np.ascontiguousarray(arr)
    vtk_type = {"float64": "Float64", "float32": "Float32", "int64": "Int64", "int32": "Int32",
                "uint8": "UInt8"}[arr.dtype.name]
    comp = f' NumberOfComponents="{ncomp}"' if ncomp else ""
    if binary:
        raw = arr.tobytes()
        payload = base64.b64encode(struct.pack("<Q", len(raw))).decode() + base64.b64encode(raw).decode()
        fmt = "binary"
    else:
        payload = " ".join(repr(v) for v in arr.reshape(-1).tolist())
        fmt = "ascii"
    return f'<DataArray type="{vtk_type}" Name="{name}"{comp} format="{fmt}">{payload}</DataArray>\n'


def write_vtu(path, V, fields, binary=True, dof_coordinates=None):
    """Write nodal fields {name: array of V.ndofs} of the function space V (FunctionSpace or
    HexFunctionSpace) to `path` (.vtu).  Returns (npoints, ncells) written."""
    X = V.tabulate_dof_coordinates() if dof_coordinates is None else dof_coordinates
    X = np.asarray(X, dtype=np.float64)
    if X.shape != (V.ndofs, 3):
        raise ValueError("dof coordinates must be (ndofs, 3)")
    dim = getattr(V, "dim", 3)
    conn = subcell_connectivity(V.dofmap, V.P, dim)
    nsub, nvc = conn.shape                             # sub-cells, vertices per sub-cell
    for name, f in fields.items():
        if np.asarray(f).shape != (V.ndofs,):
            raise ValueError(f"field {name!r} must have {V.ndofs} entries")
    with open(path, "w") as out:
        out.write('<?xml version="1.0"?>\n<VTKFile type="UnstructuredGrid" version="1.0" '
                  'byte_order="LittleEndian" header_type="UInt64">\n<UnstructuredGrid>\n')
        out.write(f'<Piece NumberOfPoints="{V.ndofs}" NumberOfCells="{nsub}">\n')
        out.write("<Points>\n" + _data_array("Points", X, 3, binary) + "</Points>\n")
        out.write("<Cells>\n")
        out.write(_data_array("connectivity", conn.reshape(-1), None, binary))
        out.write(_data_array("offsets", nvc * np.arange(1, nsub + 1, dtype=np.int64), None, binary))
        out.write(_data_array("types", np.full(nsub, 12 if dim == 3 else 9, dtype=np.uint8), None,
                              binary))                 # VTK_HEXAHEDRON / VTK_QUAD
        out.write("</Cells>\n")
        first = next(iter(fields), None)
        out.write(f'<PointData Scalars="{first}">\n' if first else "<PointData>\n")
        for name, f in fields.items():
            out.write(_data_array(name, np.asarray(f, dtype=np.float64), None, binary))
        out.write("</PointData>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n")
    return V.ndofs, nsub


class TimeSeriesWriter:
    """`.pvd` collection of `.vtu` snapshots: the role VTXWriter::write(t) plays in the reference's
    examples.  The mesh part is recomputed once and reused for every snapshot."""

    def __init__(self, path, V, binary=True):
        self.base = os.path.splitext(path)[0]
        self.V, self.binary, self.entries = V, binary, []
        self.X = V.tabulate_dof_coordinates()

    def write(self, t, **fields):
        name = f"{self.base}_{len(self.entries):06d}.vtu"
        write_vtu(name, self.V, fields, self.binary, dof_coordinates=self.X)
        self.entries.append((float(t), os.path.basename(name)))
        self._write_index()
        return name

    def _write_index(self):
        with open(self.base + ".pvd", "w") as f:
            f.write('<?xml version="1.0"?>\n<VTKFile type="Collection" version="0.1">\n<Collection>\n')
            for t, name in self.entries:
                f.write(f'<DataSet timestep="{t!r}" part="0" file="{name}"/>\n')
            f.write("</Collection>\n</VTKFile>\n")

    def close(self):
        self._write_index()


def read_vtu_arrays(path):
    """Minimal reader of the files written above (tests, round trips): {name: ndarray}."""
    import xml.etree.ElementTree as ET
    np_type = {"Float64": np.float64, "Float32": np.float32, "Int64": np.int64, "Int32": np.int32,
               "UInt8": np.uint8}
    out = {}
    for da in ET.parse(path).getroot().iter("DataArray"):
        dt = np_type[da.get("type")]
        text = (da.text or "").strip()
        if da.get("format") == "binary":
            head = base64.b64decode(text[:12])
            nbytes = struct.unpack("<Q", head)[0]
            arr = np.frombuffer(base64.b64decode(text[12:]), dtype=dt)
            assert arr.nbytes == nbytes
        else:
            arr = np.array(text.split(), dtype=dt)
        nc = da.get("NumberOfComponents")
        out[da.get("Name")] = arr.reshape(-1, int(nc)) if nc else arr
    return out
