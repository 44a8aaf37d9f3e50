#!/usr/bin/env python
"""Operator throughput through the UNSTRUCTURED path (general conforming numbering, HexMesh):
the same 54^3 P=4 box fed as an unstructured mesh, (a) in its natural cell order and (b) with the
cells randomly permuted -- how much of the operator's speed depends on mesh ordering -- and the
reference's own 6 312-cell test mesh.  One JSON object per line."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def time_operator(fus, torch, V, ncells, repeats=20):
    stream = torch.cuda.current_stream()
    ctx = V.context()
    ctx.set_stream(stream.cuda_stream)
    x = torch.rand(V.ndofs, dtype=torch.float64, device="cuda")
    y = torch.zeros_like(x)
    co = torch.full((ncells,), -1e-3, dtype=torch.float64, device="cuda")
    K = fus.StiffnessSpectral3D(V)
    for _ in range(3):
        K(x, co, y)
    ts = []
    for _ in range(repeats):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        K(x, co, y)
        e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))


def main():
    import torch

    import fenicsx_fus_b200 as fus
    from fenicsx_fus_b200.unstructured import HexFunctionSpace, HexMesh
    P, n = 4, 54
    box = fus.BoxMesh((n, n, n))
    Vb = fus.FunctionSpace(box, P, numbering=1)
    tmin, tmed = time_operator(fus, torch, Vb, box.ncells)
    print(json.dumps({"mesh": "box generator (cell-blocked numbering)", "dofs": Vb.ndofs,
                      "ms_min": tmin, "gdof_per_s": Vb.ndofs / tmin / 1e6}), flush=True)
    Vb._ctx.destroy()
    rng = np.random.default_rng(0)
    for label, perm in (("unstructured path, natural cell order", np.arange(box.ncells)),
                        ("unstructured path, random cell order", rng.permutation(box.ncells))):
        for reorder in ((None,) if label.endswith("natural cell order") else (None, "morton")):
            m = HexMesh(box.x, box.xdofmap[perm], reorder=reorder)
            V = HexFunctionSpace(m, P)
            tmin, tmed = time_operator(fus, torch, V, m.ncells)
            print(json.dumps({"mesh": label + (", Morton-reordered" if reorder else ""),
                              "dofs": V.ndofs, "ms_min": tmin,
                              "gdof_per_s": V.ndofs / tmin / 1e6}), flush=True)
            V._ctx.destroy()
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_mesh_hex6312.npz"))
    m = HexMesh(g["geometry"], g["topology_vtk"][:, (0, 1, 3, 2, 4, 5, 7, 6)], g["facet_quads"],
                g["facet_values"])
    for P in (4, 6):
        V = HexFunctionSpace(m, P)
        tmin, tmed = time_operator(fus, torch, V, m.ncells)
        print(json.dumps({"mesh": f"reference test mesh (6312 cells), P={P}", "dofs": V.ndofs,
                          "ms_min": tmin, "gdof_per_s": V.ndofs / tmin / 1e6}), flush=True)


if __name__ == "__main__":
    main()
