#!/usr/bin/env python
"""The reference's 2-D example cpp/fenicsx-sf-naive/examples/linear_planewave2d_1/main.cpp on the GPU,
on the example's own mesh (the reference ships it; a copy of its datasets is committed as
tests/golden/ref_mesh_quad8400.npz): LinearSpectral2D, P = 4, CFL 0.9, t_end = L/c + 4/f, then the
output step of the example (VTXWriter -> .vtu here) and the L2-type comparison with the analytic
plane wave on the axis.

    python examples/linear_planewave2d_1.py [--mesh /path/to/mesh.h5] [--out output_final.vtu]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import fenicsx_fus_b200 as fus
    from fenicsx_fus_b200 import sampling, vtkio
    from fenicsx_fus_b200.unstructured2d import QuadFunctionSpace, QuadMesh
    ap = argparse.ArgumentParser()
    ap.add_argument("--mesh", default=None, help="mesh.h5 of the example (default: committed copy)")
    ap.add_argument("--out", default="output_final.vtu")
    args = ap.parse_args()
    if args.mesh:
        mesh = QuadMesh.from_xdmf_h5(args.mesh, "planewave_2d_1")
    else:
        g = np.load(os.path.join(ROOT, "tests", "golden", "ref_mesh_quad8400.npz"))
        mesh = QuadMesh(g["geometry"], g["topology_vtk"][:, (0, 1, 3, 2)], g["facet_lines"],
                        g["facet_values"], g["cell_values"])
    # main.cpp:31-41
    sourceFrequency, sourceAmplitude = 0.5e6, 60000.0
    speedOfSound, density, domainLength, degreeOfBasis = 1500.0, 1000.0, 0.12, 4
    period = 1.0 / sourceFrequency
    V = QuadFunctionSpace(mesh, degreeOfBasis)
    c0 = np.where(mesh.cell_tags == 1, speedOfSound, speedOfSound)       # one medium (tag 1)
    rho0 = np.full(mesh.ncells, density)
    # main.cpp:103-110
    CFL = 0.9
    dt = CFL * mesh.h_min() / (speedOfSound * degreeOfBasis ** 2)
    dt = period / (int(period / dt) + 1)
    tf = domainLength / speedOfSound + 4.0 / sourceFrequency
    model = fus.LinearSpectral2D(V, c0, rho0, sourceFrequency, sourceAmplitude, speedOfSound)
    model.init()
    steps = model.rk4(0.0, tf, dt)
    u = model.u_sol()
    print(f"Degrees of freedom: {V.ndofs}\nTime step size: {dt}\nNumber of steps: {steps}")
    vtkio.write_vtu(args.out, V, {"u": u})
    pts, vals = sampling.eval_line(V, u, (0.0, 0.0), (domainLength, 0.0), 481)
    t_end = tf                       # the last step of rk4 is shortened to land on tf (Linear.hpp:271)
    exact = sourceAmplitude * np.sin(2 * np.pi * sourceFrequency * (t_end - pts[:, 0] / speedOfSound))
    behind = pts[:, 0] < 0.9 * speedOfSound * (t_end - 4.0 / sourceFrequency)
    err = np.sqrt(((vals - exact)[behind] ** 2).sum() / (exact[behind] ** 2).sum())
    print(f"Relative L2 error on the axis behind the front: {err:.3e}\nwrote {args.out}")


if __name__ == "__main__":
    main()
