#!/usr/bin/env python
"""Torch-free timing probe: one stiffness application per geometry mode (0 streamed G, 2 rebuilt
from the trilinear cell map), timed by the library's own CUDA-event pairs (fus_ctx_profile).

    python scripts/probe_geometry_modes.py [P:n ...]        default 4:54 2:107 6:36

Writes one JSON line per (P, mode) to stdout as soon as it is measured.  Meant for short GPU slots:
no torch import, a few seconds per degree.
"""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import fenicsx_fus_b200 as fus
    cases = [tuple(int(v) for v in a.split(":")) for a in sys.argv[1:]] or [(4, 54), (2, 107), (6, 36)]
    peak = 6555.2
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    for P, n in cases:
        t0 = time.perf_counter()
        m = fus.BoxMesh((n, n, n))
        V = fus.FunctionSpace(m, P, numbering=1)
        ctx = V.context()
        lib = ctx.lib
        nd, nc = V.ndofs, m.ncells
        dx, dy, dc = ctx.alloc(8 * nd), ctx.alloc(8 * nd), ctx.alloc(8 * nc)
        X = V.tabulate_dof_coordinates() if nd < 2_000_000 else None
        x = np.sin(np.arange(nd) * 1e-3) if X is None else np.sin(X[:, 0]) * np.cos(np.pi * X[:, 1])
        ctx.upload(dx, x)
        ctx.upload(dc, np.full(nc, -1e-3))
        setup_s = time.perf_counter() - t0
        ref = None
        for mode in (0, 2):
            ctx.set_option("geometry_mode", mode)
            if ctx.get_option("geometry_compressed") != mode:
                continue
            fus.check(lib.fus_dev_memset(ctx.h, dy, 0, 8 * nd), "memset")
            fus.check(lib.fus_stiffness_apply_dev(ctx.h, dx, dc, dy), "apply")
            y = np.zeros(nd)
            ctx.download(y, dy)
            if ref is None:
                ref = y
            for _ in range(3):
                fus.check(lib.fus_stiffness_apply_dev(ctx.h, dx, dc, dy), "apply")
            ctx.sync()
            ctx.set_option("profile_kernels", 1)
            reps = 20
            for _ in range(reps):
                fus.check(lib.fus_stiffness_apply_dev(ctx.h, dx, dc, dy), "apply")
            nl, ms = ctx.profile("stiffness")
            ctx.set_option("profile_kernels", 0)
            alg = 52.0 * nc * (P + 1) ** 3 + 16.0 * nd
            avg = ms / max(nl, 1)
            print(json.dumps({"P": P, "n": n, "dofs": nd, "geometry_mode": mode, "launches": nl,
                              "ms_avg": avg, "gdof_per_s": nd / (avg * 1e-3) / 1e9,
                              "streamed_alg_gbs": alg / (avg * 1e-3) / 1e9,
                              "frac_of_measured_peak": alg / (avg * 1e-3) / 1e9 / peak,
                              "rel_l2_vs_mode0": float(np.linalg.norm(y - ref) / np.linalg.norm(ref)),
                              "setup_s": setup_s}), flush=True)
        ctx.set_option("geometry_mode", 0)
        for p in (dx, dy, dc):
            ctx.free(p)
        V._ctx = None
        ctx.destroy()


if __name__ == "__main__":
    main()
