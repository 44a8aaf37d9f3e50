#!/usr/bin/env python
"""Secondary measurements (not the headline bench line): BASELINE.json configs 2-4 on one GPU.

  * config 2: degree sweep P=2..7, single stiffness apply on a ~10 M-dof box
    (cpp/fenicsx-sf/experiments/measure_fraction_of_peak_performance/main.cpp:44-116):
    time = min over repeats, reported as Gdof/s and as fraction of the HBM roofline with the
    algorithmic bytes 52 r + 16 per dof;
  * config 3/4: RK4 steps of the heterogeneous linear, lossy and Westervelt models at P=4;
  * --geometry-modes: the same sweep with the geometric factors streamed (0, the reference's data and
    the roofline denominator) or rebuilt per point from the trilinear cell map (2, fus_trilinear.hpp);
    --rk4-geometry-modes: RK4 throughput of the headline workload (linear, P=4, 54^3) per mode and the
    relative L2 difference of the fields between modes.
Writes one JSON object per line to stdout.  bench.py runs this in a child process for its `extras`.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SWEEP = {2: 107, 3: 71, 4: 54, 5: 43, 6: 36, 7: 31}


def peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return float(json.load(open(p))["hbm_gbs"]) if os.path.exists(p) else 6650.0


def sweep_fp32(args, fus, torch, stream, pk):
    """FP32 operator instantiation per degree (StiffnessSpectral3D<float,P>: float data select the
    FP32 kernels), after everything else -- it is the newest code: if it fails, the rows printed
    before it are already out."""
    for P in [int(s) for s in args.degrees.split(",") if s]:
        n = SWEEP[P]
        m = fus.BoxMesh((n, n, n))
        V = fus.FunctionSpace(m, P, numbering=args.numbering)
        ctx = V.context()
        ctx.set_stream(stream.cuda_stream)
        x = torch.rand(V.ndofs, dtype=torch.float64, device="cuda")
        coeffs = torch.full((m.ncells,), -1.0 / 1000.0, dtype=torch.float64, device="cuda")
        K = fus.StiffnessSpectral3D(V)
        x32, c32 = x.float(), coeffs.float()
        y32, y64 = torch.zeros_like(x32), torch.zeros_like(x)
        K(x, coeffs, y64)
        K(x32, c32, y32)
        torch.cuda.synchronize()
        err = float((torch.linalg.vector_norm(y32.double() - y64)
                     / torch.linalg.vector_norm(y64)).item())
        for _ in range(3):
            K(x32, c32, y32)
        times = []
        for _ in range(args.repeats):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            K(x32, c32, y32)
            e1.record(stream)
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        tmin, tmed = min(times), float(np.median(times))
        alg32 = 28.0 * m.ncells * (P + 1) ** 3 + 8.0 * V.ndofs      # 24 B G + 4 B dofmap; x, y
        print(json.dumps({"config": "degree_sweep_fp32", "P": P, "n": n, "dofs": V.ndofs,
                          "ms_min": tmin, "ms_median": tmed,
                          "gdof_per_s": V.ndofs / (tmin * 1e-3) / 1e9,
                          "alg_gbs_median": alg32 / (tmed * 1e-3) / 1e9,
                          "frac_of_measured_peak": alg32 / (tmed * 1e-3) / 1e9 / pk,
                          "rel_l2_vs_fp64": err}), flush=True)
        del K, x, coeffs, x32, c32, y32, y64
        V._ctx = None
        ctx.destroy()
        torch.cuda.empty_cache()


def main():
    import torch

    import fenicsx_fus_b200 as fus
    ap = argparse.ArgumentParser()
    ap.add_argument("--degrees", default="2,3,4,5,6,7")
    ap.add_argument("--repeats", type=int, default=20)
    ap.add_argument("--models", default="linear_het,lossy,westervelt")
    ap.add_argument("--variants", default="0", help="-1 = the library's own choice per degree")
    ap.add_argument("--geometry-modes", default="0")
    ap.add_argument("--col-blocks-per-sm", type=int, default=0,
                    help="option col_blocks_per_sm for the degree sweep (0 = occupancy)")
    ap.add_argument("--pipeline-variants", default="",
                    help="extra stiffness variants timed with streamed G only (3,4,5: the line kernel's "
                         "experimental software pipelines), in the degree sweep and in the RK4 runs")
    ap.add_argument("--rk4-geometry-modes", default="")
    ap.add_argument("--fp32", action="store_true",
                    help="also time the FP32 operator instantiation and report its error vs FP64")
    ap.add_argument("--numbering", type=int, default=1)
    ap.add_argument("--rk4-cells", type=int, default=54, help="cells per direction of the RK4 runs")
    ap.add_argument("--rk4-steps", type=int, default=20, help="timed steps of the RK4 runs per mode")
    args = ap.parse_args()
    pk = peak()
    stream = torch.cuda.Stream()        # the legacy default stream cannot be graph-captured
    torch.cuda.set_stream(stream)
    gmodes = [int(s) for s in args.geometry_modes.split(",") if s]
    pvariants = [int(s) for s in args.pipeline_variants.split(",") if s]
    for P in [int(s) for s in args.degrees.split(",") if s]:
        n = SWEEP[P]
        m = fus.BoxMesh((n, n, n))
        V = fus.FunctionSpace(m, P, numbering=args.numbering)
        ctx = V.context()
        ctx.set_stream(stream.cuda_stream)
        X = None
        x = torch.rand(V.ndofs, dtype=torch.float64, device="cuda")
        y = torch.zeros_like(x)
        coeffs = torch.full((m.ncells,), -1.0 / 1000.0, dtype=torch.float64, device="cuda")
        K = fus.StiffnessSpectral3D(V)
        y_first = None
        pairs = [(int(s), g) for s in args.variants.split(",") for g in gmodes]
        pairs += [(v, 0) for v in pvariants]
        for variant, gmode in pairs:
            ctx.set_option("stiffness_variant", variant)
            ctx.set_option("geometry_mode", gmode)
            ctx.set_option("col_blocks_per_sm", args.col_blocks_per_sm)
            if ctx.get_option("geometry_compressed") != gmode:
                continue
            y.zero_()
            for _ in range(3):
                K(x, coeffs, y)
            times = []
            for _ in range(args.repeats):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                K(x, coeffs, y)
                e1.record(stream)
                torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1))
            tmin, tmed = min(times), float(np.median(times))
            npts = m.ncells * (P + 1) ** 3
            alg = 52.0 * npts + 16.0 * V.ndofs
            y1 = K(x, coeffs, torch.zeros_like(x))          # one clean application: same numbers?
            if y_first is None:
                y_first = y1
            err = float((torch.linalg.vector_norm(y1 - y_first)
                         / torch.linalg.vector_norm(y_first)).item())
            print(json.dumps({"config": "degree_sweep", "P": P, "n": n, "dofs": V.ndofs,
                              "variant": variant, "geometry_mode": gmode,
                              "col_blocks_per_sm": args.col_blocks_per_sm,
                              "rel_l2_vs_first_config": err,
                              "numbering": args.numbering,
                              "y_norm_after_repeats": float(torch.linalg.vector_norm(y).item()),
                              "ms_min": tmin, "ms_median": tmed,
                              "gdof_per_s": V.ndofs / (tmin * 1e-3) / 1e9,
                              "alg_gbs_min": alg / (tmin * 1e-3) / 1e9,
                              "alg_gbs_median": alg / (tmed * 1e-3) / 1e9,
                              "frac_of_measured_peak": alg / (tmed * 1e-3) / 1e9 / pk}), flush=True)
        ctx.set_option("stiffness_variant", -1)
        ctx.set_option("geometry_mode", 0)
        del K, x, y, coeffs
        V._ctx = None
        ctx.destroy()
        torch.cuda.empty_cache()

    # ---- models at P=4 on the 54^3 box ----------------------------------------------------
    P, n = 4, args.rk4_cells
    L = 0.12 * n / 54.0                             # same cell size whatever the box
    h = L / n
    models = [s for s in args.models.split(",") if s]
    rk4_modes = [int(s) for s in args.rk4_geometry_modes.split(",") if s]
    if rk4_modes:        # headline workload (bench.py) with the geometric factors from each source
        m = fus.BoxMesh((n, n, n), (0, 0, 0), (L, L, L))
        V = fus.FunctionSpace(m, P, numbering=args.numbering)
        ctx = V.context()
        ctx.set_stream(stream.cuda_stream)
        mdl = fus.LinearSpectral3D(V, 1500.0, 1000.0, 0.5e6, 6e4, 1500.0)
        dt0 = 0.65 * np.sqrt(3) * h / (1500.0 * P * P)
        dt = 2e-6 / (int(2e-6 / dt0) + 1)
        K, ref = args.rk4_steps, None
        for gmode, variant in [(g, -1) for g in rk4_modes] + [(0, v) for v in pvariants]:
            ctx.set_option("stiffness_variant", variant)
            ctx.set_option("geometry_mode", gmode)
            if ctx.get_option("geometry_compressed") != gmode:
                continue
            mdl.init()
            mdl.rk4(0.0, 2.5 * dt, dt)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            done = mdl.rk4(3 * dt, 3 * dt + (K - 0.5) * dt, dt)
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            ctx.set_option("profile_kernels", 1)
            mdl.rk4(0.0, (K - 0.5) * dt, dt)
            torch.cuda.synchronize()
            ctx.set_option("profile_kernels", 0)
            n_st, ms_st = ctx.profile("stiffness")
            mdl.init()
            mdl.rk4(0.0, (K - 0.5) * dt, dt)
            u = mdl.u_sol()
            if ref is None:
                ref = u
            print(json.dumps({"config": ("headline_rk4_by_geometry_mode" if variant < 0
                                         else "headline_rk4_by_pipeline_variant"),
                              "geometry_mode": gmode, "variant": variant,
                              "P": P, "dofs": V.ndofs, "steps": done, "ms_per_step": ms / done,
                              "dof_updates_per_s": V.ndofs * done / (ms * 1e-3),
                              "operator_ms": ms_st / max(n_st, 1),
                              "rel_l2_vs_first_mode": float(np.linalg.norm(u - ref)
                                                            / max(np.linalg.norm(ref), 1e-300)),
                              "u_norm": float(np.linalg.norm(u))}), flush=True)
        ctx.set_option("geometry_mode", 0)
        ctx.set_option("stiffness_variant", -1)
        mdl.destroy()
        V._ctx = None
        ctx.destroy()
    if models:
        m = fus.BoxMesh((n, n, n), (0, 0, 0), (L, L, L))
        V = fus.FunctionSpace(m, P, numbering=args.numbering)
        ctx = V.context()
        ctx.set_stream(stream.cuda_stream)
        nc = m.ncells
        cx = np.arange(nc) // (n * n)                   # layers along x by cell index
        # water / skin / cortical / trabecular / brain
        # (cpp/fenicsx-sf/experiments/measure_vector_assembly_speed/main.cpp:43-83)
        lay = np.minimum(cx * 5 // n, 4)
        c_tab = np.array([1500.0, 1610.0, 2800.0, 2300.0, 1560.0])
        r_tab = np.array([1000.0, 1090.0, 1850.0, 1700.0, 1040.0])
        f0 = 0.5e6
        for name in models:
            if name == "linear_het":
                c0, rho0 = c_tab[lay], r_tab[lay]
                mdl = fus.LinearSpectral3D(V, c0, rho0, f0, 6e4, 1500.0)
                cmax, cfl = 2800.0, 0.65
            elif name == "lossy":
                c0, rho0 = c_tab[lay], r_tab[lay]
                delta = fus.compute_diffusivity_of_sound(2 * np.pi * f0, c0, 5.0)
                mdl = fus.LossySpectral3D(V, c0, rho0, delta, f0, 6e4, 1500.0)
                cmax, cfl = 2800.0, 0.2      # `ds` absorbs on every facet: see DESIGN.md (stability)
            else:
                f0 = 1.1e6                    # HITU/W-H131-WATER/main.cpp:33-46
                c0, rho0 = np.full(nc, 1480.0), np.full(nc, 1000.0)
                alpha = 0.2 / 20 * np.log(10)   # 0.2 dB/m in Np/m
                delta = np.full(nc, fus.compute_diffusivity_of_sound(2 * np.pi * f0, 1480.0, alpha))
                m.tag_source_disc((L / 2, L / 2), 0.032)   # H131 aperture radius 32 mm
                mdl = fus.WesterveltSpectral3D(V, c0, rho0, delta, np.full(nc, 3.5), f0,
                                               1000.0 * 1480.0 * 0.2726428, 1480.0)
                cmax, cfl = 1480.0, 0.2
            dt = cfl * np.sqrt(3) * h / (cmax * P * P)
            mdl.init()
            mdl.rk4(0.0, 2.5 * dt, dt)
            torch.cuda.synchronize()
            K = 20
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            done = mdl.rk4(3 * dt, 3 * dt + (K - 0.5) * dt, dt)
            e1.record(stream)
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            u = mdl.u_sol()
            print(json.dumps({"config": "model_rk4", "model": name, "P": P, "dofs": V.ndofs,
                              "steps": done, "ms_per_step": ms / done,
                              "dof_updates_per_s": V.ndofs * done / (ms * 1e-3),
                              "finite": bool(np.isfinite(u).all()),
                              "u_norm": float(np.linalg.norm(u))}), flush=True)
            mdl.destroy()
    if args.fp32:
        sweep_fp32(args, fus, torch, stream, pk)


if __name__ == "__main__":
    main()
