#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json metric "FP64 DOF-updates/sec (operator+RK4)"): LinearSpectral3D RK4 on the
P=4 GLL hex box of 54^3 cells per GPU (10.2 M dofs; SURVEY.md section 8d config 1 at the config-2
size); one "step" is one RK4 time step = 4 fused stages (operator + epilogue).  For N > 1 the mesh
is the union of Px x Py x Pz such boxes (weak scaling), partitioned one box per GPU with an NCCL
halo exchange per stage.

Prints ONE JSON line on rank 0.  `value` is device-resident throughput, `e2e` the same metric
through the host-facing call sequence init(u,v) -> rk4 -> u_sol() with host buffers and the copies
inside the timed region, `roofline` the dominant kernel (stiffness operator) against the measured
HBM peak, `cpu_baseline` the reference kernels (oracle/_ref) on this box's host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

P_BENCH = 4
N_BENCH = 54               # cells per direction per GPU
BOX_LEN = 0.12             # m (SC2-BM1/main.cpp:41)
C0, RHO0 = 1500.0, 1000.0  # water (SC2-BM1/main.cpp:37-38)
FREQ, P0 = 0.5e6, 60000.0
CFL = 0.65
METRIC = "FP64 DOF-updates/sec (operator+RK4)"
UNIT = "DOF-updates/s"
PGRID = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def timestep(P, h_edge, c, cfl=CFL, freq=FREQ):
    """Time step of every reference driver (BM7-SC1/main.cpp:112-118): CFL on the cell diameter,
    snapped to an integer number of steps per period."""
    dt0 = cfl * (np.sqrt(3.0) * h_edge) / (c * P * P)
    steps_per_period = int((1.0 / freq) / dt0) + 1
    return (1.0 / freq) / steps_per_period


def make_model(fus, name, V, facets, device):
    """The solver of the requested BASELINE config on the bench box: (model, dt, vector passes per
    stage in the SURVEY section 8d byte model).  `linear` is the headline (config 1 physics);
    `lossy` adds attenuation to it (Lossy.hpp); `westervelt` uses the HITU water parameters of
    config 4 (W-H131-WATER/main.cpp:32-46) with a planar source on x = 0 -- the bowl meshes are not
    distributed -- and, like `lossy`, the smaller CFL the all-facet absorbing term needs on a box
    (DESIGN.md section 6)."""
    h = BOX_LEN / 54
    if name == "linear":
        return (fus.LinearSpectral3D(V, C0, RHO0, FREQ, P0, C0, facets=facets, device=device),
                timestep(V.P, h, C0), 112.0)
    if name == "lossy":
        delta = fus.compute_diffusivity_of_sound(2 * np.pi * FREQ, C0, 5.0)
        return (fus.LossySpectral3D(V, C0, RHO0, delta, FREQ, P0, C0, facets=facets, device=device),
                timestep(V.P, h, C0, cfl=0.2), 128.0)
    f0, c, rho = 1.1e6, 1480.0, 1000.0
    delta = fus.compute_diffusivity_of_sound(2 * np.pi * f0, c, 0.2 / 20 * np.log(10))
    return (fus.WesterveltSpectral3D(V, c, rho, delta, 3.5, f0, rho * c * 0.2726428, c,
                                     facets=facets, device=device),
            timestep(V.P, h, c, cfl=0.2, freq=f0), 136.0)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region, polled through NVML every ~2 ms from
    a thread (nvidia-smi's own loop is too coarse for a 30 ms region); falls back to one
    nvidia-smi query when NVML is unavailable."""
    HW, HW_THERM, SW_THERM, SW_POWER = 0x8, 0x40, 0x20, 0x4

    def __init__(self, index=0):
        self.index, self.samples, self.reason_bits = index, [], 0
        self.stop_flag = threading.Event()
        self.thread, self.h, self.nv, self.max_mhz = None, None, None, None

    def start(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            # honour CUDA_VISIBLE_DEVICES-less boxes: local index == NVML index on the gpurun pods
            self.h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            self.nv = nv
        except Exception:
            self.nv = None
            return
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def _poll(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.reason_bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        if self.nv is None:
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--id={self.index}",
                     "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits"],
                    capture_output=True, text=True, timeout=10).stdout.strip().split(",")
                return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "samples": 1,
                        "reasons": [], "source": "nvidia-smi after the timed region (NVML unavailable)"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0,
                        "reasons": ["clock query unavailable"]}
        self.stop_flag.set()
        self.thread.join(timeout=2)
        reasons = [nm for nm, bit in (("hw_slowdown", self.HW), ("hw_thermal_slowdown", self.HW_THERM),
                                      ("sw_thermal_slowdown", self.SW_THERM),
                                      ("sw_power_cap", self.SW_POWER)) if self.reason_bits & bit]
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "samples": len(self.samples), "reasons": reasons,
                "source": "NVML polled every 2 ms during the timed region"}


# --------------------------------------------------------------------------------------------
# CPU side: the reference kernels on the host cores (oracle/_ref)
# --------------------------------------------------------------------------------------------
def cpu_linear_rk4(n_cells, steps, warmup, threads=None, budget_s=None, mesh=None, fields=False):
    """LinearSpectral3D RK4 on a P=4 box of n_cells^3 cells, with the cell loops running on the
    reference's own sum_factorisation.hpp (oracle/_ref), one OpenMP thread per contiguous cell
    range.  `mesh` = (x, xdofmap, dofmap, facets, ndofs) runs the same model on exactly these arrays
    (the GPU arm's own mesh and numbering: the parity check), else on the oracle's box generator.
    Returns (dof_updates_per_s, cores, steps_done, ndofs, kind, seconds[, u, v, apply])."""
    from oracle.oracle import Oracle, ref_available
    use_ref = ref_available() or os.path.isdir("/root/reference/cpp/fenicsx-sf/common")
    orc = Oracle(ref=use_ref)
    kind = "reference" if use_ref else "port"
    cores = os.cpu_count() or 1
    if threads:
        cores = threads
    if use_ref:
        orc.lib.fr_set_threads(cores)
    else:
        cores = 1
    P, n = P_BENCH, (n_cells,) * 3
    h = BOX_LEN / 54                           # same cell size as the GPU workload
    if mesh is None:
        xg, xd = orc.box_mesh(n, (0, 0, 0), (h * n_cells,) * 3)
        dm = orc.box_dofmap(P, n, 0)
        nd = int(dm.max()) + 1
        facets = orc.box_facets(n)
    else:
        xg, xd, dm, facets, nd = mesh
    G, dJ = orc.geometry(P, xg, xd)
    fn, fs = orc.facet_data(P, xg, xd, facets)
    nc = dm.shape[0]
    dphi = orc.dphi(P)
    mdl = orc.model("linear", P, nd, dm, G, dJ, dphi, np.full(nc, C0), np.full(nc, RHO0),
                    None, None, facets, fn, fs, FREQ, P0, C0, use_ref_kernels=use_ref)
    dt = timestep(P, h, C0)
    u, v = np.zeros(nd), np.zeros(nd)
    t = 0.0
    if warmup:
        mdl.rk4(t, t + (warmup - 0.5) * dt, dt, u, v)
        t += warmup * dt
    t0 = time.perf_counter()
    done = 0
    if budget_s is None:
        done = mdl.rk4(t, t + (steps - 0.5) * dt, dt, u, v)
    else:
        while done < steps and (time.perf_counter() - t0) < budget_s:
            done += mdl.rk4(t, t + 0.5 * dt, dt, u, v)
            t += dt
    el = time.perf_counter() - t0
    assert np.isfinite(u).all()
    if not fields:
        return nd * done / el, cores, done, nd, kind, el

    def apply(x, coeffs):
        return orc.stiffness_apply(P, dm, G, dphi, coeffs, x, np.zeros(nd), use_ref_kernels=use_ref)
    return nd * done / el, cores, done, nd, kind, el, u, v, apply


def rel_l2(a, b):
    nb = float(np.linalg.norm(b))
    return float(np.linalg.norm(a - b)) / (nb if nb > 0 else 1.0)


PARITY_TOL = {"apply_rel_l2": 1e-12, "u_rel_l2": 1e-10, "v_rel_l2": 1e-10}   # BASELINE.json north_star


def run_reference_arm(args):
    """--impl reference: the reference's CPU path for the same metric/config on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # full workload if it fits a few minutes, else a bounded sample of it
    n_cells = N_BENCH
    if os.environ.get("FUS_REF_CELLS"):            # tests: force a small sample
        n_cells = int(os.environ["FUS_REF_CELLS"])
    else:
        probe, cores, _, _, kind, _ = cpu_linear_rk4(18, 2, 1)
        est = (N_BENCH * P_BENCH + 1) ** 3 * (args.steps + args.warmup) / probe
        if est > 150.0:
            n_cells = 27
    val, cores, done, nd, kind, el = cpu_linear_rk4(n_cells, args.steps, args.warmup)
    sample = (f"LinearSpectral3D RK4, P={P_BENCH}, box {n_cells}^3 cells ({nd} dofs), {done} steps, "
              f"{cores} OpenMP threads as ranks, cell loops on the reference's sum_factorisation.hpp")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": done, "warmup": args.warmup, "ms_per_step": 1e3 * el / max(done, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"linear_rk4_P{P_BENCH}_box{N_BENCH}", "degree": P_BENCH,
                   "cells_per_direction": n_cells, "dofs": nd, "full_size": n_cells == N_BENCH},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def _run_sweep(extra_args, timeout_s):
    """One run of scripts/bench_sweep.py in a child process: (JSON rows, exit code, stderr tail)."""
    cmd = [sys.executable, os.path.join(ROOT, "scripts", "bench_sweep.py")] + extra_args
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    try:
        res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout_s, env=env)
        out, rc, err = res.stdout, res.returncode, res.stderr[-400:]
    except subprocess.TimeoutExpired as ex:
        out = ex.stdout.decode() if isinstance(ex.stdout, bytes) else (ex.stdout or "")
        rc, err = -9, f"timed out after {timeout_s:.0f} s"
    rows = []
    for ln in out.splitlines():
        if ln.startswith("{"):
            try:
                rows.append(json.loads(ln))
            except ValueError:
                pass
    return rows, rc, err


def child_extras(timeout_s=150.0):
    """Secondary measurements in CHILD processes (scripts/bench_sweep.py), after the headline
    numbers are in hand: the degree sweep of BASELINE config 2 (single operator application, P=2..7,
    ~10 M dofs) with the geometric factors streamed and rebuilt on the fly, the headline RK4
    workload per geometry mode and per pipeline variant, the FP32 operators.  Children so that
    nothing they do -- a fault in a newer kernel, a time-out -- can cost the headline line; the
    TMA-ring variant, whose mbarrier protocol has its first hardware run here, gets a child of its
    own so that it cannot cost the other extras either."""
    t0 = time.perf_counter()
    rows, rc, err = _run_sweep(
        ["--degrees", "2,3,4,5,6,7", "--variants=-1", "--geometry-modes", "0,1,2,3",
         "--rk4-geometry-modes", "0,1,2,3", "--pipeline-variants", "3,4,5", "--models", "",
         "--repeats", "20", "--fp32"], timeout_s)
    keep = ("P", "dofs", "geometry_mode", "variant", "ms_min", "ms_median", "gdof_per_s",
            "frac_of_measured_peak", "ms_per_step", "dof_updates_per_s", "operator_ms",
            "rel_l2_vs_first_mode", "rel_l2_vs_first_config", "rel_l2_vs_fp64")

    def pick(rws, config):
        return [{k: r[k] for k in keep if k in r} for r in rws if r.get("config") == config]

    res = {"exit": rc,
           "degree_sweep_operator_apply": pick(rows, "degree_sweep"),
           "headline_rk4_by_geometry_mode": pick(rows, "headline_rk4_by_geometry_mode"),
           # stiffness_variant 3..5: the line kernel with the software pipelines that move the
           # scoreboard wait seen in the ncu source view of the default kernel;
           # same results, first hardware timing here, default unchanged until it is in hand
           "headline_rk4_by_pipeline_variant": pick(rows, "headline_rk4_by_pipeline_variant"),
           # FP32 operator instantiation (float data, 28 B/point + 8 B/dof algorithmic): first
           # hardware run of these kernels -- their logic is covered by the host emulation tests
           "degree_sweep_operator_apply_fp32": pick(rows, "degree_sweep_fp32"),
           "note": ("geometry_mode 0 streams the reference's G (48 B/point; the roofline's bytes), "
                    "1 keeps one Ghat per affine cell (the box qualifies), "
                    "2 rebuilds G per point from the trilinear cell map (192 B/cell), 3 is the same "
                    "kernel compiled under a 128-register cap (occupancy experiment); "
                    "frac_of_measured_peak always uses the streamed algorithmic bytes")}
    if rc != 0:
        res["stderr_tail"] = err
    # stiffness_variant 6: G through a TMA bulk-copy ring in shared memory, next to the default kernel
    rows6, rc6, err6 = _run_sweep(
        ["--degrees", "2,3,4,5,6,7", "--variants=-1", "--geometry-modes", "0", "--rk4-geometry-modes",
         "0", "--pipeline-variants", "6", "--models", "", "--repeats", "20"], 90.0)
    res["tma_ring_variant"] = {"exit": rc6,
                               "degree_sweep_operator_apply": pick(rows6, "degree_sweep"),
                               "headline_rk4": (pick(rows6, "headline_rk4_by_geometry_mode")
                                                + pick(rows6, "headline_rk4_by_pipeline_variant")),
                               "note": "variant -1 = the default kernel of the same run; P=7 falls "
                                       "back to variant 5 (the ring does not fit in shared memory)"}
    if rc6 != 0:
        res["tma_ring_variant"]["stderr_tail"] = err6
    res["wall_s"] = time.perf_counter() - t0
    return res


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    import fenicsx_fus_b200 as fus
    from fenicsx_fus_b200 import partition

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N")
    if fus.device_count() < 1:
        raise SystemExit("bench.py: no B200 visible and there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank),
                                timeout=datetime.timedelta(seconds=180))

    P = P_BENCH
    pg = PGRID[world]
    if args.pgrid:                       # e.g. 8,1,1: slabs (2 neighbours per rank) instead of 2x2x2
        pg = tuple(int(v) for v in args.pgrid.split(","))
        if len(pg) != 3 or pg[0] * pg[1] * pg[2] != world:
            raise SystemExit("--pgrid must be three factors of the number of GPUs")
    h = BOX_LEN / 54            # same cell size whatever the box
    n_global = tuple(N_BENCH * p for p in pg)
    # local part of the global box, local dof numbering (owned first, then ghosts), halo lists
    part = partition.BoxPartition(P, n_global, pg, rank, lo=(0.0, 0.0, 0.0),
                                  hi=tuple(h * n for n in n_global))
    V = part.function_space(device=local_rank, lean=args.lean)
    ctx = V.context(local_rank)
    stream = torch.cuda.Stream()         # not the legacy default stream: it cannot be captured
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    transport = "none"
    if world > 1:
        part.setup_halo(ctx, dist)
        transport = "nccl send/recv"
        if os.environ.get("FUS_HALO_TRANSPORT", "peer") == "peer" and part.connect_peers(ctx, dist):
            transport = "peer-direct puts over NVLink (CUDA IPC), NCCL for set-up reductions"
    if args.geometry_mode and not args.lean:
        ctx.set_option("geometry_mode", args.geometry_mode)
    geometry = {0: "G streamed, 48 B/point (the reference's data)",
                1: "affine cells: Ghat per cell", 2: "rebuilt per point from the trilinear cell map",
                3: "rebuilt per point from the trilinear cell map (128-register build)"
                }[ctx.get_option("geometry_compressed")] + (" (lean context: no G/detJ stored)"
                                                            if args.lean else "")
    gmode_used = ctx.get_option("geometry_compressed")
    mdl, dt, stage_vector_bytes = make_model(fus, args.model, V, part.facets, local_rank)
    ndofs_global = part.ndofs_global
    K, W = args.steps, max(args.warmup, 0)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    ctx_uses_graph = os.environ.get("FUS_USE_GRAPH", "1") != "0" and (
        world == 1 or transport.startswith("peer"))
    # ---- device-resident throughput ------------------------------------------------------
    mdl.init()
    t = 0.0
    if W:
        assert mdl.rk4(t, t + (W - 0.5) * dt, dt) == W
        t += W * dt
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = fus.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    t_enq0 = time.perf_counter()
    done = mdl.rk4(t, t + (K - 0.5) * dt, dt)
    host_enqueue_ms = 1e3 * (time.perf_counter() - t_enq0)
    ev1.record(stream)
    sync_all()
    launches = fus.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    assert done == K, (done, K)
    t += K * dt
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = ndofs_global * K / (ms_total * 1e-3)
    # per-kernel device times for the roofline: the same K steps again with a CUDA event pair
    # around every launch (this pass is issued eagerly, the timed one above replays a CUDA graph)
    ctx.set_option("profile_kernels", 1)
    assert mdl.rk4(t, t + (K - 0.5) * dt, dt) == K
    sync_all()
    ctx.set_option("profile_kernels", 0)
    t += K * dt
    n_st, ms_st = ctx.profile("stiffness")
    n_ep, ms_ep = ctx.profile("stage")
    u_probe = mdl.u_sol()
    assert np.isfinite(u_probe).all()
    if rank == 0:                      # the source sits on the x = 0 face, i.e. in rank 0's block
        assert np.abs(u_probe).max() > 0.0

    # ---- end to end through the host-facing calls, host buffers, copies timed --------------
    nloc = V.ndofs
    u_host = torch.zeros(nloc, dtype=torch.float64).pin_memory()
    v_host = torch.zeros(nloc, dtype=torch.float64).pin_memory()
    u_host.numpy()[:] = u_probe
    v_host.numpy()[:] = mdl.v_sol()
    uh, vh = u_host.numpy(), v_host.numpy()
    sync_all()
    t_e0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    mdl.init(uh, vh)                                   # H2D of the state (2 vectors)
    done = mdl.rk4(t, t + (K - 0.5) * dt, dt)          # K steps on the device
    fus.check(mdl.lib.fus_model_get_state(mdl.h, uh.ctypes.data, vh.ctypes.data), "get_state")
    e1.record(stream)                                  # D2H of the result (2 vectors) done
    sync_all()
    wall_e2e = time.perf_counter() - t_e0
    ms2 = torch.tensor([max(e0.elapsed_time(e1), 0.0)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    ms_e2e = max(float(ms2.item()), 0.0)
    e2e_value = ndofs_global * K / (ms_e2e * 1e-3)
    bytes_state = 2 * 8 * nloc

    # ---- step with a host round trip of the state around EVERY step (extra information) ----
    kr = min(K, 5)
    sync_all()
    t_r0 = time.perf_counter()
    for _ in range(kr):
        mdl.init(uh, vh)
        mdl.rk4(t, t + 0.5 * dt, dt)
        fus.check(mdl.lib.fus_model_get_state(mdl.h, uh.ctypes.data, vh.ctypes.data), "get_state")
    sync_all()
    roundtrip_value = ndofs_global * kr / (time.perf_counter() - t_r0)

    # ---- extra (not the headline): opt-in affine compression of the geometric factors -----------
    extras = {}
    headline_geometry = not args.lean and not args.geometry_mode
    if headline_geometry:
        ctx.set_option("geometry_mode", 1)
    if headline_geometry and ctx.get_option("geometry_compressed") == 1:
        mdl.rk4(t, t + 2.5 * dt, dt)
        sync_all()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        mdl.rk4(t, t + (K - 0.5) * dt, dt)
        a1.record(stream)
        sync_all()
        msa = torch.tensor([a0.elapsed_time(a1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(msa, op=dist.ReduceOp.MAX)
        ctx.set_option("profile_kernels", 1)
        mdl.rk4(t, t + (K - 0.5) * dt, dt)
        sync_all()
        ctx.set_option("profile_kernels", 0)
        n_a, ms_a = ctx.profile("stiffness")
        extras["affine_compressed_geometry"] = {
            "value": ndofs_global * K / (float(msa.item()) * 1e-3), "unit": UNIT,
            "ms_per_step": float(msa.item()) / K, "operator_ms": ms_a / (4 * K),
            "note": ("option geometry_mode=1: all cells of the box are parallelepipeds, G = w_q*Ghat is "
                     "rebuilt from 6 numbers per cell instead of streamed (48 B/point); not the "
                     "headline because it depends on the mesh")}
    if headline_geometry:
        ctx.set_option("geometry_mode", 0)

    if rank != 0:
        mdl.destroy()
        ctx.destroy()
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (stiffness operator) -------------------------------
    peak, peak_src = measured_peaks()
    npts_loc = part.ncells * (P + 1) ** 3
    alg_bytes = 52.0 * npts_loc + 16.0 * nloc       # 48 B G + 4 B dofmap per point; x read, y write
    # one operator application per stage; when partitioned it is issued as three launches
    # (interior A, interface, interior B), so normalise by stages rather than by launches
    n_apply = 4 * K
    avg_ms = ms_st / n_apply
    achieved = alg_bytes / (avg_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and world == 1 and gmode_used == 0:
        with open(tpath) as f:
            traffic = json.load(f).get(f"stiffness_line_kernel<{P + 1},false>@P{P}_box{N_BENCH}",
                                       {}).get("dram_bytes_per_launch")
    # whole-step algorithmic bytes: 4 * (52 r + 112) per dof for the linear model,
    # + 16 (lossy: second gathered vector) or + 24 (Westervelt, fused-minimal flow)
    step_bytes = 4.0 * (52.0 * npts_loc + stage_vector_bytes * nloc)
    step_gbs = step_bytes * K / (ms_total * 1e-3) / 1e9

    # ---- CPU baseline on this box's host cores (bounded sample) -----------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline and (P, N_BENCH) == (4, 54) and args.model == "linear":
        try:
            val, cores, sdone, snd, kind, el = cpu_linear_rk4(30, 50, 1, budget_s=12.0)
            cpu = {"value": val, "unit": UNIT, "cores": cores, "kind": kind,
                   "sample": (f"same model on a P=4 box of 30^3 cells ({snd} dofs), {sdone} steps in "
                              f"{el:.1f} s, {cores} OpenMP threads as ranks")}
        except Exception as ex:  # the baseline must never sink the GPU measurement
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable",
                   "sample": repr(ex)[:200]}

    if (world == 1 and not args.no_extras and (P, N_BENCH) == (4, 54) and headline_geometry
            and args.model == "linear"):
        # release this process's device memory first; the child builds its own contexts
        mdl.destroy()
        ctx.destroy()
        torch.cuda.empty_cache()
        try:
            extras["child_process_sweep"] = child_extras()
        except Exception as ex:  # never at the expense of the headline line
            extras["child_process_sweep"] = {"error": repr(ex)[:300]}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak",
        # the only published time-loop figure is the Lossy solver at P=4: 122 M
        # DOF-updates/s on 76 Ice Lake ranks; the headline (linear) workload has none
        "vs_baseline": (value / 122.0e6 if (args.model, P) == ("lossy", 4) else None),
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.model}_rk4_P{P}_box{N_BENCH}_per_gpu", "degree": P,
                   "cells_per_gpu": part.ncells, "dofs_global": ndofs_global,
                   "process_grid": list(pg), "dt": dt,
                   "l2": f"inputs_exceed_l2 ({48e-6 * npts_loc:.0f} MB of geometric factors streamed per stage)",
                   "parallelism": f"mesh partition {pg[0]}x{pg[1]}x{pg[2]}", "halo": transport,
                   "geometry": geometry,
                   "issue": "one captured CUDA graph per RK4 step" if ctx_uses_graph else "eager"},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": bytes_state / K,
                "d2h_bytes_per_step": bytes_state / K,
                "note": ("init(u,v) from pinned host + rk4(K steps) + u_sol()/v_sol() to host, all "
                         "inside the timed region; the state stays resident across steps as in the "
                         "reference's rk4 loop"),
                "wall_s": wall_e2e,
                "roundtrip_every_step_value": roundtrip_value},
        "gpu_launches": int(launches), "host_enqueue_ms_per_step": host_enqueue_ms / K,
        "roofline": {"bound": "hbm",
                     "kernel": f"stiffness_line_kernel<{P + 1},false,{gmode_used}>",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "launches": int(n_st),
                     "operator_applications": n_apply, "avg_launch_ms": avg_ms,
                     "algorithmic_bytes_per_launch": alg_bytes,
                     "stage_epilogue_avg_ms": ms_ep / max(n_ep, 1),
                     "kernel_timing": "CUDA event pairs around every launch in a second pass of the "
                                      "same K steps (eager issue)",
                     "step_algorithmic_gbs": step_gbs, "step_frac": step_gbs / peak},
        "cpu_baseline": cpu,
        "extras": extras,
    }
    print(json.dumps(line), flush=True)
    mdl.destroy()
    ctx.destroy()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the child-process sweep (profiler passes)")
    # non-headline workloads for our own scaling studies (the driver never passes these)
    ap.add_argument("--degree", type=int, default=P_BENCH)
    ap.add_argument("--cells", type=int, default=N_BENCH, help="cells per direction per GPU")
    ap.add_argument("--pgrid", default="", help="process grid px,py,pz (default 1x1x1, 2x1x1, "
                    "2x2x1, 2x2x2); our own scaling studies only")
    ap.add_argument("--model", default="linear", choices=["linear", "lossy", "westervelt"],
                    help="lossy / westervelt: BASELINE configs 3-4 style runs (not the headline)")
    ap.add_argument("--geometry-mode", type=int, default=0, choices=[0, 1, 2, 3],
                    help="1/2: compressed geometric factors (not the headline: see DESIGN.md)")
    ap.add_argument("--lean", action="store_true",
                    help="context without G/detJ on the device (geometry rebuilt on the fly)")
    args = ap.parse_args()
    globals()["P_BENCH"], globals()["N_BENCH"] = args.degree, args.cells
    if args.gpus not in PGRID:
        raise SystemExit("--gpus must be 1, 2, 4 or 8")
    if args.impl == "reference":
        return run_reference_arm(args)
    try:
        return run_gpu_arm(args)
    except BaseException:
        # a failed rank must not leave its peers waiting in a collective (nor hang in teardown)
        import traceback
        traceback.print_exc()
        sys.stderr.flush()
        os._exit(1)


if __name__ == "__main__":
    sys.exit(main())
