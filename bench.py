#!/usr/bin/env python
"""bench.py -- headline benchmark of the hot path (see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json metric "FP64 DOF-updates/sec (operator+RK4)"): LinearSpectral3D RK4 on the
P=4 GLL hex box of 54^3 cells per GPU (10.2 M dofs; SURVEY.md section 8d config 1 at the config-2
size); one "step" is one RK4 time step = 4 fused stages (operator + epilogue).  For N > 1 the mesh
is the union of Px x Py x Pz such boxes (weak scaling), partitioned one box per GPU; the stage
kernels exchange the interface values themselves over NVLink peer memory (NCCL for set-up).

Prints ONE JSON line on rank 0.  `value` is device-resident throughput, `e2e` the same metric
through the host-facing call sequence init(u,v) -> rk4 -> u_sol() with host buffers and the copies
inside the timed region, `roofline` the dominant kernel (stiffness operator) against the measured
HBM peak, `cpu_baseline` the reference kernels (oracle/_ref) on this box's host cores, `parity` the
relative L2 differences between the GPU path and that CPU run on the SAME mesh, numbering and time
steps (N = 1: the full-size workload; N > 1: a reduced partitioned box against the single-domain
oracle) -- the run fails when they exceed BASELINE.json's tolerances.
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
H_CELL = BOX_LEN / 54      # same cell size whatever the box
PARITY_TOL = {"apply_rel_l2": 1e-12, "u_rel_l2": 1e-10, "v_rel_l2": 1e-10}   # BASELINE.json north_star
PARITY_CELLS_MULTI = 16    # cells per direction per GPU of the partitioned parity run
# vector passes per stage of SURVEY section 8d's byte model: 52 r + this many bytes per dof
STAGE_BYTES = {"linear": 112.0, "linear_het": 112.0, "lossy": 128.0, "westervelt": 136.0}


def timestep(P, h_edge, c, cfl=CFL, freq=FREQ):
    """Time step of every reference driver (BM7-SC1/main.cpp:112-118): CFL on the cell diameter,
    snapped to an integer number of steps per period."""
    dt0 = cfl * (np.sqrt(3.0) * h_edge) / (c * P * P)
    steps_per_period = int((1.0 / freq) / dt0) + 1
    return (1.0 / freq) / steps_per_period


def rel_l2(a, b):
    nb = float(np.linalg.norm(b))
    return float(np.linalg.norm(a - b)) / (nb if nb > 0 else 1.0)


# --------------------------------------------------------------------------------------------
# The BASELINE configs as data: the same description builds the GPU model and the CPU oracle model
# --------------------------------------------------------------------------------------------
def model_params(name, P, cell_x, nx_global):
    """Physics of one BASELINE config on the bench box.  cell_x: global x index of every cell (the
    layered media of config 3 are layers along x by cell index), nx_global: cells along x.
      linear      config 1: water, planar source on x = 0, absorbing x = L (SC2-BM1/main.cpp:32-44)
      linear_het  config 3: water / skin / cortical / trabecular / brain layers
                  (cpp/fenicsx-sf/experiments/measure_vector_assembly_speed/main.cpp:43-83), time step
                  from the fastest medium (BM7-SC1/main.cpp:112-113)
      lossy       Lossy.hpp on the same water box, attenuation 5 Np/m
      westervelt  config 4: HITU water parameters (HITU/W-H131-WATER/main.cpp:33-46), source on the
                  disc of the H131 aperture (radius 32 mm) on x = 0 -- the stand-in for the
                  focused-bowl meshes, which are not distributed
    lossy / westervelt use CFL 0.2: their all-facet absorbing term is unstable at the drivers' 0.65 on
    a box."""
    nc = cell_x.size
    full = lambda v: np.full(nc, float(v))  # noqa: E731
    d = dict(name=name, kind="linear", delta0=None, beta0=None, freq=FREQ, p0=P0, s0=C0, disc=None)
    if name == "linear":
        d.update(c0=full(C0), rho0=full(RHO0), dt=timestep(P, H_CELL, C0))
    elif name == "linear_het":
        lay = np.minimum(cell_x * 5 // nx_global, 4)
        c_tab = np.array([1500.0, 1610.0, 2800.0, 2300.0, 1560.0])
        r_tab = np.array([1000.0, 1090.0, 1850.0, 1700.0, 1040.0])
        d.update(c0=c_tab[lay], rho0=r_tab[lay], dt=timestep(P, H_CELL, 2800.0))
    elif name == "lossy":
        delta = 2 * 5.0 * C0 ** 3 / (2 * np.pi * FREQ) ** 2           # Westervelt.hpp:408-413
        d.update(kind="lossy", c0=full(C0), rho0=full(RHO0), delta0=full(delta),
                 dt=timestep(P, H_CELL, C0, cfl=0.2))
    elif name == "westervelt":
        f0, c, rho = 1.1e6, 1480.0, 1000.0
        alpha = 0.2 / 20 * np.log(10)                                  # 0.2 dB/m in Np/m
        delta = 2 * alpha * c ** 3 / (2 * np.pi * f0) ** 2
        d.update(kind="westervelt", c0=full(c), rho0=full(rho), delta0=full(delta), beta0=full(3.5),
                 freq=f0, p0=rho * c * 0.2726428, s0=c, dt=timestep(P, H_CELL, c, cfl=0.2, freq=f0),
                 disc=0.032)
    else:
        raise SystemExit(f"unknown model {name}")
    return d


def tag_source_disc(x, xdofmap, facets, centre_yz, radius):
    """Facets of the x = lo face (local facet 2, tag 1) whose centroid lies farther than `radius`
    from (y, z) = centre_yz lose the source tag (BoxMesh.tag_source_disc on plain arrays)."""
    f = np.array(facets, dtype=np.int32, copy=True)
    on = np.flatnonzero((f[:, 1] == 2) & (f[:, 2] == 1))
    if on.size:
        cen = x[xdofmap[f[on, 0]][:, (0, 2, 4, 6)]].mean(axis=1)
        far = np.hypot(cen[:, 1] - centre_yz[0], cen[:, 2] - centre_yz[1]) > radius
        f[on[far], 2] = 0
    return np.ascontiguousarray(f)


def gpu_model(fus, prm, V, facets, device):
    kw = dict(facets=facets, device=device)
    if prm["kind"] == "linear":
        return fus.LinearSpectral3D(V, prm["c0"], prm["rho0"], prm["freq"], prm["p0"], prm["s0"], **kw)
    if prm["kind"] == "lossy":
        return fus.LossySpectral3D(V, prm["c0"], prm["rho0"], prm["delta0"], prm["freq"], prm["p0"],
                                   prm["s0"], **kw)
    return fus.WesterveltSpectral3D(V, prm["c0"], prm["rho0"], prm["delta0"], prm["beta0"],
                                    prm["freq"], prm["p0"], prm["s0"], **kw)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def library_source_hash():
    """Content hash of the sources libfus_b200.so is built from (fenicsx-fus_b200/build.py)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "fus_b200_build", os.path.join(ROOT, "fenicsx-fus_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod._source_hash()


def measured_traffic(kernel_key):
    """DRAM bytes per launch of the dominant kernel from the ncu capture kept in
    profiles/traffic.json -- only if that capture was taken from THIS build of the library (the file
    records the source hash); a capture of another build is not a measurement of this one."""
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(tpath) as f:
            t = json.load(f)
        if t.get("source_hash") != library_source_hash():
            return None, "profiles/traffic.json is from another build of the library"
        return t.get(kernel_key, {}).get("dram_bytes_per_launch"), t.get("captured_with", "")
    except Exception as ex:
        return None, repr(ex)[:120]


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region, polled through NVML every ~2 ms from
    a thread (nvidia-smi's own loop is too coarse for a 30 ms region); falls back to one
    nvidia-smi query when NVML is unavailable."""
    HW, HW_THERM, SW_THERM, SW_POWER = 0x8, 0x40, 0x20, 0x4

    def __init__(self, index=0):
        self.index, self.samples, self.reason_bits = index, [], 0
        self.stop_flag = threading.Event()
        self.thread, self.h, self.nv, self.max_mhz = None, None, None, None
        self.t_begin, self.t_end = None, None     # the timed region, on the clock of the samples

    def mark_begin(self):
        self.t_begin = time.perf_counter()

    def mark_end(self):
        self.t_end = time.perf_counter()

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
                t0 = time.perf_counter()
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.samples.append((t0, time.perf_counter(), mhz, bits))
                self.reason_bits |= bits
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
        # An NVML query takes milliseconds under load, so a 24 ms timed region sees one or two of
        # them: the poller also runs through the warm-up before it and the two passes after it that
        # repeat the same K steps (per-kernel events, end to end).  Reported: the samples that
        # overlap the timed region (`samples`, and their median when there are any) and all samples
        # under this workload; throttle reasons are the union over all of them (conservative).
        def overlaps(smp):
            return (self.t_begin is not None and self.t_end is not None
                    and smp[1] >= self.t_begin and smp[0] <= self.t_end)
        timed = [smp for smp in self.samples if overlaps(smp)]
        use = timed if timed else self.samples
        reasons = [nm for nm, bit in (("hw_slowdown", self.HW), ("hw_thermal_slowdown", self.HW_THERM),
                                      ("sw_thermal_slowdown", self.SW_THERM),
                                      ("sw_power_cap", self.SW_POWER)) if self.reason_bits & bit]
        allmhz = [smp[2] for smp in self.samples]
        return {"sm_mhz": float(np.median([smp[2] for smp in use])) if use else None,
                "sm_max_mhz": self.max_mhz, "samples": len(timed), "reasons": reasons,
                "samples_same_workload": len(allmhz),
                "sm_mhz_min_same_workload": float(min(allmhz)) if allmhz else None,
                "source": ("NVML polled during the timed region (`samples`) and, for more samples, "
                           "through the warm-up before it and the two repeat passes of the same K "
                           "steps after it; reasons = union over all of them")}


# --------------------------------------------------------------------------------------------
# CPU side: the reference kernels on the host cores (oracle/_ref)
# --------------------------------------------------------------------------------------------
def cpu_model_rk4(P, steps, warmup, prm=None, mesh=None, n_cells=None, threads=None, budget_s=None,
                  fields=False):
    """One of the models' RK4 loops with the cell loops running on the reference's own
    sum_factorisation.hpp (oracle/_ref), one OpenMP thread per contiguous cell range.
    mesh = (x, xdofmap, dofmap, facets, ndofs): run on exactly these arrays (the GPU arm's own mesh
    and numbering: the parity check); else on the oracle's box of n_cells^3 cells.  prm: a
    model_params dict whose per-cell arrays match the mesh (default: the headline linear model).
    Returns dict(value, cores, steps, ndofs, kind, seconds[, u, v, apply])."""
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
    if mesh is None:
        n = (n_cells,) * 3
        xg, xd = orc.box_mesh(n, (0, 0, 0), (H_CELL * n_cells,) * 3)
        dm = orc.box_dofmap(P, n, 0)
        nd = int(dm.max()) + 1
        facets = orc.box_facets(n)
    else:
        xg, xd, dm, facets, nd = mesh
    nc = dm.shape[0]
    if prm is None:
        prm = model_params("linear", P, np.zeros(nc, dtype=np.int64), 1)
    G, dJ = orc.geometry(P, xg, xd)
    fn, fs = orc.facet_data(P, xg, xd, facets)
    dphi = orc.dphi(P)
    mdl = orc.model(prm["kind"], P, nd, dm, G, dJ, dphi, prm["c0"], prm["rho0"], prm["delta0"],
                    prm["beta0"], facets, fn, fs, prm["freq"], prm["p0"], prm["s0"],
                    use_ref_kernels=use_ref)
    dt = prm["dt"]
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
    res = dict(value=nd * done / el, cores=cores, steps=done, ndofs=nd, kind=kind, seconds=el)
    if fields:
        res.update(u=u, v=v, apply=lambda x, coeffs: orc.stiffness_apply(
            P, dm, G, dphi, coeffs, x, np.zeros(nd), use_ref_kernels=use_ref))
    return res


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
        probe = cpu_model_rk4(P_BENCH, 2, 1, n_cells=18)
        est = (N_BENCH * P_BENCH + 1) ** 3 * (args.steps + args.warmup) / probe["value"]
        if est > 150.0:
            n_cells = 27
    r = cpu_model_rk4(P_BENCH, args.steps, args.warmup, n_cells=n_cells)
    val, done, nd = r["value"], r["steps"], r["ndofs"]
    sample = (f"LinearSpectral3D RK4, P={P_BENCH}, box {n_cells}^3 cells ({nd} dofs), {done} steps, "
              f"{r['cores']} OpenMP threads as ranks, cell loops on the reference's "
              "sum_factorisation.hpp")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": done, "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / max(done, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"linear_rk4_P{P_BENCH}_box{N_BENCH}", "degree": P_BENCH,
                   "cells_per_direction": n_cells, "dofs": nd, "full_size": n_cells == N_BENCH},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                         "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------
# Child processes: secondary measurements that must never cost the headline line
# --------------------------------------------------------------------------------------------
def _child_env():
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "LOCAL_WORLD_SIZE", "GROUP_RANK", "ROLE_RANK",
              "MASTER_ADDR", "MASTER_PORT", "TORCHELASTIC_RUN_ID", "TORCHELASTIC_RESTART_COUNT",
              "TORCHELASTIC_MAX_RESTARTS", "TORCHELASTIC_USE_AGENT_STORE", "GROUP_WORLD_SIZE",
              "ROLE_WORLD_SIZE", "ROLE_NAME", "TORCHELASTIC_ERROR_FILE"):
        env.pop(k, None)
    return env


def _json_rows(out):
    rows = []
    for ln in out.splitlines():
        if ln.startswith("{"):
            try:
                rows.append(json.loads(ln))
            except ValueError:
                pass
    return rows


def _run_child(cmd, timeout_s):
    """(JSON rows printed by the child, exit code, stderr tail); the child gets its own process
    group so that a time-out takes its workers with it."""
    try:
        proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                                env=_child_env(), start_new_session=True)
    except OSError as ex:
        return [], -1, repr(ex)[:200]
    try:
        out, err = proc.communicate(timeout=timeout_s)
        return _json_rows(out), proc.returncode, err[-400:]
    except subprocess.TimeoutExpired:
        import signal
        try:
            os.killpg(proc.pid, signal.SIGKILL)
        except OSError:
            pass
        out, err = proc.communicate()
        return _json_rows(out or ""), -9, f"timed out after {timeout_s:.0f} s"


def _run_sweep(extra_args, timeout_s):
    return _run_child([sys.executable, os.path.join(ROOT, "scripts", "bench_sweep.py")] + extra_args,
                      timeout_s)


def child_extras(timeout_s=150.0):
    """scripts/bench_sweep.py in a child process: the degree sweep of BASELINE config 2 (single
    operator application, P=2..7, ~10 M dofs) with the library's kernel choice per degree, the same
    with the geometric factors compressed / rebuilt on the fly, the headline RK4 workload per
    geometry mode, the FP32 operators."""
    t0 = time.perf_counter()
    rows, rc, err = _run_sweep(
        ["--degrees", "2,3,4,5,6,7", "--variants=-1", "--geometry-modes", "0,1,2",
         "--rk4-geometry-modes", "0,1,2", "--models", "", "--repeats", "20", "--fp32"], timeout_s)
    keep = ("P", "dofs", "geometry_mode", "variant", "ms_min", "ms_median", "gdof_per_s",
            "frac_of_measured_peak", "ms_per_step", "dof_updates_per_s", "operator_ms",
            "rel_l2_vs_first_mode", "rel_l2_vs_first_config", "rel_l2_vs_fp64")

    def pick(rws, config):
        return [{k: r[k] for k in keep if k in r} for r in rws if r.get("config") == config]

    res = {"exit": rc,
           "degree_sweep_operator_apply": pick(rows, "degree_sweep"),
           "headline_rk4_by_geometry_mode": pick(rows, "headline_rk4_by_geometry_mode"),
           "degree_sweep_operator_apply_fp32": pick(rows, "degree_sweep_fp32"),
           "note": ("variant -1 = the kernel the library picks for the degree (fus_capi.cu, from the "
                    "hardware sweep in profiles/r2a_variant_sweep.jsonl); geometry_mode 0 streams the "
                    "reference's G (48 B/point; the roofline's bytes), 1 keeps one Ghat per affine "
                    "cell (the box qualifies), 2 rebuilds G per point from the trilinear cell map "
                    "(192 B/cell); frac_of_measured_peak always uses the streamed algorithmic bytes")}
    if rc != 0:
        res["stderr_tail"] = err
    res["wall_s"] = time.perf_counter() - t0
    return res


MODEL_KEYS = ("value", "ms_per_step", "n_gpus", "steps", "config", "roofline", "parity", "cpu_baseline",
              "e2e", "vs_baseline")


EXTRAS_BUDGET_S = 480.0   # all child runs of config_extras together


def _bench_child(gpus, extra, timeout_s, port):
    """bench.py itself as a child (other models / degrees / sizes): its JSON line, trimmed."""
    if gpus > 1:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               f"--nproc-per-node={gpus}", "--master-addr", "127.0.0.1", "--master-port", str(port),
               os.path.join(ROOT, "bench.py"), "--gpus", str(gpus)]
    else:
        cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", "1"]
    t0 = time.perf_counter()
    rows, rc, err = _run_child(cmd + ["--no-extras"] + extra, timeout_s)
    rows = [r for r in rows if "metric" in r]
    if rc != 0 or not rows:
        return {"exit": rc, "stderr_tail": err, "wall_s": time.perf_counter() - t0}
    d = {k: rows[-1][k] for k in MODEL_KEYS if k in rows[-1]}
    if "roofline" in d:
        d["roofline"] = {k: d["roofline"].get(k) for k in (
            "kernel", "achieved", "peak", "frac", "avg_launch_ms", "stage_epilogue_avg_ms",
            "step_algorithmic_gbs", "step_frac", "algorithmic_bytes_per_dof_per_stage")}
    d["wall_s"] = time.perf_counter() - t0
    return d


def config_extras(world, steps, warmup, port):
    """BASELINE configs 3, 4, 5 next to the headline, each a child run of this script with its own
    byte model (52 r + 112 / 128 / 136 B per dof and stage), its own parity block and CPU baseline:
      N = 1: heterogeneous linear (config 3), lossy, Westervelt (config 4) at the headline size;
      N > 1: Westervelt on N GPUs (config 4 is quoted at 1/2/4/8);
      N = 8: config 5 -- P=5, 100^3 cells per GPU (1.003 G dofs), 20 steps -- and the same box on one
             GPU, so that the line carries its own weak-scaling efficiency."""
    out = {}
    sw = ["--steps", str(steps), "--warmup", str(max(3, warmup))]
    # The children take 12-50 s each on a B200.  All of them together get EXTRAS_BUDGET_S: a child
    # that hangs must not push the headline line past the caller's own time limit.
    deadline = time.perf_counter() + EXTRAS_BUDGET_S

    def child(gpus, extra, timeout_s, child_port):
        left = deadline - time.perf_counter()
        if left < 30.0:
            return {"skipped": "the time budget of the child runs is spent"}
        return _bench_child(gpus, extra, min(timeout_s, left), child_port)

    if world == 1:
        for name in ("linear_het", "lossy", "westervelt"):
            out[name] = child(1, ["--model", name] + sw, 150.0, port)
    else:
        out["westervelt"] = child(world, ["--model", "westervelt", "--no-cpu-baseline"] + sw, 200.0,
                                  port)
    if world == 8 and os.environ.get("FUS_BENCH_CONFIG5", "1") != "0":
        c5 = ["--degree", "5", "--cells", "100", "--steps", "20", "--warmup", "3",
              "--no-cpu-baseline", "--no-parity"]
        out["config5_P5_box100_8gpu"] = r8 = child(8, c5, 300.0, port + 1)
        out["config5_P5_box100_1gpu"] = r1 = child(1, c5, 240.0, port + 2)
        if "ms_per_step" in r8 and "ms_per_step" in r1:
            out["config5_weak_scaling_efficiency"] = r1["ms_per_step"] / r8["ms_per_step"]
    return out


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def build_problem(fus, partition, dist, args, P, n_cells, pg, rank, local_rank, world, stream=None,
                  peer=True):
    """Partition of the global box, context, halo, model of the requested config on this rank."""
    n_global = tuple(n_cells * p for p in pg)
    part = partition.BoxPartition(P, n_global, pg, rank, lo=(0.0, 0.0, 0.0),
                                  hi=tuple(H_CELL * n for n in n_global))
    V = part.function_space(device=local_rank, lean=args.lean)
    ctx = V.context(local_rank)
    if stream is not None:
        ctx.set_stream(stream.cuda_stream)
    transport = "none"
    if world > 1:
        part.setup_halo(ctx, dist)
        transport = "nccl send/recv"
        if peer and os.environ.get("FUS_HALO_TRANSPORT", "peer") == "peer" and part.connect_peers(ctx, dist):
            transport = ("fused peer transport: the stage kernels exchange over NVLink peer memory "
                         "(CUDA IPC); NCCL for set-up reductions")
    cell_x = part.cell_global // (n_global[1] * n_global[2])
    prm = model_params(args.model, P, cell_x, n_global[0])
    facets = part.facets
    if prm["disc"]:
        facets = tag_source_disc(part.x, part.xdofmap, facets,
                                 (0.5 * H_CELL * n_global[1], 0.5 * H_CELL * n_global[2]), prm["disc"])
    if args.geometry_mode and not args.lean:
        ctx.set_option("geometry_mode", args.geometry_mode)
    mdl = gpu_model(fus, prm, V, facets, local_rank)
    return part, V, ctx, mdl, prm, facets, transport, n_global


def oracle_global(args, P, n_global, steps):
    """Single-domain CPU run of the same config on the global box, dof id == global node key."""
    from oracle.oracle import Oracle
    orc = Oracle()
    xg, xd = orc.box_mesh(n_global, (0, 0, 0), tuple(H_CELL * n for n in n_global))
    dm = orc.box_dofmap(P, n_global, 0)
    nd = int(dm.max()) + 1
    facets = orc.box_facets(n_global)
    nc = dm.shape[0]
    cell_x = np.arange(nc) // (n_global[1] * n_global[2])
    prm = model_params(args.model, P, cell_x, n_global[0])
    if prm["disc"]:
        facets = tag_source_disc(xg, xd, facets, (0.5 * H_CELL * n_global[1], 0.5 * H_CELL * n_global[2]),
                                 prm["disc"])
    return cpu_model_rk4(P, steps, 0, prm=prm, mesh=(xg, xd, dm, facets, nd), fields=True)


def partition_parity(fus, partition, dist, torch, args, P, pg, rank, local_rank, world, steps=10):
    """N > 1: the partitioned path (halo exchange included) on a reduced global box of
    PARITY_CELLS_MULTI^3 cells per GPU, `steps` RK4 steps from rest, owned dofs gathered to rank 0 and
    compared with the single-domain oracle.  Run before the timed passes; every SCALE line carries it."""
    part, V, ctx, mdl, prm, _, transport, n_global = build_problem(
        fus, partition, dist, args, P, PARITY_CELLS_MULTI, pg, rank, local_rank, world)
    mdl.init()
    done = mdl.rk4(0.0, (steps - 0.5) * prm["dt"], prm["dt"])
    u, v = mdl.u_sol(), mdl.v_sol()
    mdl.destroy()
    ctx.destroy()
    V._ctx = None
    gathered = [None] * world
    dist.gather_object((part.global_key[:part.nowned], u[:part.nowned], v[:part.nowned], done),
                       gathered if rank == 0 else None, dst=0)
    if rank != 0:
        return None
    ref = oracle_global(args, P, n_global, steps)
    gu, gv = np.zeros(ref["ndofs"]), np.zeros(ref["ndofs"])
    for keys, uu, vv, dn in gathered:
        assert dn == ref["steps"], (dn, ref["steps"])
        gu[keys], gv[keys] = uu, vv
    return {"apply_rel_l2": None, "u_rel_l2": rel_l2(gu, ref["u"]), "v_rel_l2": rel_l2(gv, ref["v"]),
            "steps": int(ref["steps"]), "against": "single-domain " + ("oracle/_ref" if ref["kind"] ==
                                                                        "reference" else "oracle"),
            "workload": (f"{args.model} RK4 from rest, P={P}, global box {n_global[0]}x{n_global[1]}x"
                         f"{n_global[2]} cells ({ref['ndofs']} dofs) on {pg[0]}x{pg[1]}x{pg[2]} GPUs, "
                         f"{transport.split(':')[0]}"),
            "tolerance": PARITY_TOL}


def single_parity(fus, args, P, part, V, mdl, prm, facets, steps):
    """N = 1: the full-size workload itself.  K RK4 steps from rest on the GPU and on the reference
    kernels (oracle/_ref, all host cores) on the SAME mesh arrays, numbering and time steps; one
    operator application on a seeded random vector.  The CPU run is timed: it is the cpu_baseline."""
    mdl.init()
    done = mdl.rk4(0.0, (steps - 0.5) * prm["dt"], prm["dt"])
    u, v = mdl.u_sol(), mdl.v_sol()
    ref = cpu_model_rk4(P, steps, 0, prm=prm,
                        mesh=(part.x, part.xdofmap, part.dofmap, facets, V.ndofs), fields=True)
    assert done == ref["steps"], (done, ref["steps"])
    x = np.random.default_rng(12345).uniform(-1.0, 1.0, V.ndofs)     # SURVEY section 8d seed
    coeffs = -1.0 / prm["rho0"]
    y = fus.StiffnessSpectral3D(V)(x, coeffs, np.zeros(V.ndofs))
    par = {"apply_rel_l2": rel_l2(y, ref["apply"](x, coeffs)), "u_rel_l2": rel_l2(u, ref["u"]),
           "v_rel_l2": rel_l2(v, ref["v"]), "steps": int(done),
           "against": "oracle/_ref (reference kernels)" if ref["kind"] == "reference" else "oracle port",
           "workload": f"{args.model} RK4 from rest, P={P}, box {N_BENCH}^3 cells ({V.ndofs} dofs): "
                       "the benchmark workload itself", "tolerance": PARITY_TOL}
    cpu = {"value": ref["value"], "unit": UNIT, "cores": ref["cores"], "kind": ref["kind"],
           "sample": (f"the same {args.model} model on the same P={P} box of {N_BENCH}^3 cells "
                      f"({V.ndofs} dofs), {ref['steps']} steps in {ref['seconds']:.1f} s, "
                      f"{ref['cores']} OpenMP threads as ranks")}
    return par, cpu


def parity_ok(par):
    if not par:
        return True
    return all(par.get(k) is None or par[k] <= tol for k, tol in PARITY_TOL.items())


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
    K, W = args.steps, max(args.warmup, 0)

    # ---- N > 1: parity of the partitioned path on a reduced box, before anything is timed ----
    parity = None
    if world > 1 and not args.no_parity:
        parity = partition_parity(fus, partition, dist, torch, args, P, pg, rank, local_rank, world)
        torch.cuda.empty_cache()

    stream = torch.cuda.Stream()         # not the legacy default stream: it cannot be captured
    torch.cuda.set_stream(stream)
    part, V, ctx, mdl, prm, facets, transport, n_global = build_problem(
        fus, partition, dist, args, P, N_BENCH, pg, rank, local_rank, world, stream=stream)
    dt = prm["dt"]
    stage_vector_bytes = STAGE_BYTES[args.model]
    geometry = {0: "G streamed, 48 B/point (the reference's data)",
                1: "affine cells: Ghat per cell", 2: "rebuilt per point from the trilinear cell map",
                3: "rebuilt per point from the trilinear cell map (128-register build)"
                }[ctx.get_option("geometry_compressed")] + (" (lean context: no G/detJ stored)"
                                                            if args.lean else "")
    gmode_used = ctx.get_option("geometry_compressed")
    ndofs_global = part.ndofs_global

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    ctx_uses_graph = os.environ.get("FUS_USE_GRAPH", "1") != "0" and (
        world == 1 or transport.startswith("fused"))
    # ---- device-resident throughput ------------------------------------------------------
    mdl.init()
    t = 0.0
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()          # NVML set-up happens here, not between the barrier and the timed call
    if W:
        # W warm-up steps: the first call issues step 0 eagerly and captures the step graph, the
        # second one replays it from its first step on, as the timed call does
        w1 = max(W // 2, min(W, 3))
        assert mdl.rk4(t, t + (w1 - 0.5) * dt, dt) == w1
        t += w1 * dt
        if W > w1:
            assert mdl.rk4(t, t + (W - w1 - 0.5) * dt, dt) == W - w1
            t += (W - w1) * dt
    sync_all()
    launches0 = fus.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.mark_begin()
    ev0.record(stream)
    t_enq0 = time.perf_counter()
    done = mdl.rk4(t, t + (K - 0.5) * dt, dt)
    host_enqueue_ms = 1e3 * (time.perf_counter() - t_enq0)
    ev1.record(stream)
    sync_all()
    sampler.mark_end()
    launches = fus.launch_count() - launches0
    assert done == K, (done, K)
    t += K * dt
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    per_rank_ms = [float(ms.item()) / K]
    if world > 1:
        allms = [torch.zeros_like(ms) for _ in range(world)]
        dist.all_gather(allms, ms)
        per_rank_ms = [float(a.item()) / K for a in allms]
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
    clocks = sampler.stop() if rank == 0 else None

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
    if headline_geometry and world == 1 and not args.no_extras:
        ctx.set_option("geometry_mode", 1)
        if ctx.get_option("geometry_compressed") == 1:
            mdl.rk4(t, t + 2.5 * dt, dt)
            sync_all()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record(stream)
            mdl.rk4(t, t + (K - 0.5) * dt, dt)
            a1.record(stream)
            sync_all()
            msa = a0.elapsed_time(a1)
            ctx.set_option("profile_kernels", 1)
            mdl.rk4(t, t + (K - 0.5) * dt, dt)
            sync_all()
            ctx.set_option("profile_kernels", 0)
            n_a, ms_a = ctx.profile("stiffness")
            extras["affine_compressed_geometry"] = {
                "value": ndofs_global * K / (msa * 1e-3), "unit": UNIT,
                "ms_per_step": msa / K, "operator_ms": ms_a / (4 * K),
                "note": ("option geometry_mode=1: all cells of the box are parallelepipeds, G = w_q*Ghat "
                         "is rebuilt from 6 numbers per cell instead of streamed (48 B/point); not the "
                         "headline because it depends on the mesh")}
        ctx.set_option("geometry_mode", 0)

    # ---- N = 1: parity at the full size + the CPU baseline, one oracle run for both -------------
    cpu = None
    if world == 1 and not args.no_parity and not args.lean:
        try:
            parity, cpu = single_parity(fus, args, P, part, V, mdl, prm, facets, min(K, args.parity_steps))
        except Exception as ex:  # reported, and the run fails below: parity is the first gate
            parity = {"error": repr(ex)[:300], "apply_rel_l2": float("inf")}
        if args.no_cpu_baseline:
            cpu = None

    if rank != 0:
        mdl.destroy()
        ctx.destroy()
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (stiffness operator) -------------------------------
    peak, peak_src = measured_peaks()
    npts_loc = part.ncells * (P + 1) ** 3
    # 48 B G + 4 B dofmap per point; x read, y write (+ 8 B per dof for the second gathered vector)
    alg_bytes = 52.0 * npts_loc + (16.0 if prm["kind"] == "linear" else 24.0) * nloc
    # one operator application per stage; normalise by stages rather than by launches
    n_apply = 4 * K
    avg_ms = ms_st / n_apply
    achieved = alg_bytes / (avg_ms * 1e-3) / 1e9
    fuse2 = "false" if prm["kind"] == "linear" else "true"
    traffic, traffic_note = (None, "single GPU, streamed geometry only")
    if world == 1 and gmode_used == 0:
        traffic, traffic_note = measured_traffic(f"stiffness_line_kernel<{P + 1},{fuse2}>@P{P}_box{N_BENCH}")
    # whole-step algorithmic bytes: 4 * (52 r + 112) per dof for the linear model,
    # + 16 (lossy: second gathered vector) or + 24 (Westervelt, fused-minimal flow)
    step_bytes = 4.0 * (52.0 * npts_loc + stage_vector_bytes * nloc)
    step_gbs = step_bytes * K / (ms_total * 1e-3) / 1e9

    mdl.destroy()
    ctx.destroy()
    V._ctx = None
    torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()

    # ---- children (after every other rank has left; this process holds no device memory) -----
    if not args.no_extras and (P, N_BENCH) == (4, 54) and headline_geometry and args.model == "linear":
        port = int(os.environ.get("MASTER_PORT", "29500")) + 11
        if world == 1:
            try:
                extras["child_process_sweep"] = child_extras()
            except Exception as ex:  # never at the expense of the headline line
                extras["child_process_sweep"] = {"error": repr(ex)[:300]}
        try:
            extras["baseline_configs"] = config_extras(world, K, W, port)
        except Exception as ex:
            extras["baseline_configs"] = {"error": repr(ex)[:300]}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak",
        # the only published time-loop figure is the Lossy solver at P=4: 122 M
        # DOF-updates/s on 76 Ice Lake ranks; the headline (linear) workload has none
        "vs_baseline": (value / 122.0e6 if (args.model, P) == ("lossy", 4) else None),
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.model}_rk4_P{P}_box{N_BENCH}", "degree": P,
                   "cells_per_direction": N_BENCH, "cells_per_gpu": part.ncells,
                   "dofs": ndofs_global, "dofs_per_gpu": int(part.nowned),
                   "process_grid": list(pg), "dt": dt,
                   "l2": f"inputs_exceed_l2 ({48e-6 * npts_loc:.0f} MB of geometric factors streamed per stage)",
                   "parallelism": f"mesh partition {pg[0]}x{pg[1]}x{pg[2]}", "halo": transport,
                   "geometry": geometry,
                   "issue": "one captured CUDA graph per RK4 step" if ctx_uses_graph else "eager"},
        "clocks": clocks,
        "parity": parity,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": bytes_state / K,
                "d2h_bytes_per_step": bytes_state / K,
                "note": ("init(u,v) from pinned host + rk4(K steps) + u_sol()/v_sol() to host, all "
                         "inside the timed region; the state stays resident across steps as in the "
                         "reference's rk4 loop, so its copies are amortised over K steps -- "
                         "roundtrip_every_step_value copies the state in and out around EVERY step"),
                "wall_s": wall_e2e,
                "roundtrip_every_step_value": roundtrip_value},
        "gpu_launches": int(launches), "host_enqueue_ms_per_step": host_enqueue_ms / K,
        "ms_per_step_by_rank": per_rank_ms,
        "roofline": {"bound": "hbm",
                     "kernel": f"stiffness_line_kernel<{P + 1},{fuse2},{gmode_used}>",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src,
                     "launches": int(n_st),
                     "operator_applications": n_apply, "avg_launch_ms": avg_ms,
                     "algorithmic_bytes_per_launch": alg_bytes,
                     "algorithmic_bytes_per_dof_per_stage": 52.0 * npts_loc / nloc + stage_vector_bytes,
                     "stage_epilogue_avg_ms": ms_ep / max(n_ep, 1),
                     "kernel_timing": "CUDA event pairs around every launch in a second pass of the "
                                      "same K steps (eager issue)",
                     "step_algorithmic_gbs": step_gbs, "step_frac": step_gbs / peak,
                     "step_note": ("step_frac uses SURVEY section 8d's byte model (14 vector passes per "
                                   "stage); the fused flow moves fewer (operator 3 + epilogue 8.25 on "
                                   "average), so it can exceed 1")},
        "cpu_baseline": cpu,
        "extras": extras,
    }
    print(json.dumps(line), flush=True)
    if not parity_ok(parity):
        sys.stderr.write(f"bench.py: PARITY FAILED {json.dumps(parity)}\n")
        return 3
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true",
                    help="skip the oracle comparison (sizes the oracle cannot hold)")
    ap.add_argument("--parity-steps", type=int, default=20)
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the child processes (sweeps, other BASELINE configs)")
    # non-headline workloads (the driver never passes these; bench.py's own children do)
    ap.add_argument("--degree", type=int, default=P_BENCH)
    ap.add_argument("--cells", type=int, default=N_BENCH, help="cells per direction per GPU")
    ap.add_argument("--pgrid", default="", help="process grid px,py,pz (default 1x1x1, 2x1x1, "
                    "2x2x1, 2x2x2); our own scaling studies only")
    ap.add_argument("--model", default="linear", choices=sorted(STAGE_BYTES),
                    help="linear_het / lossy / westervelt: BASELINE configs 3-4 (not the headline)")
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
