"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the
required keys, and the GPU arm refuses to run without a device (no silent CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

REQUIRED = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
            "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "impl"}


def test_reference_arm_json_line():
    env = dict(os.environ, FUS_REF_CELLS="6")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, env=env,
                         timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert REQUIRED <= set(d), REQUIRED - set(d)
    assert d["impl"] == "reference" and d["steps"] == 2 and d["value"] > 0
    assert d["metric"].startswith("FP64 DOF-updates/sec") and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", FUS_REF_CELLS="6")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--gpus", "2", "--steps", "1", "--warmup", "0"], capture_output=True,
                         text=True, env=env, timeout=120)
    assert res.returncode == 0 and not res.stdout.strip()


def test_gpu_arm_has_no_cpu_fallback():
    import fenicsx_fus_b200 as fus
    if fus.device_count() > 0:
        return
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert res.returncode != 0
    assert "no B200 visible" in (res.stderr + res.stdout) or "CUDA" in (res.stderr + res.stdout)


def test_child_extras_never_raise():
    """The secondary sweep runs in a child process; whatever happens there (here: no GPU) comes
    back as data, so the headline line is always printed."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    import fenicsx_fus_b200 as fus
    if fus.device_count() > 0:
        return
    res = bench.child_extras(timeout_s=240.0)
    assert res["exit"] != 0 and res["degree_sweep_operator_apply"] == []
    assert res["headline_rk4_by_geometry_mode"] == [] and "stderr_tail" in res
    short = bench.child_extras(timeout_s=0.05)
    assert short["exit"] == -9 and "timed out" in short["stderr_tail"]
    # the other BASELINE configs are child runs of bench.py itself: same guarantee
    cfg = bench.config_extras(1, 2, 1, 29999)
    assert set(cfg) == {"linear_het", "lossy", "westervelt"} and all(v["exit"] != 0 for v in cfg.values())


class _FakeEvent:
    clock = [0.0]

    def __init__(self, enable_timing=True):
        self.t = None

    def record(self, stream=None):
        _FakeEvent.clock[0] += 1.5
        self.t = _FakeEvent.clock[0]

    def elapsed_time(self, other):
        return other.t - self.t


def _fake_torch():
    """Just enough of torch for bench.py's control flow: streams/events that do nothing and
    tensors backed by numpy."""
    import types

    import numpy as np

    class T:
        is_cuda = True                       # "device" tensors: the emulated device shares host memory

        def __init__(self, a):
            a = np.asarray(a)
            self.a = np.ascontiguousarray(a if a.dtype == np.float32 else a.astype(np.float64))

        @property
        def dtype(self):
            return self.a.dtype

        def data_ptr(self):
            return self.a.ctypes.data

        def pin_memory(self):
            return self

        def numpy(self):
            return self.a

        def item(self):
            return float(self.a.reshape(-1)[0])

        def zero_(self):
            self.a[...] = 0.0
            return self

        def float(self):
            return T(self.a.astype(np.float32))

        def double(self):
            return T(self.a.astype(np.float64))

        def __sub__(self, o):
            return T(self.a - o.a)

        def __truediv__(self, o):
            return T(self.a / np.maximum(np.abs(o.a), 1e-300))

    cuda = types.SimpleNamespace(
        set_device=lambda i: None, synchronize=lambda: None, empty_cache=lambda: None,
        set_stream=lambda s: None, Stream=lambda: types.SimpleNamespace(cuda_stream=0),
        Event=_FakeEvent)
    t = types.ModuleType("torch")
    t.cuda, t.float64 = cuda, "float64"
    t.tensor = lambda v, dtype=None, device=None: T(v)
    t.zeros = lambda n, dtype=None: T(np.zeros(n))
    t.rand = lambda n, dtype=None, device=None: T(np.random.default_rng(0).uniform(size=n))
    t.zeros_like = lambda x: T(np.zeros_like(x.a))
    t.full = lambda shape, v, dtype=None, device=None: T(np.full(shape, v))
    t.linalg = types.SimpleNamespace(vector_norm=lambda x: T([np.linalg.norm(x.a)]))
    t.device = lambda *a: None
    dist = types.ModuleType("torch.distributed")
    t.distributed = dist
    return t, dist


@pytest.mark.parametrize("model,extra", [("linear", []), ("westervelt", ["--geometry-mode", "2"]),
                                         ("lossy", ["--lean"])])
def test_gpu_arm_control_flow_with_a_stub_device(monkeypatch, capsys, model, extra):
    """bench.py's GPU arm cannot run in the build container (no GPU) and there is no CPU fallback to
    run instead, so its Python control flow -- argument plumbing, the JSON line and its keys -- is
    exercised here with the device library and torch stubbed out.  Numbers are meaningless."""
    import importlib.util
    import sys
    import types

    import numpy as np

    import fenicsx_fus_b200 as fus
    spec = importlib.util.spec_from_file_location("bench_mod2", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    t, dist = _fake_torch()
    monkeypatch.setitem(sys.modules, "torch", t)
    monkeypatch.setitem(sys.modules, "torch.distributed", dist)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        monkeypatch.delenv(k, raising=False)
    calls = {"options": [], "lean": None}

    class Ctx:
        def __init__(self):
            self.mode = 0

        def set_stream(self, s):
            pass

        def set_option(self, name, value):
            calls["options"].append((name, value))
            if name == "geometry_mode":
                self.mode = value

        def get_option(self, name):
            return self.mode if name == "geometry_compressed" else 0

        def profile(self, kernel):
            return 8, 2.0

        def destroy(self):
            pass

    def from_mesh(cls, V, device=0, dofmap=None, ndofs=None, nowned=None, lean=False):
        calls["lean"] = lean
        c = Ctx()
        c.mode = 2 if lean else 0
        return c

    class Model:
        def __init__(self, V, *a, **kw):
            assert "facets" in kw and "device" in kw
            self.V, self.h = V, 0
            self.lib = types.SimpleNamespace(fus_model_get_state=lambda h, u, v: 0)

        def init(self, u=None, v=None):
            pass

        def rk4(self, t0, tf, dt):
            n, t = 0, t0
            while t < tf:
                t += min(dt, tf - t)
                n += 1
            return n

        def u_sol(self):
            return np.ones(self.V.ndofs)

        v_sol = u_sol

        def destroy(self):
            pass

    monkeypatch.setattr(fus, "device_count", lambda: 1)
    monkeypatch.setattr(fus, "launch_count", lambda: 0)
    monkeypatch.setattr(fus.Context, "from_mesh", classmethod(from_mesh))
    for name in ("LinearSpectral3D", "LossySpectral3D", "WesterveltSpectral3D"):
        monkeypatch.setattr(fus, name, Model)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "2", "--warmup", "1", "--cells", "3",
                                      "--no-parity", "--model", model] + extra)
    assert bench.main() == 0
    lines = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    need = (REQUIRED - {"impl"}) | {"roofline", "clocks", "gpu_launches", "extras"}
    assert need <= set(d), need - set(d)
    assert d["steps"] == 2 and d["n_gpus"] == 1 and d["dtype"] == "f64" and d["value"] > 0
    assert d["config"]["workload"] == f"{model}_rk4_P4_box3"
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    assert calls["lean"] == ("--lean" in extra)
    if "--geometry-mode" in extra:
        assert ("geometry_mode", 2) in calls["options"] and "trilinear" in d["config"]["geometry"]
        assert d["roofline"]["kernel"].endswith(",2>")
    assert "parity" in d


def test_child_sweep_script_control_flow_with_a_stub_device(monkeypatch, capsys):
    """scripts/bench_sweep.py is what bench.py runs in a child process for its `extras`; like the
    GPU arm it cannot run here, so its control flow is exercised with the device stubbed out:
    every (degree, geometry mode) row and every RK4-by-mode row comes out as JSON."""
    import importlib.util
    import sys

    import numpy as np

    import fenicsx_fus_b200 as fus
    spec = importlib.util.spec_from_file_location("sweep_mod",
                                                  os.path.join(ROOT, "scripts", "bench_sweep.py"))
    sweep = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(sweep)
    t, dist = _fake_torch()
    monkeypatch.setitem(sys.modules, "torch", t)
    monkeypatch.setattr(sweep, "SWEEP", {P: 3 for P in range(2, 8)})

    class Ctx:
        mode = 0

        def set_stream(self, s):
            pass

        def set_option(self, name, value):
            if name == "geometry_mode":
                self.mode = value

        def get_option(self, name):
            return self.mode

        def profile(self, kernel):
            return 80, 16.0

        def destroy(self):
            pass

    class Op:
        def __init__(self, V):
            pass

        def __call__(self, x, c, y):
            return y

    class Model:
        def __init__(self, V, *a, **kw):
            self.V = V

        def init(self, u=None, v=None):
            pass

        def rk4(self, t0, tf, dt):
            return int(round((tf - t0) / dt + 0.5))

        def u_sol(self):
            return np.ones(self.V.ndofs)

        def destroy(self):
            pass

    monkeypatch.setattr(fus.Context, "from_mesh", classmethod(lambda cls, V, device=0, **kw: Ctx()))
    monkeypatch.setattr(fus, "StiffnessSpectral3D", Op)
    monkeypatch.setattr(fus, "LinearSpectral3D", Model)
    monkeypatch.setattr(sys, "argv", ["bench_sweep.py", "--degrees", "2,3,4,5,6,7", "--variants=-1",
                                      "--geometry-modes", "0,1,2,3", "--rk4-geometry-modes",
                                      "0,1,2,3", "--pipeline-variants", "3,4,5", "--models", "",
                                      "--repeats", "2", "--fp32"])
    sweep.main()
    rows = [json.loads(ln) for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")]
    deg = [r for r in rows if r["config"] == "degree_sweep"]
    rk = [r for r in rows if r["config"] == "headline_rk4_by_geometry_mode"]
    assert sorted((r["P"], r["geometry_mode"]) for r in deg if r["variant"] < 0) == [
        (P, g) for P in range(2, 8) for g in (0, 1, 2, 3)]
    assert sorted((r["P"], r["variant"]) for r in deg if r["variant"] >= 0) == [
        (P, v) for P in range(2, 8) for v in (3, 4, 5)]
    assert [r["geometry_mode"] for r in rk] == [0, 1, 2, 3] and all(r["steps"] == 20 for r in rk)
    rkv = [r for r in rows if r["config"] == "headline_rk4_by_pipeline_variant"]
    assert [(r["geometry_mode"], r["variant"]) for r in rkv] == [(0, 3), (0, 4), (0, 5)]
    assert all({"ms_min", "gdof_per_s", "frac_of_measured_peak"} <= set(r) for r in deg)
    f32 = [r for r in rows if r["config"] == "degree_sweep_fp32"]
    assert [r["P"] for r in f32] == list(range(2, 8)) and all("rel_l2_vs_fp64" in r for r in f32)
    assert all({"ms_per_step", "operator_ms", "rel_l2_vs_first_mode"} <= set(r) for r in rk)


@pytest.mark.parametrize("extra", [[], ["--model", "westervelt", "--lean"]])
def test_gpu_arm_on_the_emulated_device(extra):
    """bench.py's GPU arm against the REAL library built for the CPU (tests/emu: every entry point,
    the RK4 loop, per-kernel profiling, the affine extra) with only torch stubbed out -- a small box,
    meaningless timings, but the actual call sequence the driver will run on the B200."""
    code = (
        "import os, sys, importlib.util, json\n"
        f"ROOT = {ROOT!r}\n"
        "sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))\n"
        "spec = importlib.util.spec_from_file_location('bel', os.path.join(ROOT, 'tests', 'emu', "
        "'build_emulated_library.py'))\n"
        "bel = importlib.util.module_from_spec(spec); spec.loader.exec_module(bel)\n"
        "if bel.stale(): bel.build()\n"
        "from fenicsx_fus_b200 import capi\n"
        "capi.LIB_PATH, capi._lib = bel.LIB, None\n"
        "import test_bench_contract as tb\n"
        "t, dist = tb._fake_torch()\n"
        "sys.modules['torch'] = t; sys.modules['torch.distributed'] = dist\n"
        "for k in ('RANK', 'WORLD_SIZE', 'LOCAL_RANK'): os.environ.pop(k, None)\n"
        "spec = importlib.util.spec_from_file_location('bench_emu', os.path.join(ROOT, 'bench.py'))\n"
        "bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)\n"
        f"sys.argv = ['bench.py', '--steps', '3', '--warmup', '1', '--cells', '3', '--no-extras'] + {extra!r}\n"
        "sys.exit(bench.main())\n")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900,
                         cwd=ROOT)
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-3000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["steps"] == 3 and d["value"] > 0 and d["gpu_launches"] > 0
    assert d["roofline"]["launches"] == 12 and d["roofline"]["operator_applications"] == 12
    if "--lean" not in extra:
        # the oracle ran the same steps on the same mesh: the parity block of the JSON line
        par = d["parity"]
        assert par["steps"] == 3 and par["apply_rel_l2"] < 1e-12
        assert par["u_rel_l2"] < 1e-10 and par["v_rel_l2"] < 1e-10
        assert d["cpu_baseline"]["value"] > 0 and d["cpu_baseline"]["cores"] >= 1
    if "--lean" in extra or "--geometry-mode" in extra:
        assert d["roofline"]["kernel"].endswith(",2>")


def test_child_sweep_script_on_the_emulated_device():
    """scripts/bench_sweep.py (bench.py's child process) against the real library on the emulated
    device, torch stubbed: degree sweep in every geometry mode, the RK4 runs per mode, and the FP32
    operator sweep through the Python dispatch on float32 tensors -- tiny boxes."""
    code = (
        "import os, sys, importlib.util\n"
        f"ROOT = {ROOT!r}\n"
        "sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))\n"
        "spec = importlib.util.spec_from_file_location('bel', os.path.join(ROOT, 'tests', 'emu', "
        "'build_emulated_library.py'))\n"
        "bel = importlib.util.module_from_spec(spec); spec.loader.exec_module(bel)\n"
        "if bel.stale(): bel.build()\n"
        "from fenicsx_fus_b200 import capi\n"
        "capi.LIB_PATH, capi._lib = bel.LIB, None\n"
        "import test_bench_contract as tb\n"
        "t, dist = tb._fake_torch()\n"
        "sys.modules['torch'] = t\n"
        "spec = importlib.util.spec_from_file_location('sweep_emu', os.path.join(ROOT, 'scripts', "
        "'bench_sweep.py'))\n"
        "sweep = importlib.util.module_from_spec(spec); spec.loader.exec_module(sweep)\n"
        "sweep.SWEEP = {P: 3 for P in range(2, 8)}\n"
        "sys.argv = ['bench_sweep.py', '--degrees', '2,5', '--variants=-1', '--geometry-modes', "
        "'0,1,2,3', '--rk4-geometry-modes', '0,1,2,3', '--rk4-cells', '3', '--rk4-steps', '3', "
        "'--pipeline-variants', '3,5', '--models', '', '--repeats', '1', '--fp32']\n"
        "sweep.main()\n")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=1200,
                         cwd=ROOT)
    assert res.returncode == 0, res.stdout[-1500:] + res.stderr[-3000:]
    rows = [json.loads(ln) for ln in res.stdout.splitlines() if ln.startswith("{")]
    deg = [r for r in rows if r["config"] == "degree_sweep"]
    assert sorted((r["P"], r["geometry_mode"]) for r in deg if r["variant"] < 0) == [
        (P, g) for P in (2, 5) for g in (0, 1, 2, 3)]
    assert sorted((r["P"], r["variant"]) for r in deg if r["variant"] >= 0) == [
        (P, v) for P in (2, 5) for v in (3, 5)]
    assert all(r["rel_l2_vs_first_config"] < 1e-12 for r in deg)           # same operator in every config
    rk = [r for r in rows if r["config"] == "headline_rk4_by_geometry_mode"]
    assert [r["geometry_mode"] for r in rk] == [0, 1, 2, 3]
    rkv = [r for r in rows if r["config"] == "headline_rk4_by_pipeline_variant"]
    assert [r["variant"] for r in rkv] == [3, 5]
    assert all(r["rel_l2_vs_first_mode"] < 1e-10 for r in rk + rkv)        # same fields in every mode
    f32 = [r for r in rows if r["config"] == "degree_sweep_fp32"]
    assert [r["P"] for r in f32] == [2, 5] and all(0 < r["rel_l2_vs_fp64"] < 1e-5 for r in f32)
