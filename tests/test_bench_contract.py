"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the
required keys, and the GPU arm refuses to run without a device (no silent CPU fallback)."""
import json
import os
import subprocess
import sys

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
