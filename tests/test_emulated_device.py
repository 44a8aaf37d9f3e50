"""A slice of the `-m gpu` suite on the *emulated device* as part of the CPU suite.

`pytest -m gpu --emulated-device` runs the GPU tests against tests/emu/libfus_b200_emulated.so: the
whole library (fus_capi.cu, fus_halo.cu with every kernel launch rewritten, the kernels through the
SIMT emulator, a synchronous stand-in for the CUDA runtime) built for the CPU.  That exercises the host
plumbing of every entry point and the kernels' logic without a GPU; it makes no statement about the
device.  The full emulated run takes ~15 minutes; this test runs a representative slice on every CPU
run so that the plumbing of the newest paths (FP32 operators, lean contexts, on-the-fly geometry, 2-D,
C ABI from C) cannot rot between GPU runs."""
import os
import subprocess
import sys

from conftest import ROOT

SLICE = ("test_fp32_operators_vs_oracle and (2 or 4) or test_lean_context_models_vs_oracle and lossy "
         "or test_trilinear_geometry_on_the_fly and 3 or test_gpu_operators_2d_vs_oracle and 3 "
         "or test_gpu_models_2d_vs_oracle and linear or test_c_abi_from_plain_c "
         "or test_ragged_cell_counts or test_models_golden and westervelt "
         "or test_affine_geometry_compression and 2")


def test_gpu_suite_slice_on_the_emulated_device():
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests"), "-m", "gpu",
                          "--emulated-device", "-q", "-x", "-p", "no:cacheprovider", "-k", SLICE],
                         capture_output=True, text=True, timeout=1500, cwd=ROOT)
    tail = "\n".join(res.stdout.splitlines()[-15:])
    assert res.returncode == 0, tail + res.stderr[-2000:]
    assert " passed" in tail and "failed" not in tail


GRAPH_CHECK = r"""
import ctypes, importlib.util, os, sys
import numpy as np
ROOT = sys.argv[1]
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("bel", os.path.join(ROOT, "tests", "emu",
                                                                   "build_emulated_library.py"))
bel = importlib.util.module_from_spec(spec); spec.loader.exec_module(bel)
if bel.stale():
    bel.build()
from fenicsx_fus_b200 import capi
capi.LIB_PATH, capi._lib = bel.LIB, None
import fenicsx_fus_b200 as fus
from oracle.oracle import Oracle
lib = ctypes.CDLL(bel.LIB)
lib.fus_emu_graph_launches.restype = ctypes.c_longlong
replays = lib.fus_emu_graph_launches

orc = Oracle()
P, n, h = 3, (3, 2, 2), 0.002
m = fus.BoxMesh(n, (0, 0, 0), tuple(h * k for k in n))
V = fus.FunctionSpace(m, P, numbering=1)
ctx = V.context()
G, dJ = orc.geometry(P, m.x, m.xdofmap)
fn, fs = orc.facet_data(P, m.x, m.xdofmap, m.facets)
nc = m.ncells
def rel(a, b): return np.linalg.norm(a - b) / np.linalg.norm(b)
dt = 0.65 * np.sqrt(3) * h / (1500.0 * P * P)
rng = np.random.default_rng(0)
u0, v0 = 1e3 * rng.uniform(-1, 1, V.ndofs), 1e9 * rng.uniform(-1, 1, V.ndofs)
for kind in ("linear", "westervelt"):
    if kind == "linear":
        mdl = fus.LinearSpectral3D(V, 1500.0, 1000.0, 0.5e6, 6.0e4, 1500.0)
        om = orc.model("linear", P, V.ndofs, V.dofmap, G, dJ, orc.dphi(P), np.full(nc, 1500.0),
                       np.full(nc, 1000.0), None, None, m.facets, fn, fs, 0.5e6, 6.0e4, 1500.0)
        step = dt
    else:
        mdl = fus.WesterveltSpectral3D(V, 1500.0, 1000.0, 3e-3, 5.0, 0.5e6, 6.0e4, 1500.0)
        om = orc.model("westervelt", P, V.ndofs, V.dofmap, G, dJ, orc.dphi(P), np.full(nc, 1500.0),
                       np.full(nc, 1000.0), np.full(nc, 3e-3), np.full(nc, 5.0), m.facets, fn, fs,
                       0.5e6, 6.0e4, 1500.0)
        step = 0.3 * dt
    def run(nsteps, step=step):
        mdl.init(u0.copy(), v0.copy())
        before = replays()
        # the reference loop shortens the last step (dt = min(dt, tf - t)): K - 0.5 -> K steps
        assert mdl.rk4(0.0, (nsteps - 0.5) * step, step) == nsteps
        u, v = u0.copy(), v0.copy()
        om.rk4(0.0, (nsteps - 0.5) * step, step, u, v)
        assert rel(mdl.u_sol(), u) < 1e-10 and rel(mdl.v_sol(), v) < 1e-10, kind
        return replays() - before
    # step 0 eager, step 1 captured and launched, steps 2..6 replayed, step 7 (shortened) eager
    assert run(8) == 6, kind
    # a later call with the same dt reuses the executable graph from its first full step on
    assert run(5) == 4, kind
    # another dt: the old graph must not be replayed (a new one is captured at step 1)
    assert run(6, 0.5 * step) == 4, kind
    # an option that changes the launches invalidates it as well
    ctx.set_option("stiffness_variant", 2)
    assert run(6, 0.5 * step) == 4, kind
    ctx.set_option("geometry_mode", 2)
    assert ctx.get_option("geometry_compressed") == 2 and run(6, 0.5 * step) == 4, kind
    ctx.set_option("geometry_mode", 0)
    ctx.set_option("stiffness_variant", -1)
    # per-kernel profiling and use_graph = 0 run eagerly
    ctx.set_option("use_graph", 0)
    assert run(6) == 0, kind
    ctx.set_option("use_graph", 1)
    ctx.set_option("profile_kernels", 1)
    assert run(6) == 0, kind
    ctx.set_option("profile_kernels", 0)
    assert run(6) > 0, kind
    mdl.destroy()
# capture refused (as for the legacy default stream on a device): eager issue, same numbers
os.environ["FUS_EMU_NO_GRAPH"] = "1"
mdl = fus.LinearSpectral3D(V, 1500.0, 1000.0, 0.5e6, 6.0e4, 1500.0)
om = orc.model("linear", P, V.ndofs, V.dofmap, G, dJ, orc.dphi(P), np.full(nc, 1500.0),
               np.full(nc, 1000.0), None, None, m.facets, fn, fs, 0.5e6, 6.0e4, 1500.0)
mdl.init(u0.copy(), v0.copy())
before = replays()
assert mdl.rk4(0.0, 5.5 * dt, dt) == 6 and replays() == before
u, v = u0.copy(), v0.copy()
om.rk4(0.0, 5.5 * dt, dt, u, v)
assert rel(mdl.u_sol(), u) < 1e-10
print("graph path ok")
"""


def test_step_graph_path_on_the_emulated_device():
    """fus_model_rk4's CUDA-graph path -- capture of one step, replay, the device-side step counter
    that walks the source table, invalidation on dt / option changes, eager fallbacks -- with the
    emulated runtime recording launches during capture and replaying them (tests/emu/cuda_runtime.h).
    The replay counter proves which path ran; fields are checked against the oracle every time."""
    res = subprocess.run([sys.executable, "-c", GRAPH_CHECK, ROOT], capture_output=True, text=True,
                         timeout=1200, cwd=ROOT)
    assert res.returncode == 0 and "graph path ok" in res.stdout, res.stdout[-1500:] + res.stderr[-3000:]
