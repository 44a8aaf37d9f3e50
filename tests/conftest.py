import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_addoption(parser):
    parser.addoption("--emulated-device", action="store_true", default=False,
                     help="run the -m gpu tests against tests/emu/libfus_b200_emulated.so: the whole "
                          "library built for the CPU on the SIMT emulator (host plumbing + kernel "
                          "logic; no statement about the device).  Test infrastructure only.")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "emu_skip: too large or hardware-specific for --emulated-device")
    config.addinivalue_line("markers", "first_hw_run: written after the round's GPU budget was spent "
                            "(verified on the emulated device only): collected after the tests that "
                            "already passed on a B200, so that under -x they cannot hide them")
    if config.getoption("--emulated-device"):
        import importlib.util
        spec = importlib.util.spec_from_file_location(
            "build_emulated_library", os.path.join(ROOT, "tests", "emu", "build_emulated_library.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        if mod.stale():
            mod.build()
        from fenicsx_fus_b200 import capi
        capi.LIB_PATH, capi._lib = mod.LIB, None            # tests only: the product never does this
        libdir = os.path.join(ROOT, "tests", "emu", "_gen", "lib")
        os.makedirs(libdir, exist_ok=True)
        link = os.path.join(libdir, "libfus_b200.so")
        if os.path.lexists(link):
            os.remove(link)
        os.symlink(mod.LIB, link)
        _EMU["libdir"] = libdir


def pytest_collection_modifyitems(config, items):
    items.sort(key=lambda item: "first_hw_run" in item.keywords)          # stable: order otherwise kept
    if config.getoption("--emulated-device"):
        skip = pytest.mark.skip(reason="not meaningful / too large on the emulated device")
        for item in items:
            if "emu_skip" in item.keywords:
                item.add_marker(skip)


_EMU = {"libdir": None}


def exe_env():
    """Environment for the example executables (examples/*): unchanged on a GPU box; with
    --emulated-device a directory holding the emulated build under the product library's name is put
    on LD_LIBRARY_PATH (searched before the executables' RUNPATH), so that they run on it."""
    env = dict(os.environ)
    if _EMU["libdir"]:
        env["LD_LIBRARY_PATH"] = _EMU["libdir"] + os.pathsep + env.get("LD_LIBRARY_PATH", "")
    return env


@pytest.fixture(scope="session")
def emulated(request):
    """True when the -m gpu tests run against the emulated device (smaller workloads then)."""
    return bool(request.config.getoption("--emulated-device"))


@pytest.fixture(scope="session")
def orc():
    """The plain-C oracle (oracle/fus_oracle.c)."""
    from oracle.oracle import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def orc_ref():
    """oracle/_ref: cell loops on the reference's own sum_factorisation.hpp (prebuilt or built here)."""
    from oracle.oracle import Oracle, ref_available
    if not ref_available() and not os.path.isdir("/root/reference/cpp/fenicsx-sf/common"):
        pytest.skip("oracle/_ref not built and /root/reference absent")
    return Oracle(ref=True)


@pytest.fixture(scope="session")
def fus():
    import fenicsx_fus_b200 as m
    return m


@pytest.fixture(scope="session")
def gpu(fus):
    if fus.device_count() < 1:
        pytest.fail("no sm_100 device visible: -m gpu tests need a B200 (there is no fallback)")
    return 0


def warp_vertices(x, amp=0.08, seed=7):
    """Smooth + random displacement of box vertices: non-affine trilinear cells."""
    rng = np.random.default_rng(seed)
    span = x.max(0) - x.min(0)
    n_est = max(2.0, round(len(x) ** (1 / 3)) - 1)
    h = span / n_est
    y = x.copy()
    y += amp * h * rng.uniform(-1, 1, x.shape)
    y[:, 0] += 0.05 * span[0] * np.sin(2.0 * x[:, 1] / span[1]) * np.cos(1.5 * x[:, 2] / span[2])
    return y


def rel_l2(a, b):
    nb = np.linalg.norm(b)
    return np.linalg.norm(a - b) / (nb if nb > 0 else 1.0)
