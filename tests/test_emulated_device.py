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
