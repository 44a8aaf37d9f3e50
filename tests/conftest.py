import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


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
