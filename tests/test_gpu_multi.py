"""Multi-GPU parity (needs >= 2 B200s on the box; skipped otherwise): launches
tests/mp_model_check.py under torchrun, one rank per GPU over NCCL."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 4, 8])
def test_partitioned_models_match_single_domain_oracle(fus, gpu, world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29500 + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "mp_model_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1]
    out = json.loads(line)
    assert out["status"] == "ok", out


@pytest.mark.parametrize("kind", ["linear", "westervelt"])
def test_cpp_dropin_multi_rank(fus, gpu, orc, kind, tmp_path):
    """The C++ drop-in on a partitioned mesh, no Python in the loop: examples/linear_box_mp.cpp is
    the reference's SC2-BM1 driver with mesh::create_box on a communicator; its ranks are the threads
    of one process, one per GPU (include/fus/dolfinx_shim.hpp: there is no MPI here), the halo
    set-up happens inside the function space's device context (include/fus/spectral_op.hpp) and the
    time loop exchanges over the fused peer transport.  The owned dofs every rank writes are compared
    with the single-domain oracle."""
    import numpy as np
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from conftest import exe_env, rel_l2
    exe = os.path.join(ROOT, "examples", "linear_box_mp")
    assert os.path.exists(exe), "run __graft_entry__.build() first"
    P, n, steps = 4, 8, 10
    out = str(tmp_path / "u")
    res = subprocess.run([exe, "2", "1", "1", str(n), str(steps), kind, out], capture_output=True,
                         text=True, timeout=600, env=exe_env())
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-2000:]
    vals = dict(ln.split(": ") for ln in res.stdout.strip().splitlines() if ": " in ln)
    assert int(vals["Ranks"]) == 2 and int(vals["Number of steps"]) == steps
    L = 0.12 * n / 54.0
    xg, xd = orc.box_mesh((n, n, n), (0, 0, 0), (L, L, L))
    dm = orc.box_dofmap(P, (n, n, n), 0)                  # dof id == global node index
    nd = int(dm.max()) + 1
    assert int(vals["Degrees of freedom"]) == nd
    G, dJ = orc.geometry(P, xg, xd)
    facets = orc.box_facets((n, n, n))
    fn, fs = orc.facet_data(P, xg, xd, facets)
    nc = dm.shape[0]
    c0, rho0 = np.full(nc, 1500.0), np.full(nc, 1000.0)
    w0 = 2 * np.pi * 0.5e6
    delta = np.full(nc, 2 * 5.0 * 1500.0 ** 3 / w0 / w0)
    om = orc.model(kind, P, nd, dm, G, dJ, orc.dphi(P), c0, rho0, delta if kind != "linear" else None,
                   np.full(nc, 3.5) if kind == "westervelt" else None, facets, fn, fs, 0.5e6, 60000.0,
                   1500.0)
    dt = float(vals["Time step size"])
    u, v = np.zeros(nd), np.zeros(nd)
    assert om.rk4(0.0, (steps - 0.5) * dt, dt, u, v) == steps
    got = np.full(nd, np.nan)
    for r in range(2):
        rec = np.fromfile(f"{out}.{r}.bin", dtype=np.dtype([("g", "<i8"), ("u", "<f8")]))
        assert np.isnan(got[rec["g"]]).all()              # every dof owned by exactly one rank
        got[rec["g"]] = rec["u"]
    assert not np.isnan(got).any() and np.linalg.norm(u) > 0
    assert rel_l2(got, u) < 1e-10
    assert abs(float(vals["u_l2"]) - np.linalg.norm(u)) < 1e-10 * np.linalg.norm(u)
