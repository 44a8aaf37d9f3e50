"""Oracle parity AT THE SIZES THE NUMBERS ARE QUOTED ON (BASELINE.json configs 1-4, SURVEY.md
section 8d): the GPU path against the reference's own kernels (oracle/_ref: cell loops on the
unmodified sum_factorisation.hpp, all host cores) on the same mesh arrays, numbering and time steps.

  config 1  linear, P=4, 54^3 cells (10 218 313 dofs), 100 RK4 steps from rest
  config 2  one operator application per degree P=2..7 on the ~10 M-dof boxes of the degree sweep
  config 3  heterogeneous media (five layers), P=4, 54^3 cells, 20 steps
  config 4  Westervelt (HITU water parameters, disc source), P=4, 54^3 cells, 20 steps; lossy likewise

Tolerances are BASELINE.json's: one application 1e-12, fields after N steps 1e-10 (relative L2), as the
reference's own test compares its sum-factorised operator with an independent evaluation
(cpp/fenicsx-sf/tests/test_operators3d/main.cpp:100-166).  The model set-up is bench.py's, so these
are the workloads bench.py times.  About a minute of CPU work each; skipped on the emulated device."""
import importlib.util
import os
import types

import numpy as np
import pytest

from conftest import ROOT, rel_l2

pytestmark = [pytest.mark.gpu, pytest.mark.emu_skip]


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_for_tests", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _run(fus, bench, model, steps, note=None):
    from fenicsx_fus_b200 import partition
    P, n = 4, 54
    args = types.SimpleNamespace(model=model, lean=False, geometry_mode=0)
    bench.P_BENCH, bench.N_BENCH = P, n
    part, V, ctx, mdl, prm, facets, _, _ = bench.build_problem(
        fus, partition, None, args, P, n, (1, 1, 1), 0, 0, 1)
    assert V.ndofs == 10218313
    mdl.init()
    assert mdl.rk4(0.0, (steps - 0.5) * prm["dt"], prm["dt"]) == steps
    u, v = mdl.u_sol(), mdl.v_sol()
    mdl.destroy()
    ctx.destroy()
    V._ctx = None
    ref = bench.cpu_model_rk4(P, steps, 0, prm=prm,
                              mesh=(part.x, part.xdofmap, part.dofmap, facets, V.ndofs), fields=True)
    assert ref["steps"] == steps and np.linalg.norm(ref["u"]) > 0
    eu, ev = rel_l2(u, ref["u"]), rel_l2(v, ref["v"])
    print(f"{model}: {steps} steps at 54^3, u {eu:.2e}, v {ev:.2e} ({ref['kind']} kernels, "
          f"{ref['cores']} threads, {ref['seconds']:.1f} s)")
    return eu, ev


def test_config1_linear_100_steps_at_full_size(fus, gpu, bench):
    """BASELINE config 1 as quoted: 100 steps of LinearSpectral3D on the 10.2 M-dof box."""
    eu, ev = _run(fus, bench, "linear", 100)
    assert eu < 1e-10 and ev < 1e-10


@pytest.mark.parametrize("model", ["linear_het", "lossy", "westervelt"])
def test_configs_3_and_4_at_full_size(fus, gpu, bench, model):
    """Heterogeneous media (config 3), lossy, Westervelt with the disc source (config 4), 20 steps."""
    eu, ev = _run(fus, bench, model, 20)
    assert eu < 1e-10 and ev < 1e-10


@pytest.mark.parametrize("P", [2, 3, 4, 5, 6, 7])
def test_config2_operator_apply_at_sweep_size(fus, gpu, orc_ref, P):
    """One application of the kernel the library picks for the degree, on the box of the degree
    sweep (9.8 - 10.4 M dofs), against the reference kernels, coefficient -1/1000 as in
    measure_fraction_of_peak_performance/main.cpp:97-104:
      * a seeded uniform(-1, 1) vector at BASELINE.json's 1e-12;
      * the experiment's own input u = sin(x) cos(pi y) (main.cpp:75-82) at 1e-11.  For that smooth
        field ||K u|| is ~3e3 times smaller against ||u|| than for a random one (K u = O(h^2)), and
        every implementation's rounding is amplified accordingly: the reference's own kernels
        (-Ofast) and the plain-C restatement of them differ by 5e-14 / 1.6e-13 at 12^3 / 24^3 cells
        on this input and by 1.5e-16 on the random one (measured in the build container)."""
    n = {2: 107, 3: 71, 4: 54, 5: 43, 6: 36, 7: 31}[P]
    m = fus.BoxMesh((n, n, n))
    V = fus.FunctionSpace(m, P, numbering=1)
    X = V.tabulate_dof_coordinates()
    smooth = np.sin(X[:, 0]) * np.cos(np.pi * X[:, 1])
    del X
    rnd = np.random.default_rng(12345).uniform(-1.0, 1.0, V.ndofs)
    coeffs = np.full(m.ncells, -1.0 / 1000.0)
    K = fus.StiffnessSpectral3D(V)
    y_rnd = K(rnd, coeffs, np.zeros(V.ndofs))
    y_smooth = K(smooth, coeffs, np.zeros(V.ndofs))
    ctx = V.context()
    V._ctx = None
    ctx.destroy()
    orc_ref.lib.fr_set_threads(os.cpu_count() or 1)
    G, _ = orc_ref.geometry(P, m.x, m.xdofmap, want_detJ=False)
    dphi = orc_ref.dphi(P)
    e_rnd = rel_l2(y_rnd, orc_ref.stiffness_apply(P, V.dofmap, G, dphi, coeffs, rnd, np.zeros(V.ndofs),
                                                  use_ref_kernels=True))
    e_smooth = rel_l2(y_smooth, orc_ref.stiffness_apply(P, V.dofmap, G, dphi, coeffs, smooth,
                                                        np.zeros(V.ndofs), use_ref_kernels=True))
    print(f"P={P}: {V.ndofs} dofs, apply rel L2: random {e_rnd:.2e}, sin(x)cos(pi y) {e_smooth:.2e}")
    assert e_rnd < 1e-12
    assert e_smooth < 1e-11
