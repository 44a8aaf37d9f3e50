"""Generate the golden fixtures in tests/golden/ (run in the build container, where
/root/reference is mounted):

    python tests/golden/make_golden.py

Sources of truth:
  * kat_sum_factorisation.json -- output of the reference's own known-answer program
    (cpp/mwe/sum_factorisation/main.cpp:10-62) reproduced by calling the UNMODIFIED reference
    header through oracle/_ref (fr_kat).  The expected numbers also appear in SURVEY.md section 4.
  * stiffness_P*.npz, mass_P*.npz -- operator applications computed by oracle/_ref, i.e. the cell
    loop of spectral_op.hpp:183-242 / :75-85 on top of the reference's contract<>/transpose<>
    templates, on a warped 2x2x2 box with seeded input.
  * rk4_*.npz -- a few RK4 steps of each model by the oracle's literal restatement of
    Linear.hpp / Lossy.hpp / Westervelt.hpp rk4+f1 driving those same reference kernels.
The GLL tables / Jacobians inside come from the oracle's restatement of Basix/DOLFINx (no golden
data for those exists in the reference: "parity unpinned" for them).
The arrays needed to replay each case are stored in the fixture so that the tests do not depend
on any mesh generator.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

from conftest import warp_vertices  # noqa: E402
from oracle.oracle import Oracle  # noqa: E402


def main():
    ref = Oracle(ref=True)
    ref.lib.fr_set_threads(1)

    out, out_t = np.zeros(12), np.zeros(12)
    ref.lib.fr_kat(out, out_t)
    with open(os.path.join(HERE, "kat_sum_factorisation.json"), "w") as f:
        json.dump({"source": "cpp/mwe/sum_factorisation/main.cpp:10-62 via oracle/_ref fr_kat",
                   "out": out.tolist(), "out_transposed": out_t.tolist()}, f, indent=1)

    # ---- operator applications -------------------------------------------------------
    n = (2, 2, 2)
    for P in (2, 3, 4, 5):
        rng = np.random.default_rng(1000 + P)
        xg, xd = ref.box_mesh(n)
        xg = warp_vertices(xg, amp=0.1, seed=P)
        dm = ref.box_dofmap(P, n, 1)
        nd = int(dm.max()) + 1
        G, dJ = ref.geometry(P, xg, xd)
        dphi = ref.dphi(P)
        coeffs = rng.uniform(0.5, 2.0, dm.shape[0])
        x = rng.uniform(-1, 1, nd)
        y0 = rng.uniform(-1, 1, nd)          # operator() accumulates: start from non-zero y
        y = y0.copy()
        ref.stiffness_apply(P, dm, G, dphi, coeffs, x, y, use_ref_kernels=True)
        np.savez_compressed(os.path.join(HERE, f"stiffness_P{P}.npz"), P=P, xg=xg, xd=xd, dofmap=dm,
                            G=G, detJ=dJ, dphi=dphi, coeffs=coeffs, x=x, y0=y0, y=y)
        ym = y0.copy()
        ref.mass_apply(P, dm, dJ, coeffs, x, ym, use_ref_kernels=True)
        np.savez_compressed(os.path.join(HERE, f"mass_P{P}.npz"), P=P, dofmap=dm, detJ=dJ,
                            coeffs=coeffs, x=x, y0=y0, y=ym)

    # ---- short RK4 runs ----------------------------------------------------------------
    P, n = 3, (4, 2, 2)
    lo, hi = (0.0, 0.0, 0.0), (0.008, 0.004, 0.004)
    xg, xd = ref.box_mesh(n, lo, hi)
    xg = warp_vertices(xg, amp=0.06, seed=11)
    dm = ref.box_dofmap(P, n, 1)
    nd = int(dm.max()) + 1
    nc = dm.shape[0]
    G, dJ = ref.geometry(P, xg, xd)
    dphi = ref.dphi(P)
    facets = ref.box_facets(n)
    fn, fs = ref.facet_data(P, xg, xd, facets)
    rng = np.random.default_rng(77)
    c0 = np.where(np.arange(nc) < nc // 2, 1500.0, 2300.0)      # two media
    rho0 = np.where(np.arange(nc) < nc // 2, 1000.0, 1700.0)
    freq, p0, s0 = 0.5e6, 2.0e6, 1500.0
    w0 = 2 * np.pi * freq
    delta0 = np.full(nc, 2 * 5.0 * 1500.0 ** 3 / w0 ** 2)       # alpha = 5 Np/m
    beta0 = np.full(nc, 3.5)
    h = 0.002
    # CFL 0.65 on the cell diameter (BM7-SC1/main.cpp:112-118) is stable for the wave operator but
    # NOT for the absorbing term of the lossy/Westervelt forms, which `ds` applies on every exterior
    # facet: a corner node sees the rate 3 c / (w_0 h) and explicit RK4 needs rate*dt < 2.78.
    # The fixtures therefore use 0.35 x that step for all three models.
    dt0 = 0.35 * 0.65 * (np.sqrt(3) * h) / (2300.0 * P * P)
    steps_per_period = int((1 / freq) / dt0) + 1
    dt = (1 / freq) / steps_per_period
    nsteps = 12
    # smooth non-zero initial data so that every term is active from the first stage
    # (dof coordinates are not needed: use a seeded smooth-ish random field)
    u_init = 1.0e6 * rng.uniform(-1, 1, nd)
    v_init = 1.0e12 * rng.uniform(-1, 1, nd)
    for kind in ("linear", "lossy", "westervelt"):
        mdl = ref.model(kind, P, nd, dm, G, dJ, dphi, c0, rho0, delta0, beta0, facets, fn, fs,
                        freq, p0, s0, use_ref_kernels=True)
        u, v = u_init.copy(), v_init.copy()
        kv = mdl.f1(0.3 / freq, u, v)
        taken = mdl.rk4(0.0, nsteps * dt - 0.25 * dt, dt, u, v)   # last step is a short one
        np.savez_compressed(os.path.join(HERE, f"rk4_{kind}.npz"), P=P, n=np.array(n), xg=xg, xd=xd,
                            dofmap=dm, G=G, detJ=dJ, dphi=dphi, facets=facets, fnodes=fn,
                            fscale=fs, c0=c0, rho0=rho0, delta0=delta0, beta0=beta0, freq=freq,
                            p0=p0, s0=s0, dt=dt, t0=0.0, tf=nsteps * dt - 0.25 * dt,
                            u_init=u_init, v_init=v_init, f1_t=0.3 / freq, f1=kv, mass=mdl.mass(),
                            steps=taken, u=u, v=v)
        print(kind, "steps", taken, "|u|", np.linalg.norm(u), "|v|", np.linalg.norm(v))
    tot = sum(os.path.getsize(os.path.join(HERE, f)) for f in os.listdir(HERE))
    print("golden bytes:", tot)


if __name__ == "__main__":
    main()
