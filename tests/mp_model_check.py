"""Multi-GPU worker (launched by torchrun from tests/test_gpu_multi.py or by hand):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tests/mp_model_check.py

Each rank owns one block of a small box, sets up the NCCL halo of its context and checks
  * fus_scatter_fwd_dev / fus_scatter_rev_dev against the ownership maps (bit-exact), and
  * the three models after 10 RK4 steps against the single-domain CPU oracle (<= 1e-10).
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

PGRID = {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def main():
    import torch
    import torch.distributed as dist

    import fenicsx_fus_b200 as fus
    from fenicsx_fus_b200 import capi
    from fenicsx_fus_b200.partition import BoxPartition
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = capi.load()
    P, n, h = 3, (4, 4, 4), 0.002
    hi = tuple(h * k for k in n)
    pg = PGRID[world]
    part = BoxPartition(P, n, pg, rank, lo=(0, 0, 0), hi=hi)
    V = part.function_space(device=local)
    ctx = V.context(local)
    part.setup_halo(ctx, dist)
    report = {}

    # ---- scatter_fwd / scatter_rev on the device ------------------------------------------
    key = part.global_key.astype(np.float64)
    x = np.zeros(part.ndofs)
    x[:part.nowned] = key[:part.nowned] * 0.5 + 1.0
    d = ctx.alloc(8 * part.ndofs)
    ctx.upload(d, x)
    fus.check(lib.fus_scatter_fwd_dev(ctx.h, d), "scatter_fwd")
    ctx.sync()
    got = np.zeros(part.ndofs)
    ctx.download(got, d)
    ok_fwd = bool(np.array_equal(got, key * 0.5 + 1.0))
    ones = np.ones(part.ndofs)
    ctx.upload(d, ones)
    fus.check(lib.fus_scatter_rev_dev(ctx.h, d), "scatter_rev")
    ctx.sync()
    ctx.download(got, d)
    copies = np.ones(part.ndofs)
    for s in part.send_lists:
        np.add.at(copies, s, 1.0)
    ok_rev = bool(np.array_equal(got[:part.nowned], copies[:part.nowned]))
    ctx.free(d)
    report["scatter_fwd_exact"], report["scatter_rev_exact"] = ok_fwd, ok_rev

    # ---- models ------------------------------------------------------------------------------
    f, p0, s0 = 0.5e6, 2.0e6, 1500.0
    w0 = 2 * np.pi * f
    cg = part.cell_global
    c0 = np.where(cg % 3 == 0, 2300.0, 1500.0)
    rho0 = np.where(cg % 3 == 0, 1700.0, 1000.0)
    delta0 = np.full(part.ncells, fus.compute_diffusivity_of_sound(w0, 1500.0, 5.0))
    beta0 = np.full(part.ncells, 3.5)
    dt = 0.3 * 0.65 * np.sqrt(3) * h / (2300.0 * P * P)
    u0 = 1e5 * np.sin(0.013 * key) + 3.0
    v0 = 1e11 * np.cos(0.007 * key)
    results = {}
    for overlap in (1, 0, "peer"):
        if overlap == "peer":
            # one-sided puts into the neighbours' mailboxes (CUDA IPC peer memory)
            report["peer_connected"] = bool(part.connect_peers(ctx, dist))
            if not report["peer_connected"]:
                break
        else:
            ctx.set_option("halo_overlap", overlap)
        for kind in ("linear", "lossy", "westervelt"):
            if kind == "linear":
                mdl = fus.LinearSpectral3D(V, c0, rho0, f, p0, s0, facets=part.facets, device=local)
            elif kind == "lossy":
                mdl = fus.LossySpectral3D(V, c0, rho0, delta0, f, p0, s0, facets=part.facets,
                                          device=local)
            else:
                mdl = fus.WesterveltSpectral3D(V, c0, rho0, delta0, beta0, f, p0, s0,
                                               facets=part.facets, device=local)
            mdl.init(u0.copy(), v0.copy())
            steps = mdl.rk4(0.0, 10 * dt - 0.3 * dt, dt)
            results[(kind, overlap)] = (steps, mdl.u_sol()[:part.nowned].copy(),
                                        mdl.v_sol()[:part.nowned].copy(),
                                        mdl.mass()[:part.nowned].copy())
            mdl.destroy()
    # ---- unstructured mesh (the reference's own test mesh), partitioned with HexPartition -------
    from fenicsx_fus_b200.partition import HexPartition
    from fenicsx_fus_b200.unstructured import HexFunctionSpace, HexMesh
    gm = np.load(os.path.join(ROOT, "tests", "golden", "ref_mesh_hex6312.npz"))
    umesh = HexMesh(gm["geometry"], gm["topology_vtk"][:, (0, 1, 3, 2, 4, 5, 7, 6)],
                    gm["facet_quads"], gm["facet_values"], reorder="morton")
    UP = 2
    UV = HexFunctionSpace(umesh, UP)
    hp = HexPartition(umesh, UP, world, rank, space=UV)
    hV = hp.function_space(device=local)
    hctx = hV.context(local)
    hp.setup_halo(hctx, dist)
    report["unstructured_peer_connected"] = bool(hp.connect_peers(hctx, dist))
    ucent = umesh.x[umesh.xdofmap].mean(axis=1)[:, 0]
    uc0g = np.where(ucent < 0.5, 1500.0, 2300.0)
    urho0g = np.where(ucent < 0.5, 1000.0, 1700.0)
    uf, up0 = 2.0e3, 1.0e5
    udt = 0.15 * umesh.h_min() / (2300.0 * UP * UP)
    umdl = fus.LinearSpectral3D(hV, uc0g[hp.cell_global], urho0g[hp.cell_global], uf, up0, 1500.0,
                                facets=hp.facets, device=local)
    ukey = hp.global_key.astype(np.float64)
    umdl.init(1e2 * np.sin(0.013 * ukey), 1e6 * np.cos(0.007 * ukey))
    usteps = umdl.rk4(1e-4, 1e-4 + 8 * udt, udt)
    results[("unstructured_linear", "peer")] = (usteps, umdl.u_sol()[:hp.nowned].copy(),
                                                hp.global_key[:hp.nowned].copy())
    umdl.destroy()

    gathered = [None] * world
    dist.gather_object((part.global_key[:part.nowned], results, report), gathered if rank == 0 else None,
                       dst=0)
    status = 0
    if rank == 0:
        from oracle.oracle import Oracle
        orc = Oracle()
        xg, xd = orc.box_mesh(n, (0, 0, 0), hi)
        dm = orc.box_dofmap(P, n, 0)                    # dof id == global key
        nd = dm.max() + 1
        G, dJ = orc.geometry(P, xg, xd)
        facets = orc.box_facets(n)
        fn, fs = orc.facet_data(P, xg, xd, facets)
        gc = np.arange(dm.shape[0])
        c0g = np.where(gc % 3 == 0, 2300.0, 1500.0)
        rho0g = np.where(gc % 3 == 0, 1700.0, 1000.0)
        keyg = np.arange(nd, dtype=np.float64)
        out = {"world": world, "scatter": [g[2] for g in gathered]}
        for kind in ("linear", "lossy", "westervelt"):
            om = orc.model(kind, P, nd, dm, G, dJ, orc.dphi(P), c0g, rho0g,
                           np.full(dm.shape[0], delta0[0]), np.full(dm.shape[0], 3.5), facets, fn, fs,
                           f, p0, s0)
            u, v = 1e5 * np.sin(0.013 * keyg) + 3.0, 1e11 * np.cos(0.007 * keyg)
            steps = om.rk4(0.0, 10 * dt - 0.3 * dt, dt, u, v)
            for overlap in (1, 0, "peer"):
                if (kind, overlap) not in gathered[0][1]:
                    continue
                gu, gv, gm = np.zeros(nd), np.zeros(nd), np.zeros(nd)
                for keys, res, _ in gathered:
                    st, uu, vv, mm = res[(kind, overlap)]
                    assert st == steps
                    gu[keys], gv[keys], gm[keys] = uu, vv, mm
                eu = np.linalg.norm(gu - u) / np.linalg.norm(u)
                ev = np.linalg.norm(gv - v) / np.linalg.norm(v)
                em = np.linalg.norm(gm - om.mass()) / np.linalg.norm(om.mass())
                out[f"{kind}_overlap{overlap}"] = {"u": eu, "v": ev, "mass": em}
                if not (eu < 1e-10 and ev < 1e-10 and em < 1e-12):
                    status = 1
        # unstructured: single-domain oracle on the global numbering
        Gu, dJu = orc.geometry(UP, umesh.x, umesh.xdofmap)
        fnu, fsu = orc.facet_data(UP, umesh.x, umesh.xdofmap, umesh.facets)
        omu = orc.model("linear", UP, UV.ndofs, UV.dofmap, Gu, dJu, orc.dphi(UP), uc0g, urho0g, None,
                        None, umesh.facets, fnu, fsu, uf, up0, 1500.0)
        kg = np.arange(UV.ndofs, dtype=np.float64)
        uu, vv = 1e2 * np.sin(0.013 * kg), 1e6 * np.cos(0.007 * kg)
        ust = omu.rk4(1e-4, 1e-4 + 8 * udt, udt, uu, vv)
        gu = np.zeros(UV.ndofs)
        for _, res, _ in gathered:
            st, ul, keys = res[("unstructured_linear", "peer")]
            assert st == ust
            gu[keys] = ul
        eu = np.linalg.norm(gu - uu) / np.linalg.norm(uu)
        out["unstructured_linear_peer"] = {"u": eu}
        if not eu < 1e-10:
            status = 1
        if not all(s["scatter_fwd_exact"] and s["scatter_rev_exact"] for s in out["scatter"]):
            status = 1
        if not all(s.get("peer_connected") for s in out["scatter"]):
            status = 1
        out["status"] = "ok" if status == 0 else "FAILED"
        print(json.dumps(out), flush=True)
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", f"multi_gpu_check_w{world}.json"), "w") as fjs:
            json.dump(out, fjs, indent=1)
    st = torch.tensor([status], device="cuda")
    dist.broadcast(st, src=0)
    dist.destroy_process_group()
    sys.exit(int(st.item()))


if __name__ == "__main__":
    main()
