"""CPU tests of the multi-GPU host logic: ownership, halo lists, interface-first ordering, and a
world_size-2 gloo run of the distributed operator algorithm (oracle kernels stand in for the GPU)."""
import os
import socket

import numpy as np
import pytest

from conftest import ROOT, rel_l2


def _parts(P, n, pg):
    from fenicsx_fus_b200.partition import BoxPartition
    nr = int(np.prod(pg))
    return [BoxPartition(P, n, pg, r, lo=(0, 0, 0), hi=(1.0, 0.7, 0.9)) for r in range(nr)]


def _random_partition_cases(count=10, seed=2026):
    """Seeded random (degree, box, process grid) triples: uneven splits, 3- and 4-way splits,
    single-cell blocks -- grids the 8-GPU runs (2x2x2) never exercise."""
    rng = np.random.default_rng(seed)
    cases = []
    while len(cases) < count:
        pg = tuple(int(v) for v in rng.integers(1, 5, 3))
        if np.prod(pg) > 12:
            continue
        n = tuple(int(pg[d] + rng.integers(0, 4)) for d in range(3))
        cases.append((int(rng.integers(1, 5)), n, pg))
    return cases


@pytest.mark.parametrize("P,n,pg", [(2, (4, 4, 4), (2, 2, 2)), (3, (5, 3, 2), (2, 1, 1)),
                                    (4, (4, 6, 2), (2, 2, 1)), (1, (3, 3, 3), (3, 1, 3))]
                         + _random_partition_cases())
def test_ownership_and_halo_lists(fus, P, n, pg):
    parts = _parts(P, n, pg)
    nglob = np.prod([n[d] * P + 1 for d in range(3)])
    owned = np.concatenate([p.global_key[:p.nowned] for p in parts])
    # every global dof is owned exactly once
    assert owned.size == nglob and np.array_equal(np.sort(owned), np.arange(nglob))
    for p in parts:
        assert p.ndofs_global == nglob
        assert len(np.unique(p.global_key)) == p.ndofs          # local numbering is a bijection
        assert sorted(np.unique(p.dofmap)) == list(range(p.ndofs))
        for q, s, r in zip(p.neigh, p.send_lists, p.recv_lists):
            o = parts[q]
            k = o.neigh.index(p.rank)
            # what I send is exactly what q expects from me, in the same order, and vice versa
            assert np.array_equal(p.global_key[s], o.global_key[o.recv_lists[k]])
            assert np.array_equal(p.global_key[r], o.global_key[o.send_lists[k]])
            assert np.all(s < p.nowned) and np.all(r >= p.nowned)
        # every ghost is received exactly once
        rec = np.concatenate(p.recv_lists) if p.neigh else np.zeros(0, int)
        assert np.array_equal(np.sort(rec), np.arange(p.nowned, p.ndofs))
        # interior cells touch no shared dof
        shared = np.zeros(p.ndofs, bool)
        shared[p.nowned:] = True
        for s in p.send_lists:
            shared[s] = True
        touch = shared[p.dofmap].any(1)
        assert touch[:p.ninterface_cells].all() and not touch[p.ninterface_cells:].any()
    # cells: a partition of the global cells
    cells = np.concatenate([p.cell_global for p in parts])
    assert np.array_equal(np.sort(cells), np.arange(np.prod(n)))
    # exterior facets: a partition of the global exterior facets
    nf = sum(p.facets.shape[0] for p in parts)
    assert nf == 2 * (n[0] * n[1] + n[1] * n[2] + n[0] * n[2])


@pytest.mark.parametrize("P,n,pg", [(2, (4, 3, 5), (2, 1, 1)), (3, (4, 4, 4), (2, 2, 2)),
                                    (1, (5, 3, 2), (1, 3, 2)), (4, (6, 4, 2), (3, 2, 1)),
                                    (2, (3, 3, 3), (1, 1, 1))])
def test_native_partitioner_equals_numpy_reference(fus, P, n, pg):
    """fus_box_partition_create (csrc/fus_partition.cpp, what bench.py and the GPU checks use) against
    the numpy implementation of the same rules: every array and list identical on every rank."""
    from fenicsx_fus_b200.partition import BoxPartition
    for rank in range(int(np.prod(pg))):
        for numbering in (0, 1):
            a = BoxPartition(P, n, pg, rank, lo=(0.1, 0, 0), hi=(1, 2, 3), numbering=numbering)
            b = BoxPartition(P, n, pg, rank, lo=(0.1, 0, 0), hi=(1, 2, 3), numbering=numbering,
                             native=False)
            for k in ("dofmap", "xdofmap", "cell_global", "global_key", "facets", "x", "n_local"):
                assert np.array_equal(getattr(a, k), getattr(b, k)), (rank, numbering, k)
            for k in ("ncells", "ndofs", "nowned", "ninterface_cells", "ndofs_global", "neigh"):
                assert getattr(a, k) == getattr(b, k), (rank, numbering, k)
            assert len(a.send_lists) == len(b.send_lists) == len(a.neigh)
            for la, lb in zip(a.send_lists + a.recv_lists, b.send_lists + b.recv_lists):
                assert np.array_equal(la, lb)
    with pytest.raises(ValueError):
        BoxPartition(2, (1, 2, 2), (2, 1, 1), 1)                       # this rank would have no cells
    with pytest.raises(ValueError):
        BoxPartition(2, (1, 2, 2), (2, 1, 1), 1, native=False)


def test_single_rank_partition_is_trivial(fus, orc):
    from fenicsx_fus_b200.partition import BoxPartition
    p = BoxPartition(3, (3, 2, 2), (1, 1, 1), 0)
    assert p.nowned == p.ndofs and not p.neigh and p.ninterface_cells == 0
    assert np.array_equal(p.dofmap, orc.box_dofmap(3, (3, 2, 2), 1))
    xg, xd = orc.box_mesh((3, 2, 2))
    assert np.array_equal(p.x, xg) and np.array_equal(p.xdofmap, xd)
    assert np.array_equal(p.facets, orc.box_facets((3, 2, 2)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, P, n, pg, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fenicsx_fus_b200.partition import BoxPartition
    from oracle.oracle import Oracle
    orc = Oracle()
    hi = (1.0, 0.7, 0.9)
    p = BoxPartition(P, n, pg, rank, lo=(0, 0, 0), hi=hi)
    G, dJ = orc.geometry(P, p.x, p.xdofmap)
    dphi = orc.dphi(P)
    # global field and per-cell coefficient defined through global ids
    x = np.zeros(p.ndofs)
    x[:p.nowned] = np.sin(0.37 * p.global_key[:p.nowned]) + 0.1
    p.scatter_fwd_host(dist, x)                                  # owner -> ghost
    assert np.allclose(x, np.sin(0.37 * p.global_key) + 0.1, rtol=0, atol=0)
    coeffs = 1.0 + 0.01 * (p.cell_global % 7)
    y = np.zeros(p.ndofs)
    # interface cells, then interior (two launches on the GPU); reverse exchange in between
    ni = p.ninterface_cells
    orc.stiffness_apply(P, p.dofmap[:ni], G[:ni], dphi, coeffs[:ni], x, y)
    orc.stiffness_apply(P, p.dofmap[ni:], G[ni:], dphi, coeffs[ni:], x, y)
    m = np.zeros(p.ndofs)
    orc.mass_apply(P, p.dofmap, dJ, coeffs, np.ones(p.ndofs), m)
    p.scatter_rev_host(dist, y)                                  # ghost -> owner (+=)
    p.scatter_rev_host(dist, m)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), key=p.global_key[:p.nowned],
             y=y[:p.nowned], m=m[:p.nowned])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("P,n,pg", [(3, (4, 3, 2), (2, 1, 1)), (2, (3, 4, 3), (1, 2, 1)),
                                    (2, (5, 2, 2), (3, 1, 1)), (2, (3, 3, 2), (2, 2, 1))])
def test_distributed_operator_gloo_world2(fus, orc, tmp_path, P, n, pg):
    """Two (or three, four) processes over gloo run the partitioned algorithm (scatter_fwd, local
    cells with the interface cells first, scatter_rev) and reproduce the single-domain operator.
    The 3x1x1 slab case has a middle rank with both a lower and an upper neighbour, the 2x2x1 case
    edge neighbours -- topologies beyond the 2-rank ones."""
    import torch.multiprocessing as mp
    port = _free_port()
    world = int(np.prod(pg))
    mp.spawn(_worker, args=(world, port, P, n, pg, str(tmp_path)), nprocs=world, join=True)
    hi = (1.0, 0.7, 0.9)
    xg, xd = orc.box_mesh(n, (0, 0, 0), hi)
    dm = orc.box_dofmap(P, n, 0)                                   # lexicographic: dof id == global key
    nd = dm.max() + 1
    G, dJ = orc.geometry(P, xg, xd)
    x = np.sin(0.37 * np.arange(nd)) + 0.1
    coeffs = 1.0 + 0.01 * (np.arange(dm.shape[0]) % 7)
    y = orc.stiffness_apply(P, dm, G, orc.dphi(P), coeffs, x, np.zeros(nd))
    m = orc.mass_apply(P, dm, dJ, coeffs, np.ones(nd), np.zeros(nd))
    got_y, got_m = np.full(nd, np.nan), np.full(nd, np.nan)
    for r in range(world):
        d = np.load(os.path.join(str(tmp_path), f"rank{r}.npz"))
        got_y[d["key"]] = d["y"]
        got_m[d["key"]] = d["m"]
    assert rel_l2(got_y, y) < 1e-14 and rel_l2(got_m, m) < 1e-14


def _model_worker(rank, world, port, P, n, pg, steps, out_dir):
    """One rank of the partitioned RK4 loop as fus_model_rk4 issues it (csrc/fus_capi.cu): lumped
    mass reduced once, per stage owner->ghost update of the stage input, local cells, boundary
    terms of the local facets, ghost->owner sum of b, epilogue on the owned dofs only.  Oracle
    kernels and host scatters over gloo stand in for the GPU kernels and the halo layer."""
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fenicsx_fus_b200 import capi
    from fenicsx_fus_b200.partition import BoxPartition
    from oracle.oracle import Oracle
    orc = Oracle()
    h = 0.002
    hi = tuple(h * k for k in n)
    p = BoxPartition(P, n, pg, rank, lo=(0, 0, 0), hi=hi)
    nc, nd, no = p.ncells, p.ndofs, p.nowned
    G, dJ = orc.geometry(P, p.x, p.xdofmap)
    dphi = orc.dphi(P)
    c0 = 1450.0 + 10.0 * (p.cell_global % 11)                   # heterogeneous, defined globally
    rho0 = 950.0 + 5.0 * (p.cell_global % 13)
    src, dsrc, absb, bmass = (np.zeros(nd) for _ in range(4))
    assert capi.load().fus_boundary_vectors(
        capi.KINDS["linear"], P, nc, nd, p.x, p.xdofmap, p.dofmap, p.facets.shape[0], p.facets, c0,
        rho0, None, capi.optional(src), capi.optional(dsrc), capi.optional(absb),
        capi.optional(bmass)) == 0
    m = orc.mass_apply(P, p.dofmap, dJ, 1.0 / rho0 / c0 ** 2, np.ones(nd), np.zeros(nd))
    p.scatter_rev_host(dist, m)                                  # once (Linear.hpp:134)
    f0, p0, s0 = 0.5e6, 6.0e4, 1500.0
    w0, dt = 2 * np.pi * f0, 0.5 * np.sqrt(3) * h / (1560.0 * P * P)
    u = np.zeros(nd)
    v = np.zeros(nd)
    u[:no] = 1e3 * np.sin(0.37 * p.global_key[:no])
    v[:no] = 1e9 * np.cos(0.23 * p.global_key[:no])
    a_r, b_r = (0.0, 0.5, 0.5, 1.0), (1 / 6, 1 / 3, 1 / 3, 1 / 6)
    t, tf, dt_full = 0.0, (steps - 0.5) * dt, dt
    while t < tf:                                                # Linear.hpp:270-298
        dt = min(dt, tf - t)                                     # the last step lands on tf
        u0, v0 = u.copy(), v.copy()
        ku, kv = np.zeros(nd), np.zeros(nd)
        for i in range(4):
            un, vn = u0.copy(), v0.copy()
            un[:no] += a_r[i] * dt * ku[:no]
            vn[:no] += a_r[i] * dt * kv[:no]
            tn = t + a_r[i] * dt
            window = 0.5 * (1 - np.cos(f0 * np.pi * tn / 4.0)) if tn < 4.0 / f0 else 1.0
            g = window * p0 * w0 / s0 * np.cos(w0 * tn)
            p.scatter_fwd_host(dist, un)
            p.scatter_fwd_host(dist, vn)
            b = orc.stiffness_apply(P, p.dofmap, G, dphi, -1.0 / rho0, un, np.zeros(nd))
            b += g * src - absb * vn
            p.scatter_rev_host(dist, b)
            ku = vn.copy()
            kv = np.zeros(nd)
            kv[:no] = b[:no] / m[:no]
            u[:no] += b_r[i] * dt * ku[:no]
            v[:no] += b_r[i] * dt * kv[:no]
        t += dt
    np.savez(os.path.join(out_dir, f"model_rank{rank}.npz"), key=p.global_key[:no], u=u[:no],
             v=v[:no], dt=dt_full)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("P,n,pg", [(3, (4, 3, 2), (2, 1, 1)), (2, (5, 2, 2), (3, 1, 1))])
def test_distributed_rk4_model_gloo(fus, orc, tmp_path, P, n, pg):
    """The partitioned RK4 stage flow (what fus_model_rk4 issues on N GPUs) over gloo with oracle
    kernels: heterogeneous linear model from a random state, against the oracle's single-domain
    LinearSpectral3D."""
    import torch.multiprocessing as mp
    world, steps = int(np.prod(pg)), 4
    mp.spawn(_model_worker, args=(world, _free_port(), P, n, pg, steps, str(tmp_path)), nprocs=world,
             join=True)
    h = 0.002
    xg, xd = orc.box_mesh(n, (0, 0, 0), tuple(h * k for k in n))
    dm = orc.box_dofmap(P, n, 0)                                   # lexicographic: dof id == global key
    nd, nc = dm.max() + 1, dm.shape[0]
    G, dJ = orc.geometry(P, xg, xd)
    facets = orc.box_facets(n)
    fn, fs = orc.facet_data(P, xg, xd, facets)
    cg = np.arange(nc)
    c0, rho0 = 1450.0 + 10.0 * (cg % 11), 950.0 + 5.0 * (cg % 13)
    om = orc.model("linear", P, nd, dm, G, dJ, orc.dphi(P), c0, rho0, None, None, facets, fn, fs,
                   0.5e6, 6.0e4, 1500.0)
    key = np.arange(nd)
    u, v = 1e3 * np.sin(0.37 * key), 1e9 * np.cos(0.23 * key)
    got_u, got_v = np.full(nd, np.nan), np.full(nd, np.nan)
    for r in range(world):
        d = np.load(os.path.join(str(tmp_path), f"model_rank{r}.npz"))
        got_u[d["key"]], got_v[d["key"]], dt = d["u"], d["v"], float(d["dt"])
    assert om.rk4(0.0, (steps - 0.5) * dt, dt, u, v) == steps
    assert rel_l2(got_u, u) < 1e-12 and rel_l2(got_v, v) < 1e-12


# ---- unstructured meshes ----------------------------------------------------------------------
def _ref_mesh():
    from fenicsx_fus_b200.unstructured import HexMesh
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_mesh_hex6312.npz"))
    return HexMesh(g["geometry"], g["topology_vtk"][:, (0, 1, 3, 2, 4, 5, 7, 6)], g["facet_quads"],
                   g["facet_values"], reorder="morton")


@pytest.mark.parametrize("P,nranks", [(1, 3), (2, 4)])
def test_hex_partition_invariants(fus, P, nranks):
    from fenicsx_fus_b200.partition import HexPartition
    from fenicsx_fus_b200.unstructured import HexFunctionSpace
    m = _ref_mesh()
    V = HexFunctionSpace(m, P)
    parts = [HexPartition(m, P, nranks, r, space=V) for r in range(nranks)]
    owned = np.concatenate([p.global_key[:p.nowned] for p in parts])
    assert np.array_equal(np.sort(owned), np.arange(V.ndofs))          # every dof owned exactly once
    assert np.array_equal(np.sort(np.concatenate([p.cell_global for p in parts])), np.arange(m.ncells))
    assert sum(p.facets.shape[0] for p in parts) == m.facets.shape[0]
    for p in parts:
        assert sorted(np.unique(p.dofmap)) == list(range(p.ndofs))
        for q, s, r in zip(p.neigh, p.send_lists, p.recv_lists):
            o = parts[q]
            k = o.neigh.index(p.rank)
            assert np.array_equal(p.global_key[s], o.global_key[o.recv_lists[k]])
            assert np.array_equal(p.global_key[r], o.global_key[o.send_lists[k]])
            assert np.all(s < p.nowned) and np.all(r >= p.nowned)
        rec = np.concatenate(p.recv_lists) if p.neigh else np.zeros(0, int)
        assert np.array_equal(np.sort(rec), np.arange(p.nowned, p.ndofs))
        shared = np.zeros(p.ndofs, bool)
        shared[p.nowned:] = True
        for s in p.send_lists:
            shared[s] = True
        touch = shared[p.dofmap].any(1)
        assert touch[:p.ninterface_cells].all() and not touch[p.ninterface_cells:].any()
        # local geometry is the global one restricted to the local cells
        assert np.array_equal(p.x[p.xdofmap], m.x[m.xdofmap[p.cell_global]])


def _worker_hex(rank, world, port, P, out_dir):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from fenicsx_fus_b200.partition import HexPartition
    from oracle.oracle import Oracle
    orc = Oracle()
    m = _ref_mesh()
    p = HexPartition(m, P, world, rank)
    G, dJ = orc.geometry(P, p.x, p.xdofmap)
    x = np.zeros(p.ndofs)
    x[:p.nowned] = np.sin(0.37 * p.global_key[:p.nowned]) + 0.1
    p.scatter_fwd_host(dist, x)
    assert np.array_equal(x, np.sin(0.37 * p.global_key) + 0.1)
    coeffs = 1.0 + 0.01 * (p.cell_global % 7)
    y = np.zeros(p.ndofs)
    ni = p.ninterface_cells
    orc.stiffness_apply(P, p.dofmap[:ni], G[:ni], orc.dphi(P), coeffs[:ni], x, y)
    orc.stiffness_apply(P, p.dofmap[ni:], G[ni:], orc.dphi(P), coeffs[ni:], x, y)
    p.scatter_rev_host(dist, y)
    np.savez(os.path.join(out_dir, f"hrank{rank}.npz"), key=p.global_key[:p.nowned], y=y[:p.nowned])
    dist.barrier()
    dist.destroy_process_group()


def test_distributed_operator_gloo_world2_unstructured(fus, orc, tmp_path):
    """The partitioned algorithm on the reference's unstructured test mesh, two gloo ranks."""
    import torch.multiprocessing as mp
    from fenicsx_fus_b200.unstructured import HexFunctionSpace
    P = 2
    mp.spawn(_worker_hex, args=(2, _free_port(), P, str(tmp_path)), nprocs=2, join=True)
    m = _ref_mesh()
    V = HexFunctionSpace(m, P)
    G, _ = orc.geometry(P, m.x, m.xdofmap)
    x = np.sin(0.37 * np.arange(V.ndofs)) + 0.1
    coeffs = 1.0 + 0.01 * (np.arange(m.ncells) % 7)
    y = orc.stiffness_apply(P, V.dofmap, G, orc.dphi(P), coeffs, x, np.zeros(V.ndofs))
    got = np.full(V.ndofs, np.nan)
    for r in range(2):
        d = np.load(os.path.join(str(tmp_path), f"hrank{r}.npz"))
        got[d["key"]] = d["y"]
    assert rel_l2(got, y) < 1e-14
