"""Mesh partition of a structured box over a Px x Py x Pz process grid (one block per GPU).

Mirrors the reference's distributed model: cells are partitioned without
ghost cells (GhostMode::none, BM7-SC1/main.cpp:58-59), a dof on a partition interface is owned by
exactly one rank and ghosted on the others, local vectors hold the owned entries first and the
ghosts after them (relied on by kernels::axpy, Linear.hpp:35), and each operator application is
preceded by an owner->ghost update and followed by a ghost->owner sum
(scatter_fwd / scatter_rev, Linear.hpp:196-206).

Ownership rule: an interface node belongs to the block with the lowest grid coordinates among
those sharing it, so the ghosts of a block sit on its lower faces.
Cells that touch a shared dof ("interface cells") are ordered first so that the reverse exchange
can overlap with the remaining interior cells.
"""
import ctypes as C
import itertools

import numpy as np

from . import capi


def _split(n, parts, r):
    """Balanced contiguous split of n cells in `parts`: [start, end) of part r."""
    base, rem = divmod(n, parts)
    start = r * base + min(r, rem)
    return start, start + base + (1 if r < rem else 0)


def peer_byte_offsets(q_neigh, q_soff, q_roff, q_layout, my_rank):
    """Where this rank's data go inside neighbour q's mailbox (bytes): {forward-u run, forward-v run,
    reverse run, forward flag, reverse flag, ready flag}, from q's neighbour list, offset tables and
    mailbox layout (fus_halo_peer_offsets in include/fus_b200.h)."""
    j = list(q_neigh).index(my_rank)
    out = np.zeros(6, dtype=np.int64)
    capi.check(capi.load().fus_halo_peer_offsets(
        np.ascontiguousarray(q_layout, dtype=np.int64), np.ascontiguousarray(q_soff, dtype=np.int64),
        np.ascontiguousarray(q_roff, dtype=np.int64), j, out), "fus_halo_peer_offsets")
    return tuple(int(v) for v in out)


class _PartitionBase:
    """What the operator / model / halo layers need from a partition of any mesh:
    x, xdofmap, dofmap (local numbering, owned first), ndofs, nowned, ncells, facets, P, N, rank,
    nranks, neigh, send_lists, recv_lists, ninterface_cells, n_local (box only)."""
    n_local = None

    # -------------------------------------------------------------------------------------------
    def function_space(self, device=0, lean=False):
        """FunctionSpace-like object over the local block for the operator/model classes."""
        import types

        from . import Context
        mesh = types.SimpleNamespace(x=self.x, xdofmap=self.xdofmap, facets=self.facets,
                                     ncells=self.ncells, n=self.n_local)
        V = types.SimpleNamespace(mesh=mesh, P=self.P, N=self.N, ndofs=self.ndofs,
                                  nowned=self.nowned, dofmap=self.dofmap, _ctx=None)

        def context(dev=device):
            if V._ctx is None:
                V._ctx = Context.from_mesh(V, dev, nowned=self.nowned, lean=lean)
            return V._ctx
        V.context = context
        return V

    def halo_arrays(self):
        nn = len(self.neigh)
        neigh = np.array(self.neigh, dtype=np.int32)
        soff = np.zeros(nn + 1, dtype=np.int64)
        roff = np.zeros(nn + 1, dtype=np.int64)
        for k in range(nn):
            soff[k + 1] = soff[k] + self.send_lists[k].size
            roff[k + 1] = roff[k] + self.recv_lists[k].size
        sidx = np.concatenate(self.send_lists) if nn else np.zeros(0, np.int32)
        ridx = np.concatenate(self.recv_lists) if nn else np.zeros(0, np.int32)
        return (neigh, soff, np.ascontiguousarray(sidx, dtype=np.int32), roff,
                np.ascontiguousarray(ridx, dtype=np.int32))

    def setup_halo(self, ctx, dist):
        """Create the NCCL communicator of the context: rank 0 makes the unique id, it is
        broadcast with torch.distributed, every rank calls fus_halo_setup."""
        import torch
        lib = capi.load()
        uid = np.zeros(128, dtype=np.uint8)
        if self.rank == 0:
            capi.check(lib.fus_comm_unique_id(uid.ctypes.data_as(C.c_void_p)), "fus_comm_unique_id")
        t = torch.from_numpy(uid).cuda() if dist.get_backend() == "nccl" else torch.from_numpy(uid)
        dist.broadcast(t, src=0)
        uid = t.cpu().numpy().copy()
        neigh, soff, sidx, roff, ridx = self.halo_arrays()
        p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        capi.check(lib.fus_halo_setup(ctx.h, self.rank, self.nranks, p(uid), len(self.neigh),
                                      p(neigh), p(soff), p(sidx), p(roff), p(ridx),
                                      self.ninterface_cells), "fus_halo_setup")

    def connect_peers(self, ctx, dist):
        """Switch the exchanges inside rk4 to the fused peer transport: every rank exports its
        mailbox (CUDA IPC handle + layout), all ranks gather them, each rank opens its neighbours'
        mailboxes.  Returns False (and leaves NCCL in place) if IPC is unavailable or the local
        numbering does not have the shape the fused kernels need."""
        lib = capi.load()
        handle = np.zeros(64, dtype=np.uint8)
        layout = np.zeros(6, dtype=np.int64)
        p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        rc = lib.fus_halo_peer_export(ctx.h, p(handle), p(layout), None)
        neigh, soff, _, roff, _ = self.halo_arrays()
        mine = dict(ok=(rc == 0), handle=handle.tobytes(), layout=layout.tolist(),
                    neigh=[int(q) for q in neigh], soff=soff.tolist(), roff=roff.tolist())
        allinfo = [None] * self.nranks
        dist.all_gather_object(allinfo, mine)
        if not all(i["ok"] for i in allinfo):
            return False
        nn = len(self.neigh)
        handles = np.zeros((max(nn, 1), 64), dtype=np.uint8)
        boff = np.zeros((max(nn, 1), 6), dtype=np.int64)
        for k, q in enumerate(self.neigh):
            info = allinfo[q]
            handles[k] = np.frombuffer(info["handle"], dtype=np.uint8)
            boff[k] = peer_byte_offsets(info["neigh"], info["soff"], info["roff"], info["layout"],
                                        self.rank)
        rc = lib.fus_halo_peer_connect(ctx.h, p(handles), p(boff))
        flags = [None] * self.nranks
        dist.all_gather_object(flags, rc == 0)
        if not any(flags):
            return False
        if not all(flags):
            raise capi.FusError("peer transport connected on some ranks only: "
                                + lib.fus_last_error().decode(errors="replace"))
        return True

    # ---- host-side halo (CPU tests over gloo; the device path is fus_halo.cu) -----------------
    def scatter_fwd_host(self, dist, x):
        import torch
        ops, bufs = [], []
        for q, s, r in zip(self.neigh, self.send_lists, self.recv_lists):
            if s.size:
                ops.append(dist.P2POp(dist.isend, torch.from_numpy(x[s].copy()), q))
            if r.size:
                b = torch.zeros(r.size, dtype=torch.float64)
                bufs.append((r, b))
                ops.append(dist.P2POp(dist.irecv, b, q))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for r, b in bufs:
            x[r] = b.numpy()

    def scatter_rev_host(self, dist, x):
        import torch
        ops, bufs = [], []
        for q, s, r in zip(self.neigh, self.send_lists, self.recv_lists):
            if r.size:
                ops.append(dist.P2POp(dist.isend, torch.from_numpy(x[r].copy()), q))
            if s.size:
                b = torch.zeros(s.size, dtype=torch.float64)
                bufs.append((s, b))
                ops.append(dist.P2POp(dist.irecv, b, q))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        for s, b in bufs:
            np.add.at(x, s, b.numpy())


class BoxPartition(_PartitionBase):
    """One rank's block of the global box.  The index work runs in the library
    (fus_box_partition_create, csrc/fus_partition.cpp); `native=False` selects the numpy
    implementation below, kept as the cross-check the tests compare it with array by array."""

    def __init__(self, P, n_global, pgrid, rank, lo=(0.0, 0.0, 0.0), hi=(1.0, 1.0, 1.0),
                 numbering=1, native=True):
        lib = capi.load()
        self.P, self.N = int(P), int(P) + 1
        self.n_global = tuple(int(v) for v in n_global)
        self.pgrid = tuple(int(v) for v in pgrid)
        self.rank = int(rank)
        self.nranks = int(np.prod(self.pgrid))
        if native:
            self._init_native(lib, lo, hi, numbering)
            return
        Px, Py, Pz = self.pgrid
        self.rcoord = (rank // (Py * Pz), (rank // Pz) % Py, rank % Pz)
        rng = [_split(self.n_global[d], self.pgrid[d], self.rcoord[d]) for d in range(3)]
        self.cell_lo = np.array([r[0] for r in rng])
        self.n_local = np.array([r[1] - r[0] for r in rng], dtype=np.int32)
        if np.any(self.n_local < 1):
            raise ValueError("a rank has no cells")
        nl = self.n_local
        P_, N = self.P, self.N
        lo, hi = np.asarray(lo, np.float64), np.asarray(hi, np.float64)
        self.has_lower = [self.rcoord[d] > 0 for d in range(3)]
        self.has_upper = [self.rcoord[d] < self.pgrid[d] - 1 for d in range(3)]
        M = [self.n_global[d] * P_ + 1 for d in range(3)]          # global node grid
        self.ndofs_global = int(M[0]) * int(M[1]) * int(M[2])

        # ---- local geometry: vertex coordinates from the GLOBAL formula (bitwise identical to the
        # unpartitioned box) -------------------------------------------------------------------
        nv = nl + 1
        ax = [lo[d] + (hi[d] - lo[d]) * (self.cell_lo[d] + np.arange(nv[d])) / self.n_global[d]
              for d in range(3)]
        X, Y, Z = np.meshgrid(*ax, indexing="ij")
        self.x = np.ascontiguousarray(np.stack([X, Y, Z], -1).reshape(-1, 3))
        xd = np.zeros((int(nl.prod()), 8), dtype=np.int32)
        dummy = np.zeros_like(self.x)
        capi.check(lib.fus_box_mesh(nl, np.zeros(3), np.ones(3), dummy, xd), "fus_box_mesh")

        # ---- raw local dofmap (cell-blocked numbering of the local node grid) ---------------
        ncl = int(nl.prod())
        Nd = N ** 3
        raw = np.zeros((ncl, Nd), dtype=np.int32)
        capi.check(lib.fus_box_dofmap(P_, nl, numbering, raw), "fus_box_dofmap")
        nraw = int(lib.fus_box_num_dofs(P_, nl))
        # local grid coordinates of every raw dof
        pos = np.array([0, P_] + list(range(1, P_)), dtype=np.int32)   # Basix node -> offset
        # cell index c = (cx*ny + cy)*nz + cz
        cid = np.arange(ncl, dtype=np.int32)
        cxs, cys, czs = cid // (nl[1] * nl[2]), (cid // nl[2]) % nl[1], cid % nl[2]
        del cid
        i0, i1, i2 = np.meshgrid(pos, pos, pos, indexing="ij")
        g = [np.zeros(nraw, dtype=np.int32) for _ in range(3)]   # int32 throughout: 1e9-dof runs
        for d, (cc, ii) in enumerate(((cxs, i0), (cys, i1), (czs, i2))):
            gd = (cc[:, None] * np.int32(P_) + ii.reshape(1, -1).astype(np.int32)).reshape(-1)
            g[d][raw.reshape(-1)] = gd
            del gd
        top = [int(nl[d]) * P_ for d in range(3)]
        ghost = np.zeros(nraw, dtype=bool)
        owner_rank = np.zeros(nraw, dtype=np.int32)
        mult = (Py * Pz, Pz, 1)
        for d in range(3):
            on_low = (g[d] == 0) & self.has_lower[d]
            ghost |= on_low
            owner_rank += np.int32(mult[d]) * (np.int32(self.rcoord[d]) - on_low.astype(np.int32))
            del on_low
        key = (g[0].astype(np.int64) + self.cell_lo[0] * P_) * M[1]
        key += g[1] + self.cell_lo[1] * P_
        key *= M[2]
        key += g[2] + self.cell_lo[2] * P_                           # global node id

        # ---- local numbering: owned dofs that a neighbour ghosts first, then the other owned dofs
        # (raw order within each), then ghosts grouped by owner, by global key --------------------
        shared_owned = np.zeros(nraw, dtype=bool)
        for d in range(3):
            if self.has_upper[d]:
                shared_owned |= g[d] == top[d]
        shared_owned &= ~ghost
        owned_raw = np.concatenate([np.flatnonzero(shared_owned), np.flatnonzero(~ghost & ~shared_owned)])
        ghost_raw = np.flatnonzero(ghost)
        order = np.lexsort((key[ghost_raw], owner_rank[ghost_raw]))
        ghost_raw = ghost_raw[order]
        self.nowned = int(owned_raw.size)
        self.nshared = int(shared_owned.sum())
        self.ndofs = nraw
        new_of_raw = np.empty(nraw, dtype=np.int32)
        new_of_raw[owned_raw] = np.arange(self.nowned, dtype=np.int32)
        new_of_raw[ghost_raw] = self.nowned + np.arange(ghost_raw.size, dtype=np.int32)
        self.global_key = np.empty(nraw, dtype=np.int64)
        self.global_key[new_of_raw] = key
        dofmap = new_of_raw[raw]

        # ---- halo lists ------------------------------------------------------------------------
        # recv: ghosts, already grouped by owner rank and sorted by key
        g_owner = owner_rank[ghost_raw]
        neigh = {}
        for q in np.unique(g_owner):
            sel = np.flatnonzero(g_owner == q)
            neigh.setdefault(int(q), {})["recv"] = (self.nowned + sel).astype(np.int32)
        # send: for every upper neighbour r+delta, my owned nodes on the top planes of delta
        owned_mask_new = np.zeros(nraw, dtype=bool)
        owned_mask_new[:self.nowned] = True
        gn = [np.empty(nraw, dtype=np.int32) for _ in range(3)]
        for d in range(3):
            gn[d][new_of_raw] = g[d]
        del g, key
        for delta in itertools.product((0, 1), repeat=3):
            if not any(delta):
                continue
            if any(delta[d] and not self.has_upper[d] for d in range(3)):
                continue
            q = ((self.rcoord[0] + delta[0]) * Py + self.rcoord[1] + delta[1]) * Pz \
                + self.rcoord[2] + delta[2]
            m = owned_mask_new.copy()
            for d in range(3):
                if delta[d]:
                    m &= gn[d] == top[d]
            idx = np.flatnonzero(m)
            idx = idx[np.argsort(self.global_key[idx], kind="stable")]
            neigh.setdefault(int(q), {})["send"] = idx.astype(np.int32)
        self.neigh = sorted(neigh)
        e = np.zeros(0, dtype=np.int32)
        self.send_lists = [neigh[q].get("send", e) for q in self.neigh]
        self.recv_lists = [neigh[q].get("recv", e) for q in self.neigh]

        # ---- interface cells first ------------------------------------------------------------
        shared = np.zeros(nraw, dtype=bool)
        shared[self.nowned:] = True
        for s in self.send_lists:
            shared[s] = True
        is_iface = shared[dofmap].any(axis=1)
        perm = np.concatenate([np.flatnonzero(is_iface), np.flatnonzero(~is_iface)])
        self.ninterface_cells = int(is_iface.sum())
        self.dofmap = np.ascontiguousarray(dofmap[perm])
        self.xdofmap = np.ascontiguousarray(xd[perm])
        self.ncells = ncl
        gcx, gcy, gcz = cxs + self.cell_lo[0], cys + self.cell_lo[1], czs + self.cell_lo[2]
        self.cell_global = (((gcx * self.n_global[1]) + gcy) * self.n_global[2] + gcz)[perm]
        inv = np.empty(ncl, dtype=np.int64)
        inv[perm] = np.arange(ncl)

        # ---- exterior facets of the GLOBAL box that belong to local cells ----------------------
        nf = lib.fus_box_facets(nl, None)
        f = np.zeros((nf, 3), dtype=np.int32)
        lib.fus_box_facets(nl, f.ctypes.data_as(C.c_void_p))
        fdir = np.array([2, 1, 0, 0, 1, 2])[f[:, 1]]
        fside = np.array([0, 0, 0, 1, 1, 1])[f[:, 1]]
        keep = np.ones(nf, dtype=bool)
        for d in range(3):
            keep &= ~((fdir == d) & (fside == 0) & self.has_lower[d])
            keep &= ~((fdir == d) & (fside == 1) & self.has_upper[d])
        f = f[keep]
        f[:, 0] = inv[f[:, 0]]
        self.facets = np.ascontiguousarray(f)


    def _init_native(self, lib, lo, hi, numbering):
        h = C.c_void_p()
        ng = np.array(self.n_global, dtype=np.int32)
        pg = np.array(self.pgrid, dtype=np.int32)
        rc = lib.fus_box_partition_create(self.P, ng, pg, self.rank, int(numbering), C.byref(h))
        if rc != 0:
            msg = lib.fus_last_error().decode(errors="replace")
            if "no cells" in msg:
                raise ValueError("a rank has no cells")
            raise capi.FusError(f"fus_box_partition_create failed with code {rc}: {msg}")
        try:
            sizes = np.zeros(9, dtype=np.int64)
            self.n_local = np.zeros(3, dtype=np.int32)
            cell_lo = np.zeros(3, dtype=np.int32)
            capi.check(lib.fus_box_partition_info(h, sizes, self.n_local, cell_lo), "partition info")
            ncells, ndofs, nowned, nf, nn, ns, nr, nif, ndg = (int(v) for v in sizes)
            self.cell_lo = cell_lo.astype(np.int64)
            Px, Py, Pz = self.pgrid
            self.rcoord = (self.rank // (Py * Pz), (self.rank // Pz) % Py, self.rank % Pz)
            self.has_lower = [self.rcoord[d] > 0 for d in range(3)]
            self.has_upper = [self.rcoord[d] < self.pgrid[d] - 1 for d in range(3)]
            self.ncells, self.ndofs, self.nowned = ncells, ndofs, nowned
            self.ninterface_cells, self.ndofs_global = nif, ndg
            self.dofmap = np.zeros((ncells, self.N ** 3), dtype=np.int32)
            self.xdofmap = np.zeros((ncells, 8), dtype=np.int32)
            self.cell_global = np.zeros(ncells, dtype=np.int64)
            self.global_key = np.zeros(ndofs, dtype=np.int64)
            self.facets = np.zeros((nf, 3), dtype=np.int32)
            neigh = np.zeros(nn, dtype=np.int32)
            soff, roff = np.zeros(nn + 1, dtype=np.int64), np.zeros(nn + 1, dtype=np.int64)
            sidx, ridx = np.zeros(ns, dtype=np.int32), np.zeros(nr, dtype=np.int32)
            p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
            capi.check(lib.fus_box_partition_arrays(
                h, p(self.dofmap), p(self.xdofmap), p(self.cell_global), p(self.global_key),
                p(self.facets), p(neigh), p(soff), p(sidx), p(roff), p(ridx)), "partition arrays")
        finally:
            lib.fus_box_partition_destroy(h)
        self.neigh = [int(q) for q in neigh]
        self.send_lists = [sidx[soff[k]:soff[k + 1]] for k in range(nn)]
        self.recv_lists = [ridx[roff[k]:roff[k + 1]] for k in range(nn)]
        self.nshared = int(np.unique(sidx).size)
        # vertex coordinates from the GLOBAL formula (bitwise identical to the unpartitioned box)
        lo, hi = np.asarray(lo, np.float64), np.asarray(hi, np.float64)
        nv = self.n_local + 1
        ax = [lo[d] + (hi[d] - lo[d]) * (self.cell_lo[d] + np.arange(nv[d])) / self.n_global[d]
              for d in range(3)]
        X, Y, Z = np.meshgrid(*ax, indexing="ij")
        self.x = np.ascontiguousarray(np.stack([X, Y, Z], -1).reshape(-1, 3))


class HexPartition(_PartitionBase):
    """Partition of an unstructured HexMesh over `nranks` GPUs.

    Cells are dealt out as contiguous chunks of the (Morton-ordered) cell list; a dof shared by
    several ranks is owned by the lowest of them.  Every rank builds the global numbering
    (fine for meshes that fit one host) and keeps its own part: owned dofs first in order of first
    appearance, then ghosts grouped by owner and sorted by global id; interface cells first."""

    def __init__(self, mesh, P, nranks, rank, space=None):
        from .unstructured import HexFunctionSpace
        self.P, self.N = int(P), int(P) + 1
        self.rank, self.nranks = int(rank), int(nranks)
        V = space if space is not None else HexFunctionSpace(mesh, P)
        nc = mesh.ncells
        cell_rank = (np.arange(nc, dtype=np.int64) * self.nranks // nc).astype(np.int32)
        mine = np.flatnonzero(cell_rank == self.rank)
        if mine.size == 0:
            raise ValueError("a rank has no cells")
        gdm = V.dofmap                                               # global ids (nc, Nd)
        Nd = gdm.shape[1]
        # which ranks touch which dof
        pairs = np.unique(np.stack([gdm.reshape(-1).astype(np.int64),
                                    np.repeat(cell_rank, Nd).astype(np.int64)], axis=1), axis=0)
        owner = np.full(V.ndofs, self.nranks, dtype=np.int32)
        np.minimum.at(owner, pairs[:, 0], pairs[:, 1].astype(np.int32))
        touched_by_me = pairs[pairs[:, 1] == self.rank, 0]
        # local numbering
        loc = gdm[mine]
        flat = loc.reshape(-1)
        uniq, first = np.unique(flat, return_index=True)
        uniq = uniq[np.argsort(first, kind="stable")]                 # first appearance order
        is_owned = owner[uniq] == self.rank
        owned = uniq[is_owned]
        # owned dofs that another rank touches ("shared": ghosted there) come first
        nranks_touching = np.bincount(pairs[:, 0], minlength=V.ndofs)
        sh = nranks_touching[owned] > 1
        owned = np.concatenate([owned[sh], owned[~sh]])
        self.nshared = int(sh.sum())
        ghosts = uniq[~is_owned]
        ghosts = ghosts[np.lexsort((ghosts, owner[ghosts]))]
        self.nowned, self.ndofs = int(owned.size), int(uniq.size)
        self.ndofs_global = int(V.ndofs)
        self.global_key = np.concatenate([owned, ghosts]).astype(np.int64)
        g2l = np.full(V.ndofs, -1, dtype=np.int64)
        g2l[self.global_key] = np.arange(self.ndofs)
        dofmap = g2l[loc]
        # halo lists
        neigh = {}
        for q in np.unique(owner[ghosts]):
            sel = ghosts[owner[ghosts] == q]                         # sorted by global id
            neigh.setdefault(int(q), {})["recv"] = g2l[sel].astype(np.int32)
        owned_mask = np.zeros(V.ndofs, dtype=bool)
        owned_mask[owned] = True
        for q in range(self.nranks):
            if q == self.rank:
                continue
            tq = pairs[pairs[:, 1] == q, 0]
            sel = np.sort(tq[owned_mask[tq]])
            if sel.size:
                neigh.setdefault(int(q), {})["send"] = g2l[sel].astype(np.int32)
        self.neigh = sorted(neigh)
        e = np.zeros(0, dtype=np.int32)
        self.send_lists = [neigh[q].get("send", e) for q in self.neigh]
        self.recv_lists = [neigh[q].get("recv", e) for q in self.neigh]
        # interface cells first
        shared = np.zeros(self.ndofs, dtype=bool)
        shared[self.nowned:] = True
        for s_ in self.send_lists:
            shared[s_] = True
        is_iface = shared[dofmap].any(axis=1)
        perm = np.concatenate([np.flatnonzero(is_iface), np.flatnonzero(~is_iface)])
        self.ninterface_cells = int(is_iface.sum())
        self.dofmap = np.ascontiguousarray(dofmap[perm], dtype=np.int32)
        self.cell_global = mine[perm]
        self.ncells = int(mine.size)
        # geometry: keep only the vertices the local cells use
        xd = mesh.xdofmap[self.cell_global]
        vu, vinv = np.unique(xd.reshape(-1), return_inverse=True)
        self.x = np.ascontiguousarray(mesh.x[vu])
        self.xdofmap = np.ascontiguousarray(vinv.reshape(xd.shape), dtype=np.int32)
        # exterior facets of local cells
        inv = np.full(nc, -1, dtype=np.int64)
        inv[self.cell_global] = np.arange(self.ncells)
        f = mesh.facets[cell_rank[mesh.facets[:, 0]] == self.rank].copy()
        f[:, 0] = inv[f[:, 0]]
        self.facets = np.ascontiguousarray(f, dtype=np.int32)
        del touched_by_me
