// fus_partition.cpp -- native partitioner of the structured box over a Px x Py x Pz process grid.
//
// Host-side set-up (runs once per rank), the counterpart of what DOLFINx builds in C++ for the
// reference: a cell partition without ghost cells (GhostMode::none, BM7-SC1/main.cpp:58-59), the
// local dof numbering "owned entries first, then ghosts" (common::IndexMap; relied on by
// kernels::axpy, Linear.hpp:35) and the neighbour lists behind la::Vector::scatter_fwd /
// scatter_rev (Linear.hpp:196-206).  Rules (same as fenicsx-fus_b200/partition.py, which keeps a
// numpy implementation that the tests compare against array by array):
//   * an interface node is owned by the block with the lowest grid coordinates sharing it, so a
//     block's ghosts sit on its lower faces;
//   * local numbering: owned dofs that some neighbour ghosts ("shared") first, then the other owned
//     dofs, both in the order of the block's own cell-blocked numbering, then ghosts grouped by
//     owner rank and sorted by global node id (so that the fused epilogue + exchange kernels find
//     the dofs they send at [0, nshared) and every neighbour's ghosts in one contiguous run);
//   * receive list of neighbour q: its ghosts in that order; send list to an upper neighbour: the
//     owned nodes on the shared top planes, sorted by global node id (matching orders on both sides);
//   * cells touching a shared dof come first (the reverse exchange overlaps with the rest).
#include "fus_internal.hpp"

#include <algorithm>
#include <cstring>
#include <new>
#include <vector>

struct fus_partition {
  int P = 0, rank = 0;
  int n_global[3] = {0, 0, 0}, pgrid[3] = {1, 1, 1}, rcoord[3] = {0, 0, 0};
  int32_t n_local[3] = {0, 0, 0}, cell_lo[3] = {0, 0, 0};
  int64_t ncells = 0, ndofs = 0, nowned = 0, nshared = 0, ninterface = 0, ndofs_global = 0;
  std::vector<int32_t> dofmap, xdofmap, facets, neigh, send_idx, recv_idx;
  std::vector<int64_t> cell_global, global_key, send_off, recv_off;
};

namespace {
void split(int n, int parts, int r, int& start, int& end) { // balanced contiguous split
  const int base = n / parts, rem = n % parts;
  start = r * base + std::min(r, rem);
  end = start + base + (r < rem ? 1 : 0);
}
} // namespace

extern "C" {

int fus_box_partition_create(int P, const int n_global[3], const int pgrid[3], int rank,
                             int numbering, fus_partition** out) {
  using namespace fus;
  if (out)
    *out = nullptr;
  if (!out || !n_global || !pgrid || P < 1 || numbering < 0 || numbering > 1) {
    set_error("fus_box_partition_create: bad argument");
    return FUS_ERR_ARG;
  }
  const int nranks = pgrid[0] * pgrid[1] * pgrid[2];
  if (pgrid[0] < 1 || pgrid[1] < 1 || pgrid[2] < 1 || rank < 0 || rank >= nranks) {
    set_error("fus_box_partition_create: rank %d outside the %dx%dx%d process grid", rank,
              pgrid[0], pgrid[1], pgrid[2]);
    return FUS_ERR_ARG;
  }
  fus_partition* p = new (std::nothrow) fus_partition();
  if (!p)
    return FUS_ERR_ARG;
  try {
    p->P = P;
    p->rank = rank;
    const int Py = pgrid[1], Pz = pgrid[2];
    p->rcoord[0] = rank / (Py * Pz);
    p->rcoord[1] = (rank / Pz) % Py;
    p->rcoord[2] = rank % Pz;
    bool has_lower[3], has_upper[3];
    int64_t M[3];
    for (int d = 0; d < 3; ++d) {
      p->n_global[d] = n_global[d];
      p->pgrid[d] = pgrid[d];
      int s, e;
      split(n_global[d], pgrid[d], p->rcoord[d], s, e);
      p->cell_lo[d] = s;
      p->n_local[d] = e - s;
      if (e - s < 1) {
        delete p;
        set_error("fus_box_partition_create: a rank has no cells");
        return FUS_ERR_ARG;
      }
      has_lower[d] = p->rcoord[d] > 0;
      has_upper[d] = p->rcoord[d] < pgrid[d] - 1;
      M[d] = (int64_t)n_global[d] * P + 1;
    }
    p->ndofs_global = M[0] * M[1] * M[2];
    const int nl[3] = {p->n_local[0], p->n_local[1], p->n_local[2]};
    const int N = P + 1, Nd = N * N * N;
    const int64_t ncl = (int64_t)nl[0] * nl[1] * nl[2];
    const int64_t nraw = box_num_dofs(P, nl);
    if (nraw > INT32_MAX) {
      delete p;
      set_error("local dof count exceeds int32");
      return FUS_ERR_UNSUPPORTED;
    }
    p->ncells = ncl;
    p->ndofs = nraw;

    // raw local dofmap and the local grid coordinates of every raw dof
    std::vector<int32_t> raw((size_t)ncl * Nd);
    int rc = box_dofmap(P, nl, numbering, raw.data());
    if (rc != FUS_OK) {
      delete p;
      return rc;
    }
    std::vector<int> pos(N);
    pos[0] = 0;
    pos[1] = P;
    for (int i = 2; i < N; ++i)
      pos[i] = i - 1;
    std::vector<int32_t> g[3];
    for (int d = 0; d < 3; ++d)
      g[d].assign((size_t)nraw, 0);
    {
      // (threads may write the same entry of g: always with the same value)
#pragma omp parallel for schedule(static)
      for (int cx = 0; cx < nl[0]; ++cx)
        for (int cy = 0; cy < nl[1]; ++cy)
          for (int cz = 0; cz < nl[2]; ++cz) {
            const int64_t c = ((int64_t)cx * nl[1] + cy) * nl[2] + cz;
            const int32_t* row = raw.data() + c * Nd;
            for (int a = 0; a < N; ++a)
              for (int b = 0; b < N; ++b)
                for (int e = 0; e < N; ++e) {
                  const int32_t r = row[(a * N + b) * N + e];
                  g[0][r] = cx * P + pos[a];
                  g[1][r] = cy * P + pos[b];
                  g[2][r] = cz * P + pos[e];
                }
          }
    }
    const int mult[3] = {Py * Pz, Pz, 1};
    auto key_of = [&](int64_t r) {
      return ((g[0][r] + (int64_t)p->cell_lo[0] * P) * M[1] + (g[1][r] + (int64_t)p->cell_lo[1] * P))
                 * M[2]
             + (g[2][r] + (int64_t)p->cell_lo[2] * P);
    };

    // ownership and the new local numbering
    struct Ghost {
      int32_t owner;
      int64_t key;
      int32_t raw;
    };
    std::vector<Ghost> ghosts;
    std::vector<int32_t> new_of_raw((size_t)nraw);
    const int top[3] = {nl[0] * P, nl[1] * P, nl[2] * P};
    // an owned node on a top plane that an upper neighbour shares is ghosted there: those come
    // first in the local numbering (the fused epilogue + exchange handles [0, nshared) specially)
    auto shared_owned = [&](int64_t r) {
      for (int d = 0; d < 3; ++d)
        if (g[d][r] == top[d] && has_upper[d])
          return true;
      return false;
    };
    int32_t nowned = 0, nshared = 0;
    std::vector<char> kind((size_t)nraw, 0); // 0 private, 1 shared owned, 2 ghost
    for (int64_t r = 0; r < nraw; ++r) {
      bool ghost = false;
      int owner = 0;
      for (int d = 0; d < 3; ++d) {
        const bool on_low = (g[d][r] == 0) && has_lower[d];
        ghost = ghost || on_low;
        owner += mult[d] * (p->rcoord[d] - (on_low ? 1 : 0));
      }
      if (ghost) {
        ghosts.push_back({(int32_t)owner, key_of(r), (int32_t)r});
        kind[r] = 2;
      } else {
        ++nowned;
        if (shared_owned(r)) {
          kind[r] = 1;
          ++nshared;
        }
      }
    }
    {
      int32_t next_shared = 0, next_private = nshared;
      for (int64_t r = 0; r < nraw; ++r) {
        if (kind[r] == 1)
          new_of_raw[r] = next_shared++;
        else if (kind[r] == 0)
          new_of_raw[r] = next_private++;
      }
    }
    p->nshared = nshared;
    std::sort(ghosts.begin(), ghosts.end(), [](const Ghost& a, const Ghost& b) {
      return a.owner != b.owner ? a.owner < b.owner : a.key < b.key;
    });
    p->nowned = nowned;
    for (size_t k = 0; k < ghosts.size(); ++k)
      new_of_raw[ghosts[k].raw] = nowned + (int32_t)k;
    p->global_key.resize((size_t)nraw);
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < nraw; ++r)
      p->global_key[new_of_raw[r]] = key_of(r);

    // receive lists: ghosts are already grouped by owner and sorted by key
    struct Lists {
      std::vector<int32_t> send, recv;
    };
    std::vector<std::pair<int, Lists>> nb; // kept sorted by rank
    auto lists_of = [&](int q) -> Lists& {
      auto it = std::lower_bound(nb.begin(), nb.end(), q,
                                 [](const std::pair<int, Lists>& a, int v) { return a.first < v; });
      if (it == nb.end() || it->first != q)
        it = nb.insert(it, {q, Lists()});
      return it->second;
    };
    for (size_t k = 0; k < ghosts.size(); ++k)
      lists_of(ghosts[k].owner).recv.push_back(nowned + (int32_t)k);

    // send lists: owned nodes on the top planes shared with each upper neighbour
    std::vector<std::pair<int64_t, int32_t>> cand[8]; // per delta mask: (global key, new index)
    for (int64_t r = 0; r < nraw; ++r) {
      const int32_t nw = new_of_raw[r];
      if (nw >= nowned)
        continue;
      int t = 0;
      for (int d = 0; d < 3; ++d)
        if (g[d][r] == top[d] && has_upper[d])
          t |= 1 << d;
      if (!t)
        continue;
      for (int m = 1; m < 8; ++m)
        if ((m & t) == m)
          cand[m].push_back({p->global_key[nw], nw});
    }
    for (int m = 1; m < 8; ++m) {
      const int delta[3] = {m & 1, (m >> 1) & 1, (m >> 2) & 1};
      bool ok = true;
      for (int d = 0; d < 3; ++d)
        ok = ok && (!delta[d] || has_upper[d]);
      if (!ok)
        continue;
      const int q = ((p->rcoord[0] + delta[0]) * Py + p->rcoord[1] + delta[1]) * Pz + p->rcoord[2]
                    + delta[2];
      std::sort(cand[m].begin(), cand[m].end());
      Lists& L = lists_of(q);
      for (auto& kv : cand[m])
        L.send.push_back(kv.second);
    }
    p->send_off.assign(1, 0);
    p->recv_off.assign(1, 0);
    for (auto& e : nb) {
      p->neigh.push_back(e.first);
      p->send_idx.insert(p->send_idx.end(), e.second.send.begin(), e.second.send.end());
      p->recv_idx.insert(p->recv_idx.end(), e.second.recv.begin(), e.second.recv.end());
      p->send_off.push_back((int64_t)p->send_idx.size());
      p->recv_off.push_back((int64_t)p->recv_idx.size());
    }

    // interface cells first
    std::vector<char> shared((size_t)nraw, 0);
    for (int64_t i = nowned; i < nraw; ++i)
      shared[i] = 1;
    for (int32_t s : p->send_idx)
      shared[s] = 1;
    std::vector<int64_t> perm;
    perm.reserve((size_t)ncl);
    std::vector<char> iface((size_t)ncl, 0);
#pragma omp parallel for schedule(static)
    for (int64_t c = 0; c < ncl; ++c) {
      const int32_t* row = raw.data() + c * Nd;
      char any = 0;
      for (int i = 0; i < Nd && !any; ++i)
        any = shared[new_of_raw[row[i]]];
      iface[c] = any;
    }
    for (int64_t c = 0; c < ncl; ++c)
      if (iface[c])
        perm.push_back(c);
    p->ninterface = (int64_t)perm.size();
    for (int64_t c = 0; c < ncl; ++c)
      if (!iface[c])
        perm.push_back(c);
    std::vector<int64_t> inv((size_t)ncl);
    for (int64_t k = 0; k < ncl; ++k)
      inv[perm[k]] = k;

    p->dofmap.resize((size_t)ncl * Nd);
    p->xdofmap.resize((size_t)ncl * 8);
    p->cell_global.resize((size_t)ncl);
    const int64_t vy = nl[1] + 1, vz = nl[2] + 1;
#pragma omp parallel for schedule(static)
    for (int64_t k = 0; k < ncl; ++k) {
      const int64_t c = perm[k];
      const int32_t* row = raw.data() + c * Nd;
      int32_t* o = p->dofmap.data() + k * Nd;
      for (int i = 0; i < Nd; ++i)
        o[i] = new_of_raw[row[i]];
      const int64_t cx = c / ((int64_t)nl[1] * nl[2]), cy = (c / nl[2]) % nl[1], cz = c % nl[2];
      for (int v = 0; v < 8; ++v) // local vertex numbering of fus_box_mesh on the local block
        p->xdofmap[k * 8 + v]
            = (int32_t)(((cx + (v & 1)) * vy + (cy + ((v >> 1) & 1))) * vz + (cz + (v >> 2)));
      p->cell_global[k] = ((cx + p->cell_lo[0]) * (int64_t)n_global[1] + (cy + p->cell_lo[1]))
                              * (int64_t)n_global[2]
                          + (cz + p->cell_lo[2]);
    }

    // exterior facets of the GLOBAL box that belong to local cells
    const int64_t nf = box_facets(nl, nullptr);
    std::vector<int32_t> f((size_t)3 * nf);
    box_facets(nl, f.data());
    static const int fdir[6] = {2, 1, 0, 0, 1, 2}, fside[6] = {0, 0, 0, 1, 1, 1};
    for (int64_t k = 0; k < nf; ++k) {
      const int lf = f[3 * k + 1], d = fdir[lf];
      if ((fside[lf] == 0 && has_lower[d]) || (fside[lf] == 1 && has_upper[d]))
        continue; // an internal face of the global box
      p->facets.push_back((int32_t)inv[f[3 * k]]);
      p->facets.push_back(lf);
      p->facets.push_back(f[3 * k + 2]);
    }
  } catch (const std::bad_alloc&) {
    delete p;
    set_error("fus_box_partition_create: out of host memory");
    return FUS_ERR_ARG;
  }
  *out = p;
  return FUS_OK;
}

int fus_box_partition_info(const fus_partition* p, int64_t sizes[9], int32_t n_local[3],
                           int32_t cell_lo[3]) {
  if (!p || !sizes)
    return FUS_ERR_ARG;
  sizes[0] = p->ncells;
  sizes[1] = p->ndofs;
  sizes[2] = p->nowned;
  sizes[3] = (int64_t)p->facets.size() / 3;
  sizes[4] = (int64_t)p->neigh.size();
  sizes[5] = (int64_t)p->send_idx.size();
  sizes[6] = (int64_t)p->recv_idx.size();
  sizes[7] = p->ninterface;
  sizes[8] = p->ndofs_global;
  for (int d = 0; d < 3; ++d) {
    if (n_local)
      n_local[d] = p->n_local[d];
    if (cell_lo)
      cell_lo[d] = p->cell_lo[d];
  }
  return FUS_OK;
}

int fus_box_partition_arrays(const fus_partition* p, int32_t* dofmap, int32_t* xdofmap,
                             int64_t* cell_global, int64_t* global_key, int32_t* facets,
                             int32_t* neigh, int64_t* send_off, int32_t* send_idx,
                             int64_t* recv_off, int32_t* recv_idx) {
  if (!p)
    return FUS_ERR_ARG;
  auto copy = [](auto* dst, const auto& v) {
    if (dst && !v.empty())
      std::memcpy(dst, v.data(), sizeof(v[0]) * v.size());
  };
  copy(dofmap, p->dofmap);
  copy(xdofmap, p->xdofmap);
  copy(cell_global, p->cell_global);
  copy(global_key, p->global_key);
  copy(facets, p->facets);
  copy(neigh, p->neigh);
  copy(send_off, p->send_off);
  copy(send_idx, p->send_idx);
  copy(recv_off, p->recv_off);
  copy(recv_idx, p->recv_idx);
  return FUS_OK;
}

int fus_box_partition_destroy(fus_partition* p) {
  delete p;
  return FUS_OK;
}

} // extern "C"
