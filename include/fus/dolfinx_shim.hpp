// dolfinx_shim.hpp -- the slice of the DOLFINx / Basix API that the hot path touches.
//
// DOLFINx, Basix and FFCx are not installed in this image, so the C++ mirror of the reference's
// operator and solver classes (fus/spectral_op.hpp, fus/Linear.hpp, ...) is written against this
// shim: same namespaces, class names and member names as DOLFINx for everything the reference's
// cpp/fenicsx-sf/common/*.hpp uses, backed by this repository's own structured-box mesh and dofmap
// generator (C ABI in fus_b200.h).  Members that exist only here are marked [shim].
#pragma once

#include "../fus_b200.h"

#include <algorithm>
#include <array>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <memory>
#include <mutex>
#include <numeric>
#include <span>
#include <stdexcept>
#include <string>
#include <thread>
#include <utility>
#include <vector>

namespace fus {
inline void check(int rc, const char* what) {
  if (rc != FUS_OK)
    throw std::runtime_error(std::string(what) + ": " + fus_last_error());
}

/// [shim] What the reference gets from MPI_COMM_WORLD.  There is no MPI in this image, so the ranks
/// of a partitioned run are the threads of ONE process, one per GPU (rank r drives device r):
/// `fus::run_ranks(size, pgrid, fn)` starts them and hands each its `Comm`.  A default-constructed
/// Comm is the single-rank world.  The few collectives the reference's drivers use on the host
/// (barrier, min/max/sum reductions of scalars: BM7-SC1/main.cpp:77-78) are provided; the ghost
/// exchanges of the time loop never come here -- they run between the GPUs (fus_halo_*).
class Comm {
public:
  struct PeerInfo { // what a rank publishes for its neighbours (halo set-up)
    void* mailbox = nullptr;
    int device = 0;
    std::vector<std::int64_t> layout, soff, roff;
    std::vector<int> neigh;
  };
  struct World {
    explicit World(int n, std::array<int, 3> pg)
        : size(n), pgrid(pg), bar(n), scratch(n, 0.0), peers(n), nccl_id(128, 0) {}
    int size;
    std::array<int, 3> pgrid;
    std::barrier<> bar;
    std::vector<double> scratch;
    std::vector<PeerInfo> peers;
    std::vector<unsigned char> nccl_id;
  };
  Comm() : _world(std::make_shared<World>(1, std::array<int, 3>{1, 1, 1})), _rank(0) {}
  Comm(std::shared_ptr<World> w, int rank) : _world(std::move(w)), _rank(rank) {}
  int rank() const { return _rank; }
  int size() const { return _world->size; }
  const std::array<int, 3>& pgrid() const { return _world->pgrid; }
  World& world() const { return *_world; }
  void barrier() const { _world->bar.arrive_and_wait(); }
  /// MPI_Allreduce of one scalar: op = 0 min, 1 max, 2 sum
  double allreduce(double v, int op) const {
    _world->scratch[_rank] = v;
    barrier();
    double r = _world->scratch[0];
    for (int q = 1; q < size(); ++q)
      r = op == 0 ? std::min(r, _world->scratch[q])
                  : (op == 1 ? std::max(r, _world->scratch[q]) : r + _world->scratch[q]);
    barrier();
    return r;
  }

private:
  std::shared_ptr<World> _world;
  int _rank;
};

/// [shim] mpirun: fn(comm) on `size` threads; the first exception of any rank is rethrown.
inline void run_ranks(int size, std::array<int, 3> pgrid, const std::function<void(Comm&)>& fn) {
  if (pgrid[0] * pgrid[1] * pgrid[2] != size)
    throw std::runtime_error("run_ranks: the process grid does not match the number of ranks");
  auto world = std::make_shared<Comm::World>(size, pgrid);
  std::vector<std::thread> th;
  std::vector<std::string> err(size);
  for (int r = 0; r < size; ++r)
    th.emplace_back([&, r] {
      Comm comm(world, r);
      try {
        fn(comm);
      } catch (const std::exception& e) {
        err[r] = e.what();
        std::fprintf(stderr, "rank %d: %s\n", r, e.what());
        std::fflush(stderr);
        std::_Exit(1); // a rank that died would leave the others waiting in a barrier
      }
    });
  for (auto& t : th)
    t.join();
}

/// Minimal row-major 2-D view with the mdspan members the reference uses
/// (`extent`, `operator()`, `size`, `data_handle`).
template <typename T>
class View2D {
public:
  View2D() = default;
  View2D(T* p, std::size_t n0, std::size_t n1) : _p(p), _n0(n0), _n1(n1) {}
  std::size_t extent(int i) const { return i == 0 ? _n0 : _n1; }
  std::size_t size() const { return _n0 * _n1; }
  T& operator()(std::size_t i, std::size_t j) const { return _p[i * _n1 + j]; }
  T* data_handle() const { return _p; }

private:
  T* _p = nullptr;
  std::size_t _n0 = 0, _n1 = 0;
};
} // namespace fus

namespace basix {
namespace cell {
enum class type { interval, quadrilateral, hexahedron };
}
namespace element {
enum class family { P };
enum class lagrange_variant { gll_warped };
enum class dpc_variant { unset };
} // namespace element

/// basix::FiniteElement<T>: only the degree is consumed by the hot path.
template <typename T>
class FiniteElement {
public:
  explicit FiniteElement(int degree) : _degree(degree) {}
  int degree() const { return _degree; }

private:
  int _degree;
};

template <typename T>
FiniteElement<T> create_element(element::family, cell::type, int degree, element::lagrange_variant,
                                element::dpc_variant, bool) {
  return FiniteElement<T>(degree);
}
} // namespace basix

namespace dolfinx {

namespace common {
/// common::IndexMap: owned entries first, then ghosts.
class IndexMap {
public:
  IndexMap(std::int64_t size_local, std::int64_t num_ghosts, std::int64_t size_global)
      : _local(size_local), _ghosts(num_ghosts), _global(size_global) {}
  std::int32_t size_local() const { return (std::int32_t)_local; }
  std::int32_t num_ghosts() const { return (std::int32_t)_ghosts; }
  std::int64_t size_global() const { return _global; }
  /// global index of local entries (owned and ghosts); identity on a single rank
  void local_to_global(std::span<const std::int32_t> local, std::span<std::int64_t> global) const {
    for (std::size_t i = 0; i < local.size(); ++i)
      global[i] = _l2g.empty() ? (std::int64_t)local[i] : _l2g[local[i]];
  }
  /// [shim] set by the partitioned function space: global node id of every local dof
  void set_global_indices(std::vector<std::int64_t> l2g) { _l2g = std::move(l2g); }
  /// [shim] la::Vector::scatter_fwd / scatter_rev of host vectors go through the device context of
  /// the function space (registered by fus::detail::SpaceContext); forward = owner -> ghost
  using Scatter = std::function<void(double* x, bool forward)>;
  void set_scatter(Scatter s) const { _scatter = std::move(s); }
  const Scatter& scatter() const { return _scatter; }

private:
  std::int64_t _local, _ghosts, _global;
  std::vector<std::int64_t> _l2g;
  mutable Scatter _scatter;
};
} // namespace common

namespace la {
/// la::Vector<T>: host array of owned + ghost entries (single rank: no ghosts, scatters are no-ops).
template <typename T, typename Alloc = std::allocator<T>>
class Vector {
public:
  Vector(std::shared_ptr<const common::IndexMap> map, int bs)
      : _map(std::move(map)), _bs(bs), _x((_map->size_local() + _map->num_ghosts()) * bs, T(0)) {}
  std::span<const T> array() const { return std::span<const T>(_x.data(), _x.size()); }
  std::span<T> mutable_array() { return std::span<T>(_x.data(), _x.size()); }
  std::shared_ptr<const common::IndexMap> index_map() const { return _map; }
  int bs() const { return _bs; }
  void set(T v) { std::fill(_x.begin(), _x.end(), v); }
  /// owner -> ghost (Linear.hpp:196,199); a no-op without ghosts
  void scatter_fwd() { scatter(true); }
  /// ghost -> owner with std::plus (Linear.hpp:134,206); any other operation is not provided
  template <typename Op>
  void scatter_rev(Op) {
    scatter(false);
  }

private:
  void scatter(bool forward) {
    if (_map->num_ghosts() == 0)
      return;
    if (!_map->scatter())
      throw std::runtime_error("la::Vector [shim]: no device context registered for this index map");
    if constexpr (std::is_same_v<T, double>) {
      _map->scatter()(_x.data(), forward);
    } else {
      std::vector<double> w(_x.begin(), _x.end());
      _map->scatter()(w.data(), forward);
      std::transform(w.begin(), w.end(), _x.begin(), [](double v) { return (T)v; });
    }
  }
  std::shared_ptr<const common::IndexMap> _map;
  int _bs;
  std::vector<T, Alloc> _x;
};
} // namespace la

namespace mesh {
enum class CellType { hexahedron, quadrilateral };
enum class GhostMode { none };

class Topology {
public:
  Topology(std::int64_t ncells, int dim = 3)
      : _cells(std::make_shared<common::IndexMap>(ncells, 0, ncells)), _dim(dim) {}
  int dim() const { return _dim; }
  std::shared_ptr<const common::IndexMap> index_map(int) const { return _cells; }
  void create_connectivity(int, int) const {}

private:
  std::shared_ptr<const common::IndexMap> _cells;
  int _dim;
};

template <typename T>
class Geometry {
public:
  Geometry(std::vector<T> x, std::vector<std::int32_t> dofmap, int gdim = 3)
      : _x(std::move(x)), _dofmap(std::move(dofmap)), _gdim(gdim) {}
  int dim() const { return _gdim; }
  /// padded to 3 coordinates per vertex whatever the geometric dimension, as in DOLFINx
  std::span<const T> x() const { return std::span<const T>(_x.data(), _x.size()); }
  fus::View2D<const std::int32_t> dofmap() const {
    const std::size_t nv = _gdim == 3 ? 8 : 4;
    return fus::View2D<const std::int32_t>(_dofmap.data(), _dofmap.size() / nv, nv);
  }

private:
  std::vector<T> _x;
  std::vector<std::int32_t> _dofmap;
  int _gdim;
};

template <typename T>
class Mesh {
public:
  Mesh(std::array<int, 3> n, Geometry<T> g, std::vector<std::int32_t> exterior_facets,
       int tdim = 3)
      : _n(n), _geometry(std::move(g)),
        _topology(std::make_shared<Topology>((std::int64_t)n[0] * n[1] * n[2], tdim)),
        _ext(std::move(exterior_facets)) {}
  const Geometry<T>& geometry() const { return _geometry; }
  std::shared_ptr<Topology> topology() const { return _topology; }
  std::shared_ptr<Topology> topology_mutable() const { return _topology; }
  /// [shim] cells per direction of the structured box (third entry 1 for a rectangle)
  const std::array<int, 3>& box_cells() const { return _n; }
  /// [shim] facets per cell: the stride of the facet "index" cell * facets_per_cell + local facet
  int facets_per_cell() const { return _topology->dim() == 3 ? 6 : 4; }
  /// [shim] exterior facets as {cell, local facet, box tag} triplets
  const std::vector<std::int32_t>& exterior_facets() const { return _ext; }
  /// [shim] the partition this mesh is the local block of (single rank: the whole box)
  struct Partition {
    fus::Comm comm;
    std::array<int, 3> n_global{0, 0, 0};
    std::vector<std::int64_t> cell_global; // global cell index of every local cell
    std::int64_t ninterface_cells = 0;
  };
  void set_partition(Partition p) { _part = std::make_shared<Partition>(std::move(p)); }
  const Partition* partition() const { return _part.get(); }
  const fus::Comm& comm() const {
    static const fus::Comm self;
    return _part ? _part->comm : self;
  }
  /// [shim] cells of the GLOBAL box along x and the global x index of a local cell
  int cells_x_global() const { return _part ? _part->n_global[0] : _n[0]; }
  int cell_x_global(std::int64_t c) const {
    if (!_part)
      return (int)(c / ((std::int64_t)_n[1] * _n[2]));
    return (int)(_part->cell_global[c] / ((std::int64_t)_part->n_global[1] * _part->n_global[2]));
  }

private:
  std::shared_ptr<Partition> _part;
  std::array<int, 3> _n;
  Geometry<T> _geometry;
  std::shared_ptr<Topology> _topology;
  std::vector<std::int32_t> _ext;
};

/// mesh::create_box (experiments/measure_fraction_of_peak_performance/main.cpp:61-65)
template <typename T>
Mesh<T> create_box(std::array<std::array<T, 3>, 2> p, std::array<std::size_t, 3> n, CellType) {
  static_assert(std::is_floating_point_v<T>, "real scalar types only");
  const int nn[3] = {(int)n[0], (int)n[1], (int)n[2]};
  const std::int64_t nv = (std::int64_t)(nn[0] + 1) * (nn[1] + 1) * (nn[2] + 1);
  const std::int64_t nc = (std::int64_t)nn[0] * nn[1] * nn[2];
  std::vector<double> xdbl(3 * nv);
  std::vector<std::int32_t> xd(8 * nc);
  const double lo[3] = {(double)p[0][0], (double)p[0][1], (double)p[0][2]};
  const double hi[3] = {(double)p[1][0], (double)p[1][1], (double)p[1][2]};
  fus::check(fus_box_mesh(nn, lo, hi, xdbl.data(), xd.data()), "fus_box_mesh");
  std::vector<T> x(xdbl.begin(), xdbl.end()); // the mesh stores coordinates in its own scalar type
  const std::int64_t nf = fus_box_facets(nn, nullptr);
  std::vector<std::int32_t> f(3 * nf);
  fus_box_facets(nn, f.data());
  return Mesh<T>({nn[0], nn[1], nn[2]}, Geometry<T>(std::move(x), std::move(xd)), std::move(f));
}

/// mesh::create_box(MPI_COMM_WORLD, ...) (BM7-SC1/main.cpp:58-59, GhostMode::none): this rank's
/// block of the box, partitioned over comm.pgrid() by fus_box_partition_* -- cells that touch a dof
/// shared with a neighbour first, vertex coordinates from the global formula (bitwise those of the
/// unpartitioned box).
template <typename T>
Mesh<T> create_box(const fus::Comm& comm, std::array<std::array<T, 3>, 2> p,
                   std::array<std::size_t, 3> n, CellType ct) {
  if (comm.size() == 1)
    return create_box<T>(p, n, ct);
  const int ng[3] = {(int)n[0], (int)n[1], (int)n[2]};
  const int pg[3] = {comm.pgrid()[0], comm.pgrid()[1], comm.pgrid()[2]};
  fus_partition* part = nullptr;
  fus::check(fus_box_partition_create(1, ng, pg, comm.rank(), 1, &part), "fus_box_partition_create");
  std::int64_t sizes[9];
  std::int32_t nl[3], lo_c[3];
  fus_box_partition_info(part, sizes, nl, lo_c);
  const std::int64_t nc = sizes[0], nf = sizes[3];
  std::vector<std::int32_t> xd(8 * nc), f(3 * nf);
  typename Mesh<T>::Partition P;
  P.comm = comm;
  P.n_global = {ng[0], ng[1], ng[2]};
  P.cell_global.resize(nc);
  P.ninterface_cells = sizes[7];
  fus_box_partition_arrays(part, nullptr, xd.data(), P.cell_global.data(), nullptr, f.data(), nullptr,
                           nullptr, nullptr, nullptr, nullptr);
  fus_box_partition_destroy(part);
  const std::int64_t nv[3] = {nl[0] + 1, nl[1] + 1, nl[2] + 1};
  std::vector<T> x(3 * nv[0] * nv[1] * nv[2]);
  for (std::int64_t i = 0; i < nv[0]; ++i)
    for (std::int64_t j = 0; j < nv[1]; ++j)
      for (std::int64_t k = 0; k < nv[2]; ++k) {
        const std::int64_t v = (i * nv[1] + j) * nv[2] + k;
        const std::int64_t g[3] = {lo_c[0] + i, lo_c[1] + j, lo_c[2] + k};
        for (int d = 0; d < 3; ++d)
          x[3 * v + d] = (T)((double)p[0][d]
                             + ((double)p[1][d] - (double)p[0][d]) * (double)g[d] / (double)ng[d]);
      }
  Mesh<T> m({nl[0], nl[1], nl[2]}, Geometry<T>(std::move(x), std::move(xd)), std::move(f));
  m.set_partition(std::move(P));
  return m;
}

/// mesh::create_rectangle (quadrilateral cells) for the 2-D operators of cpp/fenicsx-sf-naive
template <typename T>
Mesh<T> create_rectangle(std::array<std::array<T, 2>, 2> p, std::array<std::size_t, 2> n, CellType) {
  static_assert(std::is_floating_point_v<T>, "real scalar types only");
  const int nn[2] = {(int)n[0], (int)n[1]};
  const std::int64_t nv = (std::int64_t)(nn[0] + 1) * (nn[1] + 1), nc = (std::int64_t)nn[0] * nn[1];
  std::vector<double> xdbl(3 * nv);
  std::vector<std::int32_t> xd(4 * nc);
  const double lo[2] = {(double)p[0][0], (double)p[0][1]}, hi[2] = {(double)p[1][0], (double)p[1][1]};
  fus::check(fus_rect_mesh(nn, lo, hi, xdbl.data(), xd.data()), "fus_rect_mesh");
  std::vector<T> x(xdbl.begin(), xdbl.end());
  const std::int64_t nf = fus_rect_facets(nn, nullptr);
  std::vector<std::int32_t> f(3 * nf);
  fus_rect_facets(nn, f.data());
  return Mesh<T>({nn[0], nn[1], 1}, Geometry<T>(std::move(x), std::move(xd), 2), std::move(f), 2);
}

/// mesh::MeshTags<int32_t> over exterior facets.  [shim] a facet "index" is
/// cell * facets_per_cell + local facet.
template <typename V>
class MeshTags {
public:
  MeshTags(std::vector<std::int32_t> indices, std::vector<V> values)
      : _indices(std::move(indices)), _values(std::move(values)) {}
  std::span<const std::int32_t> indices() const { return _indices; }
  std::span<const V> values() const { return _values; }
  std::vector<std::int32_t> find(V value) const {
    std::vector<std::int32_t> out;
    for (std::size_t i = 0; i < _values.size(); ++i)
      if (_values[i] == value)
        out.push_back(_indices[i]);
    return out;
  }

private:
  std::vector<std::int32_t> _indices;
  std::vector<V> _values;
};

/// mesh::h (BM7-SC1/main.cpp:72-73): cell diameter = largest distance between two vertices.
template <typename T>
std::vector<T> h(const Mesh<T>& mesh, std::span<const int> entities, int dim) {
  if (dim != mesh.topology()->dim())
    throw std::runtime_error("mesh::h [shim]: cells only");
  auto x = mesh.geometry().x();
  auto xd = mesh.geometry().dofmap();
  const int nv = (int)xd.extent(1);
  std::vector<T> out;
  out.reserve(entities.size());
  for (int c : entities) {
    T d2 = 0;
    for (int a = 0; a < nv; ++a)
      for (int b = a + 1; b < nv; ++b) {
        T s = 0;
        for (int r = 0; r < 3; ++r) {
          const T e = x[3 * xd(c, a) + r] - x[3 * xd(c, b) + r];
          s += e * e;
        }
        d2 = std::max(d2, s);
      }
    out.push_back(std::sqrt(d2));
  }
  return out;
}

/// [shim] cell tags of the box: `nlayers` slabs of equal thickness along x, tagged 1..nlayers
/// (stands in for the cell MeshTags the reference drivers read next to their meshes,
/// BM7-SC1/main.cpp:60-63).  A tag "index" is the local cell index, as in DOLFINx.
template <typename T>
MeshTags<std::int32_t> box_cell_layers(const Mesh<T>& mesh, int nlayers) {
  const auto& n = mesh.box_cells();
  const std::int64_t nc = (std::int64_t)n[0] * n[1] * n[2];
  std::vector<std::int32_t> idx, val;
  for (std::int64_t c = 0; c < nc; ++c) { // local cell order (interface cells first when partitioned)
    idx.push_back((std::int32_t)c);
    val.push_back(1 + std::min(nlayers - 1, mesh.cell_x_global(c) * nlayers / mesh.cells_x_global()));
  }
  return MeshTags<std::int32_t>(std::move(idx), std::move(val));
}

/// [shim] facet tags of the box: 1 on x = lo, 2 on x = hi
template <typename T>
MeshTags<std::int32_t> box_facet_tags(const Mesh<T>& mesh) {
  std::vector<std::int32_t> idx, val;
  const auto& f = mesh.exterior_facets();
  for (std::size_t k = 0; k < f.size() / 3; ++k)
    if (f[3 * k + 2] != 0) {
      idx.push_back(f[3 * k] * mesh.facets_per_cell() + f[3 * k + 1]);
      val.push_back(f[3 * k + 2]);
    }
  return MeshTags<std::int32_t>(std::move(idx), std::move(val));
}
} // namespace mesh

namespace fem {
class DofMap {
public:
  DofMap(std::vector<std::int32_t> tensor_map, int nd, std::shared_ptr<const common::IndexMap> im)
      : index_map(std::move(im)), _map(std::move(tensor_map)), _nd(nd) {}
  /// [shim] the map is already in tensor-product order (reorder_dofmap is the identity)
  fus::View2D<const std::int32_t> map() const {
    return fus::View2D<const std::int32_t>(_map.data(), _map.size() / _nd, _nd);
  }
  int index_map_bs() const { return 1; }
  std::shared_ptr<const common::IndexMap> index_map;
  /// [shim] the neighbour lists behind la::Vector::scatter_fwd / scatter_rev on this rank (what
  /// DOLFINx keeps in the IndexMap's scatterer): the arguments of fus_halo_setup
  struct Halo {
    std::vector<int> neigh;
    std::vector<std::int64_t> send_off, recv_off;
    std::vector<std::int32_t> send_idx, recv_idx;
    std::int64_t ninterface_cells = 0;
  };
  std::shared_ptr<const Halo> halo;

private:
  std::vector<std::int32_t> _map;
  int _nd;
};

template <typename T>
class FunctionSpace {
public:
  FunctionSpace(std::shared_ptr<mesh::Mesh<T>> mesh, int degree, std::shared_ptr<DofMap> dm)
      : _mesh(std::move(mesh)), _degree(degree), _dofmap(std::move(dm)) {}
  std::shared_ptr<const mesh::Mesh<T>> mesh() const { return _mesh; }
  std::shared_ptr<const DofMap> dofmap() const { return _dofmap; }
  int degree() const { return _degree; }

private:
  std::shared_ptr<mesh::Mesh<T>> _mesh;
  int _degree;
  std::shared_ptr<DofMap> _dofmap;
};

/// fem::create_functionspace(mesh, element) (Linear.hpp:82-83).  degree 0 gives the DG0 space
/// of per-cell coefficients.
template <typename T>
FunctionSpace<T> create_functionspace(std::shared_ptr<mesh::Mesh<T>> mesh,
                                      const basix::FiniteElement<T>& e) {
  const auto& n = mesh->box_cells();
  const int nn[3] = {n[0], n[1], n[2]};
  const std::int64_t nc = (std::int64_t)n[0] * n[1] * n[2];
  const int P = e.degree();
  if (P == 0) {
    std::vector<std::int32_t> dm(nc);
    std::iota(dm.begin(), dm.end(), 0);
    auto im = std::make_shared<common::IndexMap>(nc, 0, nc);
    return FunctionSpace<T>(mesh, 0, std::make_shared<DofMap>(std::move(dm), 1, im));
  }
  if (mesh->topology()->dim() == 2) {
    const int Nd2 = (P + 1) * (P + 1);
    std::vector<std::int32_t> dm2((std::size_t)nc * Nd2);
    fus::check(fus_rect_dofmap(P, nn, dm2.data()), "fus_rect_dofmap");
    const std::int64_t nd2 = fus_rect_num_dofs(P, nn);
    auto im2 = std::make_shared<common::IndexMap>(nd2, 0, nd2);
    return FunctionSpace<T>(mesh, P, std::make_shared<DofMap>(std::move(dm2), Nd2, im2));
  }
  const int Nd = (P + 1) * (P + 1) * (P + 1);
  if (const auto* part = mesh->partition()) { // this rank's share of the global space
    const int ng[3] = {part->n_global[0], part->n_global[1], part->n_global[2]};
    const int pg[3] = {part->comm.pgrid()[0], part->comm.pgrid()[1], part->comm.pgrid()[2]};
    fus_partition* bp = nullptr;
    fus::check(fus_box_partition_create(P, ng, pg, part->comm.rank(), 1, &bp),
               "fus_box_partition_create");
    std::int64_t sizes[9];
    fus_box_partition_info(bp, sizes, nullptr, nullptr);
    auto H = std::make_shared<DofMap::Halo>();
    std::vector<std::int32_t> dmp((std::size_t)sizes[0] * Nd), nb(sizes[4]);
    std::vector<std::int64_t> l2g(sizes[1]);
    H->send_off.resize(sizes[4] + 1), H->recv_off.resize(sizes[4] + 1);
    H->send_idx.resize(sizes[5]), H->recv_idx.resize(sizes[6]);
    H->ninterface_cells = sizes[7];
    fus_box_partition_arrays(bp, dmp.data(), nullptr, nullptr, l2g.data(), nullptr, nb.data(),
                             H->send_off.data(), H->send_idx.data(), H->recv_off.data(),
                             H->recv_idx.data());
    fus_box_partition_destroy(bp);
    H->neigh.assign(nb.begin(), nb.end());
    auto imp = std::make_shared<common::IndexMap>(sizes[2], sizes[1] - sizes[2], sizes[8]);
    imp->set_global_indices(std::move(l2g));
    auto dmap = std::make_shared<DofMap>(std::move(dmp), Nd, imp);
    dmap->halo = H;
    return FunctionSpace<T>(mesh, P, dmap);
  }
  std::vector<std::int32_t> dm((std::size_t)nc * Nd);
  fus::check(fus_box_dofmap(P, nn, 1, dm.data()), "fus_box_dofmap");
  const std::int64_t nd = fus_box_num_dofs(P, nn);
  auto im = std::make_shared<common::IndexMap>(nd, 0, nd);
  return FunctionSpace<T>(mesh, P, std::make_shared<DofMap>(std::move(dm), Nd, im));
}

template <typename T>
class Function {
public:
  explicit Function(std::shared_ptr<const FunctionSpace<T>> V)
      : _V(std::move(V)),
        _x(std::make_shared<la::Vector<T>>(_V->dofmap()->index_map, _V->dofmap()->index_map_bs())) {}
  std::shared_ptr<la::Vector<T>> x() { return _x; }
  std::shared_ptr<const la::Vector<T>> x() const { return _x; }
  std::shared_ptr<const FunctionSpace<T>> function_space() const { return _V; }

private:
  std::shared_ptr<const FunctionSpace<T>> _V;
  std::shared_ptr<la::Vector<T>> _x;
};
} // namespace fem

} // namespace dolfinx
