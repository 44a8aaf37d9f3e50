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
#include <cmath>
#include <cstdint>
#include <functional>
#include <memory>
#include <numeric>
#include <span>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace fus {
inline void check(int rc, const char* what) {
  if (rc != FUS_OK)
    throw std::runtime_error(std::string(what) + ": " + fus_last_error());
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

private:
  std::int64_t _local, _ghosts, _global;
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
  void scatter_fwd() {}
  template <typename Op>
  void scatter_rev(Op) {}

private:
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

private:
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
  std::vector<std::int32_t> idx, val;
  std::int32_t c = 0;
  for (int i = 0; i < n[0]; ++i)
    for (int j = 0; j < n[1]; ++j)
      for (int k = 0; k < n[2]; ++k, ++c) {
        idx.push_back(c);
        val.push_back(1 + std::min(nlayers - 1, i * nlayers / n[0]));
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
