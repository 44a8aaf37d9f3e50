// fus/spectral_op.hpp -- drop-in for cpp/fenicsx-sf/common/spectral_op.hpp of the reference.
//
// Same class names, template parameters and call signatures:
//   MassSpectral3D<T,P>(std::shared_ptr<fem::FunctionSpace<T>>& V)            (spectral_op.hpp:29-32)
//   StiffnessSpectral3D<T,P>(std::shared_ptr<fem::FunctionSpace<T>>& V)       (spectral_op.hpp:132-135)
//   void operator()(const la::Vector<T,Alloc>& x, std::span<T> coeffs, la::Vector<T,Alloc>& y)
//                                                                             (:69-70, :173-174)
// with the same semantics: y += A(coeffs) x over the local cells, no communication, the caller
// zero-fills y and refreshes the ghosts of x.  The arithmetic runs in the CUDA library through the
// C ABI (fus_b200.h) in FP64.  T = float (the reference's test_operators3d and its float timing
// runs instantiate the operators with it) is accepted by the two operator classes: vectors and
// geometry are widened on the way in and the result is rounded once on the way out, so float
// drivers get at least the accuracy of the reference's float arithmetic, not its speed (there are
// no FP32 kernels yet).  Construction uploads the cell data once (the reference
// precomputes G / detJ in its constructor too); each call moves x, coeffs and y across PCIe --
// the solver classes (fus/Linear.hpp, ...) keep everything resident instead.
#pragma once

#include "dolfinx_shim.hpp"

namespace fus::detail {
/// One device context per function space, shared by the operators built on it.
template <typename T>
class SpaceContext {
public:
  explicit SpaceContext(const dolfinx::fem::FunctionSpace<T>& V, int device = 0) {
    static_assert(std::is_floating_point_v<T>, "real scalar types only");
    auto mesh = V.mesh();
    auto dm = V.dofmap()->map();
    auto im = V.dofmap()->index_map;
    auto xd = mesh->geometry().dofmap();
    auto xT = mesh->geometry().x();
    const std::vector<double> x(xT.begin(), xT.end()); // the device path is FP64
    auto create = mesh->topology()->dim() == 2 ? fus_ctx_create_from_mesh_2d
                                               : fus_ctx_create_from_mesh;
    check(create(V.degree(), (std::int64_t)dm.extent(0), im->size_local() + im->num_ghosts(),
                 im->size_local(), dm.data_handle(), (std::int64_t)x.size() / 3, x.data(),
                 xd.data_handle(), device, &_ctx),
          "fus_ctx_create_from_mesh");
  }
  ~SpaceContext() { fus_ctx_destroy(_ctx); }
  SpaceContext(const SpaceContext&) = delete;
  SpaceContext& operator=(const SpaceContext&) = delete;
  fus_ctx* get() const { return _ctx; }

private:
  fus_ctx* _ctx = nullptr;
};

/// y += A(coeffs) x through a `*_apply_host` entry point; scalar types other than double are
/// widened first and the accumulated result is rounded back once.
template <typename T>
void apply_host(int (*entry)(fus_ctx*, const double*, const double*, double*), const char* what,
                fus_ctx* ctx, std::span<const T> x, std::span<T> coeffs, std::span<T> y) {
  if constexpr (std::is_same_v<T, double>) {
    check(entry(ctx, x.data(), coeffs.data(), y.data()), what);
  } else {
    const std::vector<double> xd(x.begin(), x.end()), cd(coeffs.begin(), coeffs.end());
    std::vector<double> yd(y.begin(), y.end());
    check(entry(ctx, xd.data(), cd.data(), yd.data()), what);
    std::transform(yd.begin(), yd.end(), y.begin(), [](double v) { return (T)v; });
  }
}
} // namespace fus::detail

using namespace dolfinx;

/// 3D Spectral Mass operator (spectral_op.hpp:28-107)
template <typename T, int P>
class MassSpectral3D {
public:
  MassSpectral3D(std::shared_ptr<fem::FunctionSpace<T>>& V)
      : _ctx(std::make_shared<fus::detail::SpaceContext<T>>(*V)) {
    static_assert(P >= 1 && P <= 7, "supported degrees: 1..7");
    if (V->degree() != P)
      throw std::runtime_error("MassSpectral3D: function space degree != P");
  }

  /// Operator y += M x
  template <typename Alloc>
  void operator()(const la::Vector<T, Alloc>& x, std::span<T> coeffs, la::Vector<T, Alloc>& y) {
    fus::detail::apply_host<T>(fus_mass_apply_host, "fus_mass_apply_host", _ctx->get(), x.array(),
                               coeffs, y.mutable_array());
  }

private:
  std::shared_ptr<fus::detail::SpaceContext<T>> _ctx;
};

/// 3D Spectral Stiffness operator (spectral_op.hpp:132-284)
template <typename T, int P>
class StiffnessSpectral3D {
public:
  StiffnessSpectral3D(std::shared_ptr<fem::FunctionSpace<T>>& V)
      : _ctx(std::make_shared<fus::detail::SpaceContext<T>>(*V)) {
    static_assert(P >= 1 && P <= 7, "supported degrees: 1..7");
    if (V->degree() != P)
      throw std::runtime_error("StiffnessSpectral3D: function space degree != P");
  }

  /// Operator y += K x
  template <typename Alloc>
  void operator()(const la::Vector<T, Alloc>& x, std::span<T> coeffs, la::Vector<T, Alloc>& y) {
    fus::detail::apply_host<T>(fus_stiffness_apply_host, "fus_stiffness_apply_host", _ctx->get(),
                               x.array(), coeffs, y.mutable_array());
  }

private:
  std::shared_ptr<fus::detail::SpaceContext<T>> _ctx;
};

/// 2D Spectral Mass operator (cpp/fenicsx-sf-naive/common/spectral_op.hpp:28-107): the same
/// device path on a quadrilateral function space
template <typename T, int P>
class MassSpectral2D : public MassSpectral3D<T, P> {
public:
  MassSpectral2D(std::shared_ptr<fem::FunctionSpace<T>>& V) : MassSpectral3D<T, P>(V) {
    if (V->mesh()->topology()->dim() != 2)
      throw std::runtime_error("MassSpectral2D: quadrilateral mesh expected");
  }
};

/// 2D Spectral Stiffness operator (cpp/fenicsx-sf-naive/common/spectral_op.hpp:226-359)
template <typename T, int P>
class StiffnessSpectral2D : public StiffnessSpectral3D<T, P> {
public:
  StiffnessSpectral2D(std::shared_ptr<fem::FunctionSpace<T>>& V) : StiffnessSpectral3D<T, P>(V) {
    if (V->mesh()->topology()->dim() != 2)
      throw std::runtime_error("StiffnessSpectral2D: quadrilateral mesh expected");
  }
};
