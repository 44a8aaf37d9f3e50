// fus/model_base.hpp -- shared implementation of the three solver mirrors.
#pragma once

#include "spectral_op.hpp"

namespace fus::detail {

/// Everything {Linear,Lossy,Westervelt}Spectral3D have in common (Linear.hpp:52-347): function
/// space, device context, lumped boundary vectors, device model, init / rk4 / u_sol.  The 2-D
/// classes of cpp/fenicsx-sf-naive/common (LinearSpectral2D, ...) are the same flow on a
/// quadrilateral mesh, so they share this base: the dimension comes from the mesh.
template <typename T, int P>
class SpectralModel3D {
public:
  SpectralModel3D(int kind, basix::FiniteElement<T> element,
                  std::shared_ptr<mesh::Mesh<T>> Mesh,
                  std::shared_ptr<mesh::MeshTags<std::int32_t>> FacetTags,
                  std::shared_ptr<fem::Function<T>> speedOfSound,
                  std::shared_ptr<fem::Function<T>> density,
                  std::shared_ptr<fem::Function<T>> diffusivityOfSound,
                  std::shared_ptr<fem::Function<T>> coefficientOfNonlinearity,
                  const T& sourceFrequency, const T& sourceAmplitude, const T& sourceSpeed)
      : mesh(Mesh), ft(FacetTags) {
    static_assert(std::is_same_v<T, double>, "the B200 path is FP64 only");
    if (element.degree() != P)
      throw std::runtime_error("element degree != P");
    V = std::make_shared<fem::FunctionSpace<T>>(fem::create_functionspace(mesh, element));
    index_map = V->dofmap()->index_map;
    bs = V->dofmap()->index_map_bs();
    u_n = std::make_shared<fem::Function<T>>(V);
    v_n = std::make_shared<fem::Function<T>>(V);
    ctx = std::make_shared<SpaceContext<T>>(*V);

    // exterior facets {cell, local facet, tag}: the tag of a facet comes from FacetTags
    // (fem::compute_integration_domains, Linear.hpp:101-124); untagged exterior facets get 0 and
    // only take part in the `ds` integrals without an id.
    std::vector<std::int32_t> facets = mesh->exterior_facets();
    for (std::size_t k = 0; k < facets.size() / 3; ++k)
      facets[3 * k + 2] = 0;
    auto idx = ft->indices();
    auto val = ft->values();
    for (std::size_t i = 0; i < idx.size(); ++i)
      for (std::size_t k = 0; k < facets.size() / 3; ++k)
        if (facets[3 * k] * mesh->facets_per_cell() + facets[3 * k + 1] == idx[i]) {
          facets[3 * k + 2] = val[i];
          break;
        }

    const std::int64_t nd = index_map->size_local() + index_map->num_ghosts();
    const int tdim = mesh->topology()->dim();
    const std::int64_t nc = mesh->topology()->index_map(tdim)->size_local();
    auto x = mesh->geometry().x();
    auto xd = mesh->geometry().dofmap();
    auto dm = V->dofmap()->map();
    std::vector<T> src(nd), dsrc(nd), absb(nd), bmass(nd);
    const T* delta = diffusivityOfSound ? diffusivityOfSound->x()->array().data() : nullptr;
    const T* beta = coefficientOfNonlinearity ? coefficientOfNonlinearity->x()->array().data()
                                              : nullptr;
    auto boundary_vectors = tdim == 2 ? fus_boundary_vectors_2d : fus_boundary_vectors;
    check(boundary_vectors(kind, P, nc, nd, x.data(), xd.data_handle(), dm.data_handle(),
                               (std::int64_t)facets.size() / 3, facets.data(),
                               speedOfSound->x()->array().data(), density->x()->array().data(),
                               delta, src.data(), dsrc.data(), absb.data(), bmass.data()),
          "fus_boundary_vectors");
    check(fus_model_create(ctx->get(), kind, speedOfSound->x()->array().data(),
                           density->x()->array().data(), delta, beta, src.data(), dsrc.data(),
                           absb.data(), bmass.data(), sourceFrequency, sourceAmplitude,
                           sourceSpeed, &model),
          "fus_model_create");
  }

  ~SpectralModel3D() { fus_model_destroy(model); }
  SpectralModel3D(const SpectralModel3D&) = delete;
  SpectralModel3D& operator=(const SpectralModel3D&) = delete;

  /// Set the initial values of u and v, i.e. u_0 and v_0 (Linear.hpp:161-164)
  void init() {
    u_n->x()->set(0.0);
    v_n->x()->set(0.0);
  }

  /// Evaluate du/dt = f0(t, u, v) (Linear.hpp:171-174)
  void f0(T&, std::shared_ptr<la::Vector<T>>, std::shared_ptr<la::Vector<T>> v,
          std::shared_ptr<la::Vector<T>> result) {
    std::copy(v->array().begin(), v->array().end(), result->mutable_array().begin());
  }

  /// Evaluate dv/dt = f1(t, u, v) (Linear.hpp:181-222) on the device
  void f1(T& t, std::shared_ptr<la::Vector<T>> u, std::shared_ptr<la::Vector<T>> v,
          std::shared_ptr<la::Vector<T>> result) {
    check(fus_model_f1(model, t, u->array().data(), v->array().data(),
                       result->mutable_array().data()),
          "fus_model_f1");
  }

  /// Runge-Kutta 4th order solver (Linear.hpp:228-314): u_n, v_n go to the device, the whole
  /// time loop runs there, the final fields come back into u_n, v_n.
  void rk4(const T& startTime, const T& finalTime, const T& timeStep) {
    check(fus_model_set_state(model, u_n->x()->array().data(), v_n->x()->array().data()),
          "fus_model_set_state");
    check(fus_model_rk4(model, startTime, finalTime, timeStep, &steps_taken), "fus_model_rk4");
    check(fus_model_get_state(model, u_n->x()->mutable_array().data(),
                              v_n->x()->mutable_array().data()),
          "fus_model_get_state");
  }

  std::shared_ptr<fem::Function<T>> u_sol() const { return u_n; }
  std::shared_ptr<fem::Function<T>> v_sol() const { return v_n; } // [shim]
  std::int64_t number_of_dofs() const { return V->dofmap()->index_map->size_global(); }
  int number_of_steps() const { return steps_taken; } // [shim]

protected:
  int bs = 1, steps_taken = 0;
  std::shared_ptr<mesh::Mesh<T>> mesh;
  std::shared_ptr<mesh::MeshTags<std::int32_t>> ft;
  std::shared_ptr<const common::IndexMap> index_map;
  std::shared_ptr<fem::FunctionSpace<T>> V;
  std::shared_ptr<fem::Function<T>> u_n, v_n;
  std::shared_ptr<SpaceContext<T>> ctx;
  fus_model* model = nullptr;
};

} // namespace fus::detail
