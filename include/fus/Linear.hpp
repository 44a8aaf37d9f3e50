// fus/Linear.hpp -- drop-in for cpp/fenicsx-sf/common/Linear.hpp:52-347 of the reference.
#pragma once
#include "model_base.hpp"

/// Solver for the 3D second order linear wave equation (GLL lattice + GLL quadrature, diagonal
/// mass matrix).  Same constructor and methods as the reference class.
template <typename T, int P>
class LinearSpectral3D : public fus::detail::SpectralModel3D<T, P> {
public:
  LinearSpectral3D(basix::FiniteElement<T> element, std::shared_ptr<mesh::Mesh<T>> Mesh,
                   std::shared_ptr<mesh::MeshTags<std::int32_t>> FacetTags,
                   std::shared_ptr<fem::Function<T>> speedOfSound,
                   std::shared_ptr<fem::Function<T>> density, const T& sourceFrequency,
                   const T& sourceAmplitude, const T& sourceSpeed)
      : fus::detail::SpectralModel3D<T, P>(FUS_LINEAR, element, Mesh, FacetTags, speedOfSound,
                                           density, nullptr, nullptr, sourceFrequency,
                                           sourceAmplitude, sourceSpeed) {}
};

/// LinearSpectral2D<T,P> (cpp/fenicsx-sf-naive/common/Linear.hpp:52-350): quadrilateral mesh, same flow
template <typename T, int P>
class LinearSpectral2D : public fus::detail::SpectralModel3D<T, P> {
public:
  LinearSpectral2D(basix::FiniteElement<T> element, std::shared_ptr<mesh::Mesh<T>> Mesh,
                  std::shared_ptr<mesh::MeshTags<std::int32_t>> FacetTags,
                  std::shared_ptr<fem::Function<T>> speedOfSound,
                  std::shared_ptr<fem::Function<T>> density,
                  const T& sourceFrequency, const T& sourceAmplitude, const T& sourceSpeed)
      : fus::detail::SpectralModel3D<T, P>(FUS_LINEAR, element, Mesh, FacetTags, speedOfSound, density,
                                           nullptr, nullptr, sourceFrequency, sourceAmplitude,
                                           sourceSpeed) {
    if (Mesh->topology()->dim() != 2)
      throw std::runtime_error("LinearSpectral2D: quadrilateral mesh expected");
  }
};
