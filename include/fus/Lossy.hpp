// fus/Lossy.hpp -- drop-in for cpp/fenicsx-sf/common/Lossy.hpp:54-380 of the reference.
#pragma once
#include "model_base.hpp"

/// Solver for the 3D second order linear wave equation with attenuation.
template <typename T, int P>
class LossySpectral3D : public fus::detail::SpectralModel3D<T, P> {
public:
  LossySpectral3D(basix::FiniteElement<T> element, std::shared_ptr<mesh::Mesh<T>> Mesh,
                  std::shared_ptr<mesh::MeshTags<std::int32_t>> FacetTags,
                  std::shared_ptr<fem::Function<T>> speedOfSound,
                  std::shared_ptr<fem::Function<T>> density,
                  std::shared_ptr<fem::Function<T>> diffusivityOfSound, const T& sourceFrequency,
                  const T& sourceAmplitude, const T& sourceSpeed)
      : fus::detail::SpectralModel3D<T, P>(FUS_LOSSY, element, Mesh, FacetTags, speedOfSound,
                                           density, diffusivityOfSound, nullptr, sourceFrequency,
                                           sourceAmplitude, sourceSpeed) {}
};

/// LossySpectral2D<T,P> (cpp/fenicsx-sf-naive/common/Lossy.hpp): quadrilateral mesh, same flow
template <typename T, int P>
class LossySpectral2D : public fus::detail::SpectralModel3D<T, P> {
public:
  LossySpectral2D(basix::FiniteElement<T> element, std::shared_ptr<mesh::Mesh<T>> Mesh,
                  std::shared_ptr<mesh::MeshTags<std::int32_t>> FacetTags,
                  std::shared_ptr<fem::Function<T>> speedOfSound,
                  std::shared_ptr<fem::Function<T>> density,
                  std::shared_ptr<fem::Function<T>> diffusivityOfSound,
                  const T& sourceFrequency, const T& sourceAmplitude, const T& sourceSpeed)
      : fus::detail::SpectralModel3D<T, P>(FUS_LOSSY, element, Mesh, FacetTags, speedOfSound, density,
                                           diffusivityOfSound, nullptr, sourceFrequency, sourceAmplitude,
                                           sourceSpeed) {
    if (Mesh->topology()->dim() != 2)
      throw std::runtime_error("LossySpectral2D: quadrilateral mesh expected");
  }
};

#ifndef FUS_HAVE_COMPUTE_DIFFUSIVITY
#define FUS_HAVE_COMPUTE_DIFFUSIVITY
/// Lossy.hpp:376-380
template <typename T>
const T compute_diffusivity_of_sound(const T w0, const T c0, const T alpha) {
  return 2 * alpha * c0 * c0 * c0 / w0 / w0;
}
#endif
