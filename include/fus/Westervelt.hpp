// fus/Westervelt.hpp -- drop-in for cpp/fenicsx-sf/common/Westervelt.hpp:56-413 of the reference.
#pragma once
#include "model_base.hpp"

/// Solver for the 3D second order Westervelt equation.
template <typename T, int P>
class WesterveltSpectral3D : public fus::detail::SpectralModel3D<T, P> {
public:
  WesterveltSpectral3D(basix::FiniteElement<T> element, std::shared_ptr<mesh::Mesh<T>> Mesh,
                       std::shared_ptr<mesh::MeshTags<std::int32_t>> FacetTags,
                       std::shared_ptr<fem::Function<T>> speedOfSound,
                       std::shared_ptr<fem::Function<T>> density,
                       std::shared_ptr<fem::Function<T>> diffusivityOfSound,
                       std::shared_ptr<fem::Function<T>> coefficientOfNonlinearity,
                       const T& sourceFrequency, const T& sourceAmplitude, const T& sourceSpeed)
      : fus::detail::SpectralModel3D<T, P>(FUS_WESTERVELT, element, Mesh, FacetTags, speedOfSound,
                                           density, diffusivityOfSound, coefficientOfNonlinearity,
                                           sourceFrequency, sourceAmplitude, sourceSpeed) {}
};

/// WesterveltSpectral2D<T,P> (cpp/fenicsx-sf-naive/common/Westervelt.hpp): quadrilateral mesh, same flow
template <typename T, int P>
class WesterveltSpectral2D : public fus::detail::SpectralModel3D<T, P> {
public:
  WesterveltSpectral2D(basix::FiniteElement<T> element, std::shared_ptr<mesh::Mesh<T>> Mesh,
                  std::shared_ptr<mesh::MeshTags<std::int32_t>> FacetTags,
                  std::shared_ptr<fem::Function<T>> speedOfSound,
                  std::shared_ptr<fem::Function<T>> density,
                  std::shared_ptr<fem::Function<T>> diffusivityOfSound,
                  std::shared_ptr<fem::Function<T>> coefficientOfNonlinearity,
                  const T& sourceFrequency, const T& sourceAmplitude, const T& sourceSpeed)
      : fus::detail::SpectralModel3D<T, P>(FUS_WESTERVELT, element, Mesh, FacetTags, speedOfSound, density,
                                           diffusivityOfSound, coefficientOfNonlinearity, sourceFrequency, sourceAmplitude,
                                           sourceSpeed) {
    if (Mesh->topology()->dim() != 2)
      throw std::runtime_error("WesterveltSpectral2D: quadrilateral mesh expected");
  }
};

#ifndef FUS_HAVE_COMPUTE_DIFFUSIVITY
#define FUS_HAVE_COMPUTE_DIFFUSIVITY
/// Westervelt.hpp:408-413 (the reference defines the same function in Lossy.hpp:376-380 as well,
/// so its two headers cannot be included together; here whichever comes first defines it)
template <typename T>
const T compute_diffusivity_of_sound(const T w0, const T c0, const T alpha) {
  return 2 * alpha * c0 * c0 * c0 / w0 / w0;
}
#endif
