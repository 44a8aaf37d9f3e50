// ref_driver2d.cpp -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
//
// 2-D quadrilateral cell loop on the reference's own 2-D tensor kernels: `contract<T,Na,Nb,Nk>` and
// `transpose<T,Na,Nb,offa,offb>` come from cpp/fenicsx-sf-naive/common/sum_factorisation.hpp,
// included UNMODIFIED by include path from /root/reference (never copied).  It sits in its own
// translation unit and namespace because that header and cpp/fenicsx-sf/common/
// sum_factorisation.hpp (ref_driver.cpp) define 3-D templates with identical signatures.
// The loop restates StiffnessSpectral2D::operator() (fenicsx-sf-naive spectral_op.hpp:275-318),
// whose surroundings need DOLFINx/Basix, call for call on plain arrays.  Used to validate
// fo_stiffness_apply_2d in fus_oracle.c.
#include <array>
#include <cstdint>

namespace naive {
#include "fenicsx-sf-naive/common/sum_factorisation.hpp" // -I/root/reference/cpp
}

namespace {
template <typename T, int P>
void stiffness2d_p(std::int64_t Nc, const std::int32_t* tensor_dofmap, const T* G_, const T* dphi_,
                   const T* coeffs, const T* x_array, T* y_array) {
  using namespace naive;
  constexpr int N = P + 1, Nd = N * N;
  std::array<T, Nd> fw0_, fw1_, x_, y0_, y1_, T1, T2, dphiT;
  T* fw0 = fw0_.data();
  T* fw1 = fw1_.data();
  T* dphiT_ = dphiT.data();
  transpose<T, N, N, 1, N>(dphi_, dphiT_);
  for (std::int64_t c = 0; c < Nc; ++c) {
    for (std::int32_t i = 0; i < Nd; ++i)
      x_[i] = x_array[tensor_dofmap[c * Nd + i]];
    T1.fill(0.0);
    T2.fill(0.0);
    fw0_.fill(0.0);
    contract<T, N, N, N>(x_.data(), dphi_, fw0);
    fw1_.fill(0.0);
    transpose<T, N, N, 1, N>(x_.data(), T1.data());
    contract<T, N, N, N>(T1.data(), dphi_, T2.data());
    transpose<T, N, N, 1, N>(T2.data(), fw1);
    // stiffness::transform, 2-D overload (spectral_op.hpp:196-208)
    const T* G = G_ + c * Nd * 3;
    const T coeff = coeffs[c];
    for (int iq = 0; iq < Nd; ++iq) {
      const T* _G = G + iq * 3;
      const T w0 = fw0[iq];
      const T w1 = fw1[iq];
      fw0[iq] = coeff * (_G[2] * w0 + _G[1] * w1);
      fw1[iq] = coeff * (_G[1] * w0 + _G[0] * w1);
    }
    T1.fill(0.0);
    T2.fill(0.0);
    y0_.fill(0.0);
    contract<T, N, N, N>(fw0, dphiT_, y0_.data());
    y1_.fill(0.0);
    transpose<T, N, N, 1, N>(fw1, T1.data());
    contract<T, N, N, N>(T1.data(), dphiT_, T2.data());
    transpose<T, N, N, 1, N>(T2.data(), y1_.data());
    for (std::int32_t i = 0; i < Nd; ++i)
      y_array[tensor_dofmap[c * Nd + i]] += y0_[i] + y1_[i];
  }
}
} // namespace

extern "C" void fr_stiffness_apply_2d(int P, std::int64_t nc, const std::int32_t* dofmap,
                                      const double* G, const double* dphi, const double* coeffs,
                                      const double* x, double* y) {
  switch (P) {
  case 1: stiffness2d_p<double, 1>(nc, dofmap, G, dphi, coeffs, x, y); break;
  case 2: stiffness2d_p<double, 2>(nc, dofmap, G, dphi, coeffs, x, y); break;
  case 3: stiffness2d_p<double, 3>(nc, dofmap, G, dphi, coeffs, x, y); break;
  case 4: stiffness2d_p<double, 4>(nc, dofmap, G, dphi, coeffs, x, y); break;
  case 5: stiffness2d_p<double, 5>(nc, dofmap, G, dphi, coeffs, x, y); break;
  case 6: stiffness2d_p<double, 6>(nc, dofmap, G, dphi, coeffs, x, y); break;
  case 7: stiffness2d_p<double, 7>(nc, dofmap, G, dphi, coeffs, x, y); break;
  default: break;
  }
}
