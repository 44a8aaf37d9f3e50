// ref_driver.cpp -- CPU ORACLE / CPU BASELINE (TEST INFRASTRUCTURE, NOT PRODUCT CODE)
//
// Builds oracle/_ref/libfus_ref*.so.  The tensor kernels `contract<>` / `transpose<>` come from
// the reference's own header cpp/fenicsx-sf/common/sum_factorisation.hpp, included UNMODIFIED
// by include path from /root/reference (never copied into this repository).  The rest of the
// reference's operator (spectral_op.hpp) needs DOLFINx/Basix, which are not installed, so the
// cell loops below restate spectral_op.hpp:75-85 (mass) and :183-242 (stiffness) making the
// same calls in the same order on plain arrays.
//
// Two uses:
//   * validate oracle/fus_oracle.c (tests compare the two on identical inputs);
//   * the timed CPU baseline / `bench.py --impl reference`: the same cell loop run by one
//     OpenMP thread per contiguous cell range ("rank"), each with private scratch and a private
//     window of the output vector that is reduced afterwards -- the shared-memory stand-in for
//     the reference's MPI ranks + scatter_rev (Linear.hpp:206).
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "sum_factorisation.hpp" // resolved with -I/root/reference/cpp/fenicsx-sf/common

namespace {

template <typename T, int P>
struct StiffnessCell {
  static constexpr int N = P + 1;
  static constexpr int Nd = N * N * N;
  std::array<T, Nd> fw0_, fw1_, fw2_, x_, y0_, y1_, y2_, T1, T2, T3, T4;

  // one cell of spectral_op.hpp:183-242
  template <typename Y>
  inline void operator()(const std::int32_t* dofs, const T* G, T coeff, const T* dphi_,
                         const T* x_array, Y&& y_add) {
    T* fw0 = fw0_.data();
    T* fw1 = fw1_.data();
    T* fw2 = fw2_.data();
    for (std::int32_t i = 0; i < Nd; ++i)
      x_[i] = x_array[dofs[i]];
    T1.fill(0.0);
    T2.fill(0.0);
    T3.fill(0.0);
    T4.fill(0.0);
    fw0_.fill(0.0);
    contract<T, N, N, N, N, true>(dphi_, x_.data(), fw0);
    fw1_.fill(0.0);
    transpose<T, N, N, N, N, N * N, 1>(x_.data(), T1.data());
    contract<T, N, N, N, N, true>(dphi_, T1.data(), T2.data());
    transpose<T, N, N, N, N, N * N, 1>(T2.data(), fw1);
    fw2_.fill(0.0);
    transpose<T, N, N, N, 1, N, N * N>(x_.data(), T3.data());
    contract<T, N, N, N, N, true>(dphi_, T3.data(), T4.data());
    transpose<T, N, N, N, 1, N, N * N>(T4.data(), fw2);
    // stiffness::transform (spectral_op.hpp:113-130)
    for (int iq = 0; iq < Nd; ++iq) {
      const T* _G = G + iq * 6;
      const T w0 = fw0[iq];
      const T w1 = fw1[iq];
      const T w2 = fw2[iq];
      fw0[iq] = coeff * (_G[0] * w0 + _G[1] * w1 + _G[2] * w2);
      fw1[iq] = coeff * (_G[1] * w0 + _G[3] * w1 + _G[4] * w2);
      fw2[iq] = coeff * (_G[2] * w0 + _G[4] * w1 + _G[5] * w2);
    }
    T1.fill(0.0);
    T2.fill(0.0);
    T3.fill(0.0);
    T4.fill(0.0);
    y0_.fill(0.0);
    contract<T, N, N, N, N, false>(dphi_, fw0, y0_.data());
    y1_.fill(0.0);
    transpose<T, N, N, N, N, N * N, 1>(fw1, T1.data());
    contract<T, N, N, N, N, false>(dphi_, T1.data(), T2.data());
    transpose<T, N, N, N, N, N * N, 1>(T2.data(), y1_.data());
    y2_.fill(0.0);
    transpose<T, N, N, N, 1, N, N * N>(fw2, T3.data());
    contract<T, N, N, N, N, false>(dphi_, T3.data(), T4.data());
    transpose<T, N, N, N, 1, N, N * N>(T4.data(), y2_.data());
    for (std::int32_t i = 0; i < Nd; ++i)
      y_add(dofs[i], y0_[i] + y1_[i] + y2_[i]);
  }
};

int g_threads = 1;

// Cell ranges [c0,c1) per thread and the window [lo,hi] of dofs each range touches.
struct Ranges {
  std::vector<std::int64_t> c0, c1;
  std::vector<std::int32_t> lo, hi;
};

Ranges make_ranges(int nt, std::int64_t nc, int Nd, const std::int32_t* dofmap) {
  Ranges r;
  r.c0.resize(nt);
  r.c1.resize(nt);
  r.lo.resize(nt);
  r.hi.resize(nt);
  for (int t = 0; t < nt; ++t) {
    r.c0[t] = nc * t / nt;
    r.c1[t] = nc * (t + 1) / nt;
  }
#pragma omp parallel for num_threads(nt) schedule(static, 1)
  for (int t = 0; t < nt; ++t) {
    std::int32_t lo = INT32_MAX, hi = -1;
    for (std::int64_t i = r.c0[t] * Nd; i < r.c1[t] * Nd; ++i) {
      lo = std::min(lo, dofmap[i]);
      hi = std::max(hi, dofmap[i]);
    }
    r.lo[t] = lo;
    r.hi[t] = hi;
  }
  return r;
}

// After the cell loops: y[d] += sum over ranges covering d of their private window
void reduce_windows(int nt, const Ranges& r, const std::vector<std::vector<double>>& win,
                    double* y) {
  std::int32_t glo = INT32_MAX, ghi = -1;
  for (int t = 0; t < nt; ++t)
    if (r.hi[t] >= r.lo[t]) {
      glo = std::min(glo, r.lo[t]);
      ghi = std::max(ghi, r.hi[t]);
    }
#pragma omp parallel for num_threads(nt) schedule(static)
  for (std::int64_t d = glo; d <= ghi; ++d) {
    double s = 0.0;
    for (int t = 0; t < nt; ++t)
      if (d >= r.lo[t] && d <= r.hi[t])
        s += win[t][d - r.lo[t]];
    y[d] += s;
  }
}

template <int P>
void stiffness_p(std::int64_t nc, const std::int32_t* dofmap, const double* G, const double* dphi,
                 const double* coeffs, const double* x, double* y) {
  constexpr int Nd = (P + 1) * (P + 1) * (P + 1);
  int nt = std::max(1, (int)std::min<std::int64_t>(g_threads, nc));
  if (nt == 1) {
    StiffnessCell<double, P> cell;
    for (std::int64_t c = 0; c < nc; ++c)
      cell(dofmap + c * Nd, G + c * Nd * 6, coeffs[c], dphi, x,
           [&](std::int32_t d, double v) { y[d] += v; });
    return;
  }
  Ranges r = make_ranges(nt, nc, Nd, dofmap);
  std::vector<std::vector<double>> win(nt);
#pragma omp parallel num_threads(nt)
  {
#ifdef _OPENMP
    int t = omp_get_thread_num();
#else
    int t = 0;
#endif
    win[t].assign(r.hi[t] >= r.lo[t] ? (std::size_t)(r.hi[t] - r.lo[t] + 1) : 0, 0.0);
    double* w = win[t].data();
    const std::int32_t lo = r.lo[t];
    StiffnessCell<double, P> cell;
    for (std::int64_t c = r.c0[t]; c < r.c1[t]; ++c)
      cell(dofmap + c * Nd, G + c * Nd * 6, coeffs[c], dphi, x,
           [&](std::int32_t d, double v) { w[d - lo] += v; });
  }
  reduce_windows(nt, r, win, y);
}

template <int P>
void mass_p(std::int64_t nc, const std::int32_t* dofmap, const double* detJ, const double* coeffs,
            const double* x, double* y) {
  constexpr int Nd = (P + 1) * (P + 1) * (P + 1);
  int nt = std::max(1, (int)std::min<std::int64_t>(g_threads, nc));
  auto one_cell = [&](std::int64_t c, std::array<double, Nd>& x_, auto&& y_add) {
    // spectral_op.hpp:75-85 with mass::transform :19-26
    for (std::int32_t i = 0; i < Nd; ++i)
      x_[i] = x[dofmap[c * Nd + i]];
    const double* sdetJ = detJ + c * Nd;
    for (int iq = 0; iq < Nd; ++iq)
      x_[iq] = coeffs[c] * x_[iq] * sdetJ[iq];
    for (std::int32_t i = 0; i < Nd; ++i)
      y_add(dofmap[c * Nd + i], x_[i]);
  };
  if (nt == 1) {
    std::array<double, Nd> x_;
    for (std::int64_t c = 0; c < nc; ++c)
      one_cell(c, x_, [&](std::int32_t d, double v) { y[d] += v; });
    return;
  }
  Ranges r = make_ranges(nt, nc, Nd, dofmap);
  std::vector<std::vector<double>> win(nt);
#pragma omp parallel num_threads(nt)
  {
#ifdef _OPENMP
    int t = omp_get_thread_num();
#else
    int t = 0;
#endif
    win[t].assign(r.hi[t] >= r.lo[t] ? (std::size_t)(r.hi[t] - r.lo[t] + 1) : 0, 0.0);
    double* w = win[t].data();
    const std::int32_t lo = r.lo[t];
    std::array<double, Nd> x_;
    for (std::int64_t c = r.c0[t]; c < r.c1[t]; ++c)
      one_cell(c, x_, [&](std::int32_t d, double v) { w[d - lo] += v; });
  }
  reduce_windows(nt, r, win, y);
}

} // namespace

extern "C" {

void fr_set_threads(int n) {
  g_threads = n < 1 ? 1 : n;
#ifdef _OPENMP
  omp_set_num_threads(g_threads);
#endif
}

int fr_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// Same signature as fo_stiffness_apply / fo_mass_apply in fus_oracle.c
void fr_stiffness_apply(int P, std::int64_t nc, const std::int32_t* dofmap, const double* G,
                        const double* dphi, const double* coeffs, const double* x, double* y) {
  switch (P) {
  case 1: stiffness_p<1>(nc, dofmap, G, dphi, coeffs, x, y); break;
  case 2: stiffness_p<2>(nc, dofmap, G, dphi, coeffs, x, y); break;
  case 3: stiffness_p<3>(nc, dofmap, G, dphi, coeffs, x, y); break;
  case 4: stiffness_p<4>(nc, dofmap, G, dphi, coeffs, x, y); break;
  case 5: stiffness_p<5>(nc, dofmap, G, dphi, coeffs, x, y); break;
  case 6: stiffness_p<6>(nc, dofmap, G, dphi, coeffs, x, y); break;
  case 7: stiffness_p<7>(nc, dofmap, G, dphi, coeffs, x, y); break;
  default: break;
  }
}

void fr_mass_apply(int P, std::int64_t nc, const std::int32_t* dofmap, const double* detJ,
                   const double* coeffs, const double* x, double* y) {
  switch (P) {
  case 1: mass_p<1>(nc, dofmap, detJ, coeffs, x, y); break;
  case 2: mass_p<2>(nc, dofmap, detJ, coeffs, x, y); break;
  case 3: mass_p<3>(nc, dofmap, detJ, coeffs, x, y); break;
  case 4: mass_p<4>(nc, dofmap, detJ, coeffs, x, y); break;
  case 5: mass_p<5>(nc, dofmap, detJ, coeffs, x, y); break;
  case 6: mass_p<6>(nc, dofmap, detJ, coeffs, x, y); break;
  case 7: mass_p<7>(nc, dofmap, detJ, coeffs, x, y); break;
  default: break;
  }
}

// The reference's own known-answer test (cpp/mwe/sum_factorisation/main.cpp:10-62):
// contract<double,2,3,2,2,true> on iota data, then transpose<double,3,2,2,2,1,6>.
void fr_kat(double* out12, double* out_t12) {
  constexpr int M = 3, N = 2;
  std::array<double, N * N * N> x;
  std::array<double, M * N> dphi;
  for (int i = 0; i < N * N * N; i++)
    x[i] = i;
  for (int i = 0; i < M * N; i++)
    dphi[i] = i;
  std::array<double, M * N * N> out{0}, out_t{0};
  contract<double, N, M, N, N, true>(dphi.data(), x.data(), out.data());
  transpose<double, M, N, N, N, 1, M * N>(out.data(), out_t.data());
  std::copy(out.begin(), out.end(), out12);
  std::copy(out_t.begin(), out_t.end(), out_t12);
}

// Generic-size entry to the reference kernels for cross-checking fo_contract/fo_transpose
// at the operator's own sizes (N = P+1, P = 2..7), both contraction flavours.
void fr_contract_cube(int N, int transposeA, const double* A, const double* B, double* C) {
#define FR_CASE(n)                                                                                 \
  case n:                                                                                          \
    if (transposeA)                                                                                \
      contract<double, n, n, n, n, true>(A, B, C);                                                 \
    else                                                                                           \
      contract<double, n, n, n, n, false>(A, B, C);                                                \
    break;
  switch (N) {
    FR_CASE(2) FR_CASE(3) FR_CASE(4) FR_CASE(5) FR_CASE(6) FR_CASE(7) FR_CASE(8)
  default: break;
  }
#undef FR_CASE
}

// which: 0 -> transpose<N,N,N, N,N*N,1> (swap first two), 1 -> transpose<N,N,N, 1,N,N*N> (reverse)
void fr_transpose_cube(int N, int which, double* A, double* B) {
#define FR_CASE(n)                                                                                 \
  case n:                                                                                          \
    if (which == 0)                                                                                \
      transpose<double, n, n, n, n, n * n, 1>(A, B);                                               \
    else                                                                                           \
      transpose<double, n, n, n, 1, n, n * n>(A, B);                                               \
    break;
  switch (N) {
    FR_CASE(2) FR_CASE(3) FR_CASE(4) FR_CASE(5) FR_CASE(6) FR_CASE(7) FR_CASE(8)
  default: break;
  }
#undef FR_CASE
}

} // extern "C"
