// float_operators.cpp -- the reference's operator acceptance set-up with T = float
// (cpp/fenicsx-sf/tests/test_operators3d/main.cpp:13,59-79: P = 4, u = sin(x) cos(pi y),
// c0 = 1.5e-3, rho0 = 1e-3) against the drop-in headers: MassSpectral3D<float,P> and
// StiffnessSpectral3D<float,P> next to their double instantiations.  The float classes run the FP32
// instantiation of the device kernels, the double classes the FP64 one (see fus/spectral_op.hpp).
//
//   ./float_operators [cells_per_direction=6]
#include <fus/spectral_op.hpp>

#include <cstdio>
#include <cstdlib>

template <typename T>
static void run(std::size_t n, const char* name) {
  constexpr int P = 4;
  auto mesh = std::make_shared<mesh::Mesh<T>>(mesh::create_box<T>(
      {{{T(0), T(0), T(0)}, {T(1), T(1), T(1)}}}, {n, n, n}, mesh::CellType::hexahedron));
  auto element = basix::create_element<T>(basix::element::family::P, basix::cell::type::hexahedron,
                                          P, basix::element::lagrange_variant::gll_warped,
                                          basix::element::dpc_variant::unset, false);
  auto V = std::make_shared<fem::FunctionSpace<T>>(fem::create_functionspace(mesh, element));
  const std::size_t nd = V->dofmap()->index_map->size_local(), nc = n * n * n;
  // the reference interpolates u = sin(x) cos(pi y); the shim has no interpolation, so a bounded
  // oscillating function of the dof index stands in (the comparison below is float vs double on the
  // same data, not against an exact field)
  la::Vector<T> u(V->dofmap()->index_map, 1), ym(V->dofmap()->index_map, 1),
      ys(V->dofmap()->index_map, 1);
  auto ua = u.mutable_array();
  for (std::size_t i = 0; i < nd; ++i)
    ua[i] = (T)(std::sin(0.37 * (double)i) * std::cos(0.011 * (double)i));
  const T c0 = T(1.5e-3), rho0 = T(1e-3);
  std::vector<T> m_coeffs(nc, T(1) / rho0 / c0 / c0), s_coeffs(nc, T(-1) / rho0); // main.cpp:86-88,129-131
  MassSpectral3D<T, P> mass(V);
  StiffnessSpectral3D<T, P> stiffness(V);
  mass(u, m_coeffs, ym);
  stiffness(u, s_coeffs, ys);
  double m2 = 0, s2 = 0;
  for (T v : ym.array())
    m2 += (double)v * (double)v;
  for (T v : ys.array())
    s2 += (double)v * (double)v;
  std::printf("%s_mass_l2: %.17g\n%s_stiffness_l2: %.17g\n", name, std::sqrt(m2), name, std::sqrt(s2));
}

int main(int argc, char* argv[]) {
  const std::size_t n = argc > 1 ? std::atoi(argv[1]) : 6;
  run<float>(n, "float");
  run<double>(n, "double");
  return 0;
}
