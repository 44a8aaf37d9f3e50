// linear_box.cpp -- the reference's planar-wave driver (cpp/fenicsx-sf-naive/benchmarks/PH1/SC2-BM1/
// main.cpp:25-149) on a synthetic box, written against the drop-in headers of this repository.
// Only the #include lines and the mesh source differ from a driver written for the reference.
//
//   g++ -std=c++20 -O2 -Iinclude examples/linear_box.cpp -Lfenicsx-fus_b200/lib -lfus_b200
//       -Wl,-rpath,$PWD/fenicsx-fus_b200/lib -o examples/linear_box
//   ./linear_box [cells_per_direction=8] [steps=20]
#include <fus/Linear.hpp>
#include <fus/Lossy.hpp>
#include <fus/Westervelt.hpp>

#include <chrono>
#include <cstdio>
#include <cstdlib>

using T = double;

int main(int argc, char* argv[]) {
  const std::size_t n = argc > 1 ? std::atoi(argv[1]) : 8;
  const int nsteps = argc > 2 ? std::atoi(argv[2]) : 20;

  // Source parameters (SC2-BM1/main.cpp:32-44)
  const T sourceFrequency = 0.5e6; // (Hz)
  const T sourceAmplitude = 60000; // (Pa)
  const T period = 1 / sourceFrequency;
  const T speedOfSound = 1500; // (m/s)
  const T density = 1000;      // (kg/m^3)
  const T domainLength = 0.12 * n / 54.0;
  constexpr int degreeOfBasis = 4;

  auto mesh = std::make_shared<mesh::Mesh<T>>(mesh::create_box<T>(
      {{{0.0, 0.0, 0.0}, {domainLength, domainLength, domainLength}}}, {n, n, n},
      mesh::CellType::hexahedron));
  auto mt_facet = std::make_shared<mesh::MeshTags<std::int32_t>>(mesh::box_facet_tags(*mesh));

  auto element = basix::create_element<T>(basix::element::family::P, basix::cell::type::hexahedron,
                                          degreeOfBasis, basix::element::lagrange_variant::gll_warped,
                                          basix::element::dpc_variant::unset, false);
  auto V_DG = std::make_shared<fem::FunctionSpace<T>>(
      fem::create_functionspace(mesh, basix::FiniteElement<T>(0)));
  auto c0 = std::make_shared<fem::Function<T>>(V_DG);
  auto rho0 = std::make_shared<fem::Function<T>>(V_DG);
  std::span<T> c0_ = c0->x()->mutable_array();
  std::fill(c0_.begin(), c0_.end(), speedOfSound);
  std::span<T> rho0_ = rho0->x()->mutable_array();
  std::fill(rho0_.begin(), rho0_.end(), density);

  // Temporal parameters (SC2-BM1/main.cpp:87-94); mesh::h of a cube cell is sqrt(3) h
  const T CFL = 0.65;
  const T meshSizeMinGlobal = std::sqrt(3.0) * domainLength / n;
  T timeStepSize = CFL * meshSizeMinGlobal / (speedOfSound * degreeOfBasis * degreeOfBasis);
  const int stepPerPeriod = period / timeStepSize + 1;
  timeStepSize = period / stepPerPeriod;
  const T startTime = 0.0;
  const T finalTime = startTime + (nsteps - 0.5) * timeStepSize;

  auto model = LinearSpectral3D<T, degreeOfBasis>(element, mesh, mt_facet, c0, rho0, sourceFrequency,
                                                  sourceAmplitude, speedOfSound);
  std::printf("Degrees of freedom: %lld\n", (long long)model.number_of_dofs());
  std::printf("Time step size: %.17g\n", timeStepSize);

  model.init();
  auto t0 = std::chrono::steady_clock::now();
  model.rk4(startTime, finalTime, timeStepSize);
  const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  std::printf("Number of steps: %d\n", model.number_of_steps());
  std::printf("Solve time: %g\nTime per step: %g\n", el, el / model.number_of_steps());

  auto u = model.u_sol()->x()->array();
  double s2 = 0.0, s1 = 0.0;
  for (double v : u) {
    s2 += v * v;
    s1 += v;
  }
  std::printf("u_l2: %.17g\nu_sum: %.17g\n", std::sqrt(s2), s1);

  // one stiffness application through the operator class (spectral_op.hpp:173-243)
  auto Vp = std::make_shared<fem::FunctionSpace<T>>(fem::create_functionspace(mesh, element));
  StiffnessSpectral3D<T, degreeOfBasis> stiffness(Vp);
  la::Vector<T> x(Vp->dofmap()->index_map, 1), y(Vp->dofmap()->index_map, 1);
  auto xa = x.mutable_array();
  for (std::size_t i = 0; i < xa.size(); ++i)
    xa[i] = std::sin(0.001 * (double)i);
  std::vector<T> s_coeffs(c0_.size());
  for (std::size_t i = 0; i < s_coeffs.size(); ++i)
    s_coeffs[i] = -1.0 / rho0_[i];
  stiffness(x, s_coeffs, y);
  double k2 = 0.0;
  for (double v : y.array())
    k2 += v * v;
  std::printf("Kx_l2: %.17g\n", std::sqrt(k2));
  return 0;
}
