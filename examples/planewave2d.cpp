// planewave2d.cpp -- the reference's 2-D example driver
// (cpp/fenicsx-sf-naive/examples/linear_planewave2d_1/main.cpp:31-165) on a synthetic rectangle,
// written against the drop-in headers of this repository: LinearSpectral2D, and one application of
// StiffnessSpectral2D / MassSpectral2D (cpp/fenicsx-sf-naive/common/spectral_op.hpp).
//
//   ./planewave2d [cells_per_direction=8] [steps=20]
#include <fus/Linear.hpp>

#include <cstdio>
#include <cstdlib>

using T = double;

int main(int argc, char* argv[]) {
  const std::size_t n = argc > 1 ? std::atoi(argv[1]) : 8;
  const int nsteps = argc > 2 ? std::atoi(argv[2]) : 20;

  // Source, material and domain parameters (main.cpp:31-41)
  const T sourceFrequency = 0.5e6, sourceAmplitude = 60000, period = 1 / sourceFrequency;
  const T speedOfSound = 1500, density = 1000;
  const T domainLength = 0.12 * n / 54.0;
  constexpr int degreeOfBasis = 4;

  auto mesh = std::make_shared<mesh::Mesh<T>>(mesh::create_rectangle<T>(
      {{{0.0, 0.0}, {domainLength, domainLength}}}, {n, n}, mesh::CellType::quadrilateral));
  mesh->topology()->create_connectivity(1, 2);
  auto mt_cell = std::make_shared<mesh::MeshTags<std::int32_t>>(mesh::box_cell_layers(*mesh, 1));
  auto mt_facet = std::make_shared<mesh::MeshTags<std::int32_t>>(mesh::box_facet_tags(*mesh));

  // Mesh parameters (main.cpp:57-69)
  const int tdim = mesh->topology()->dim();
  const int num_cell = mesh->topology()->index_map(tdim)->size_local();
  std::vector<int> num_cell_range(num_cell);
  std::iota(num_cell_range.begin(), num_cell_range.end(), 0);
  std::vector<T> mesh_size_local = mesh::h(*mesh, num_cell_range, tdim);
  const T meshSizeMinGlobal = *std::min_element(mesh_size_local.begin(), mesh_size_local.end());

  auto element = basix::create_element<T>(
      basix::element::family::P, basix::cell::type::quadrilateral, degreeOfBasis,
      basix::element::lagrange_variant::gll_warped, basix::element::dpc_variant::unset, false);
  auto V_DG = std::make_shared<fem::FunctionSpace<T>>(
      fem::create_functionspace(mesh, basix::FiniteElement<T>(0)));
  auto c0 = std::make_shared<fem::Function<T>>(V_DG);
  auto rho0 = std::make_shared<fem::Function<T>>(V_DG);
  auto cells_1 = mt_cell->find(1);
  std::span<T> c0_ = c0->x()->mutable_array();
  std::for_each(cells_1.begin(), cells_1.end(), [&](std::int32_t& i) { c0_[i] = speedOfSound; });
  std::span<T> rho0_ = rho0->x()->mutable_array();
  std::for_each(cells_1.begin(), cells_1.end(), [&](std::int32_t& i) { rho0_[i] = density; });

  // Temporal parameters (main.cpp:103-110)
  const T CFL = 0.9;
  T timeStepSize = CFL * meshSizeMinGlobal / (speedOfSound * degreeOfBasis * degreeOfBasis);
  const int stepPerPeriod = period / timeStepSize + 1;
  timeStepSize = period / stepPerPeriod;
  const T startTime = 0.0, finalTime = startTime + (nsteps - 0.5) * timeStepSize;

  auto model = LinearSpectral2D<T, degreeOfBasis>(element, mesh, mt_facet, c0, rho0,
                                                  sourceFrequency, sourceAmplitude, speedOfSound);
  std::printf("Degrees of freedom: %lld\n", (long long)model.number_of_dofs());
  std::printf("Time step size: %.17g\n", timeStepSize);
  model.init();
  model.rk4(startTime, finalTime, timeStepSize);
  std::printf("Number of steps: %d\n", model.number_of_steps());
  double s2 = 0.0;
  for (double v : model.u_sol()->x()->array())
    s2 += v * v;
  std::printf("u_l2: %.17g\n", std::sqrt(s2));

  // one application of each 2-D operator class
  auto Vp = std::make_shared<fem::FunctionSpace<T>>(fem::create_functionspace(mesh, element));
  StiffnessSpectral2D<T, degreeOfBasis> stiffness(Vp);
  MassSpectral2D<T, degreeOfBasis> mass(Vp);
  la::Vector<T> x(Vp->dofmap()->index_map, 1), y(Vp->dofmap()->index_map, 1),
      z(Vp->dofmap()->index_map, 1);
  auto xa = x.mutable_array();
  for (std::size_t i = 0; i < xa.size(); ++i)
    xa[i] = std::sin(0.01 * (double)i);
  std::vector<T> coeffs(c0_.size(), -1.0e-3);
  stiffness(x, coeffs, y);
  mass(x, coeffs, z);
  double k2 = 0.0, m2 = 0.0;
  for (double v : y.array())
    k2 += v * v;
  for (double v : z.array())
    m2 += v * v;
  std::printf("Kx_l2: %.17g\nMx_l2: %.17g\n", std::sqrt(k2), std::sqrt(m2));
  return 0;
}
