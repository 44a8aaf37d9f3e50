// media_box.cpp -- the reference's lossy and Westervelt benchmark drivers
// (cpp/fenicsx-sf/benchmarks/PH1/BM7-SC1/main.cpp:31-152 and HITU/W-H131-WATER/main.cpp:32-150)
// on a synthetic box, written against the drop-in headers of this repository.  The material
// set-up through cell tags, the time-step arithmetic and the solver calls are the reference's;
// only the #include lines and the mesh source differ (its meshes are not distributed).
//
//   ./media_box lossy|westervelt [cells_per_direction=6] [steps=10]
#include <fus/Lossy.hpp>
#include <fus/Westervelt.hpp>

#include <cstdio>
#include <cstdlib>
#include <cstring>

using T = double;

template <typename Model>
static void solve_and_report(Model& model, T startTime, T finalTime, T timeStepSize) {
  std::printf("Degrees of freedom: %lld\n", (long long)model.number_of_dofs());
  std::printf("Time step size: %.17g\n", timeStepSize);
  model.init();
  model.rk4(startTime, finalTime, timeStepSize);
  std::printf("Number of steps: %d\n", model.number_of_steps());
  double s2 = 0.0, v2 = 0.0;
  for (double v : model.u_sol()->x()->array())
    s2 += v * v;
  for (double v : model.v_sol()->x()->array())
    v2 += v * v;
  std::printf("u_l2: %.17g\nv_l2: %.17g\n", std::sqrt(s2), std::sqrt(v2));
}

int main(int argc, char* argv[]) {
  const bool westervelt = argc > 1 && !std::strcmp(argv[1], "westervelt");
  const std::size_t n = argc > 2 ? std::atoi(argv[2]) : 6;
  const int nsteps = argc > 3 ? std::atoi(argv[3]) : 10;

  auto make_mesh = [&](T domainLength) {
    return std::make_shared<mesh::Mesh<T>>(mesh::create_box<T>(
        {{{0.0, 0.0, 0.0}, {domainLength, domainLength, domainLength}}}, {n, n, n},
        mesh::CellType::hexahedron));
  };
  auto min_mesh_size = [](const mesh::Mesh<T>& mesh) {
    const int num_cell = mesh.topology()->index_map(3)->size_local();
    std::vector<int> num_cell_range(num_cell);
    std::iota(num_cell_range.begin(), num_cell_range.end(), 0);
    std::vector<T> mesh_size_local = mesh::h(mesh, num_cell_range, 3);
    return *std::min_element(mesh_size_local.begin(), mesh_size_local.end());
  };

  if (!westervelt) {
    // ---- BM7-SC1/main.cpp:31-118: water | cortical bone, attenuation in the bone -------------
    const T sourceFrequency = 0.5e6, sourceAmplitude = 60000, period = 1 / sourceFrequency;
    const T angularFrequency = 2 * M_PI * sourceFrequency;
    const T speedOfSoundWater = 1500.0, speedOfSoundCortBone = 2800.0;
    const T densityWater = 1000.0, densityCortBone = 1850.0;
    const T attenuationCoefficientdBCortBone = 400.0;
    const T attenuationCoefficientNpCortBone = attenuationCoefficientdBCortBone / 20 * std::log(10);
    const T diffusivityOfSoundCortBone = compute_diffusivity_of_sound(
        angularFrequency, speedOfSoundCortBone, attenuationCoefficientNpCortBone);
    const T domainLength = 0.12 * n / 54.0;
    constexpr int degreeOfBasis = 4;

    auto mesh = make_mesh(domainLength);
    auto mt_cell = std::make_shared<mesh::MeshTags<std::int32_t>>(mesh::box_cell_layers(*mesh, 2));
    auto mt_facet = std::make_shared<mesh::MeshTags<std::int32_t>>(mesh::box_facet_tags(*mesh));
    const T meshSizeMinGlobal = min_mesh_size(*mesh);

    auto V_DG = std::make_shared<fem::FunctionSpace<T>>(
        fem::create_functionspace(mesh, basix::FiniteElement<T>(0)));
    auto c0 = std::make_shared<fem::Function<T>>(V_DG);
    auto rho0 = std::make_shared<fem::Function<T>>(V_DG);
    auto delta0 = std::make_shared<fem::Function<T>>(V_DG);
    auto cells_1 = mt_cell->find(1);
    auto cells_2 = mt_cell->find(2);
    std::span<T> c0_ = c0->x()->mutable_array();
    std::for_each(cells_1.begin(), cells_1.end(), [&](std::int32_t& i) { c0_[i] = speedOfSoundWater; });
    std::for_each(cells_2.begin(), cells_2.end(),
                  [&](std::int32_t& i) { c0_[i] = speedOfSoundCortBone; });
    std::span<T> rho0_ = rho0->x()->mutable_array();
    std::for_each(cells_1.begin(), cells_1.end(), [&](std::int32_t& i) { rho0_[i] = densityWater; });
    std::for_each(cells_2.begin(), cells_2.end(), [&](std::int32_t& i) { rho0_[i] = densityCortBone; });
    std::span<T> delta0_ = delta0->x()->mutable_array();
    std::for_each(cells_1.begin(), cells_1.end(), [&](std::int32_t& i) { delta0_[i] = 0.0; });
    std::for_each(cells_2.begin(), cells_2.end(),
                  [&](std::int32_t& i) { delta0_[i] = diffusivityOfSoundCortBone; });

    const T CFL = 0.15;
    T timeStepSize
        = CFL * meshSizeMinGlobal / (speedOfSoundCortBone * degreeOfBasis * degreeOfBasis);
    const int stepPerPeriod = period / timeStepSize + 1;
    timeStepSize = period / stepPerPeriod;
    const T startTime = 0.0, finalTime = startTime + (nsteps - 0.5) * timeStepSize;

    auto element = basix::create_element<T>(
        basix::element::family::P, basix::cell::type::hexahedron, degreeOfBasis,
        basix::element::lagrange_variant::gll_warped, basix::element::dpc_variant::unset, false);
    auto model = LossySpectral3D<T, degreeOfBasis>(element, mesh, mt_facet, c0, rho0, delta0,
                                                   sourceFrequency, sourceAmplitude,
                                                   speedOfSoundWater);
    std::printf("Model: lossy\nDiffusivity of sound: %.17g\n", diffusivityOfSoundCortBone);
    solve_and_report(model, startTime, finalTime, timeStepSize);
  } else {
    // ---- W-H131-WATER/main.cpp:32-118: water, weak attenuation, beta = 3.5, degree 6 -----------
    const T speedOfSound = 1480.0, density = 1000.0, sourceFrequency = 1.1e6;
    const T sourceVelocity = 0.2726428, sourceAmplitude = density * speedOfSound * sourceVelocity;
    const T period = 1 / sourceFrequency, angularFrequency = 2 * M_PI * sourceFrequency;
    const T nonlinearCoefficient = 3.5, attenuationCoefficientdB = 0.2;
    const T attenuationCoefficientNp = attenuationCoefficientdB / 20 * std::log(10);
    const T diffusivityOfSound
        = compute_diffusivity_of_sound(angularFrequency, speedOfSound, attenuationCoefficientNp);
    const T domainLength = 0.08 * n / 100.0;
    constexpr int degreeOfBasis = 6;

    auto mesh = make_mesh(domainLength);
    auto mt_cell = std::make_shared<mesh::MeshTags<std::int32_t>>(mesh::box_cell_layers(*mesh, 1));
    auto mt_facet = std::make_shared<mesh::MeshTags<std::int32_t>>(mesh::box_facet_tags(*mesh));
    const T meshSizeMinGlobal = min_mesh_size(*mesh);

    auto V_DG = std::make_shared<fem::FunctionSpace<T>>(
        fem::create_functionspace(mesh, basix::FiniteElement<T>(0)));
    auto c0 = std::make_shared<fem::Function<T>>(V_DG);
    auto rho0 = std::make_shared<fem::Function<T>>(V_DG);
    auto delta0 = std::make_shared<fem::Function<T>>(V_DG);
    auto beta0 = std::make_shared<fem::Function<T>>(V_DG);
    auto cells_1 = mt_cell->find(1);
    std::span<T> c0_ = c0->x()->mutable_array();
    std::for_each(cells_1.begin(), cells_1.end(), [&](std::int32_t& i) { c0_[i] = speedOfSound; });
    std::span<T> rho0_ = rho0->x()->mutable_array();
    std::for_each(cells_1.begin(), cells_1.end(), [&](std::int32_t& i) { rho0_[i] = density; });
    std::span<T> beta0_ = beta0->x()->mutable_array();
    std::for_each(cells_1.begin(), cells_1.end(),
                  [&](std::int32_t& i) { beta0_[i] = nonlinearCoefficient; });
    std::span<T> delta0_ = delta0->x()->mutable_array();
    std::for_each(cells_1.begin(), cells_1.end(),
                  [&](std::int32_t& i) { delta0_[i] = diffusivityOfSound; });

    // the reference runs CFL = 0.55 on its bowl mesh; on a box the all-facet absorbing term of
    // the lossy/Westervelt forms limits the stable step, hence 0.2
    const T CFL = 0.2;
    T timeStepSize = CFL * meshSizeMinGlobal / (speedOfSound * degreeOfBasis * degreeOfBasis);
    const int stepPerPeriod = period / timeStepSize + 1;
    timeStepSize = period / stepPerPeriod;
    const T startTime = 0.0, finalTime = startTime + (nsteps - 0.5) * timeStepSize;

    auto element = basix::create_element<T>(
        basix::element::family::P, basix::cell::type::hexahedron, degreeOfBasis,
        basix::element::lagrange_variant::gll_warped, basix::element::dpc_variant::unset, false);
    auto model = WesterveltSpectral3D<T, degreeOfBasis>(element, mesh, mt_facet, c0, rho0, delta0,
                                                        beta0, sourceFrequency, sourceAmplitude,
                                                        speedOfSound);
    std::printf("Model: westervelt\nDiffusivity of sound: %.17g\n", diffusivityOfSound);
    solve_and_report(model, startTime, finalTime, timeStepSize);
  }
  return 0;
}
