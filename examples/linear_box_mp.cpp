// linear_box_mp.cpp -- the reference's planar-wave driver (cpp/fenicsx-sf-naive/benchmarks/PH1/SC2-BM1/
// main.cpp:25-149) on a box PARTITIONED over several GPUs, written against the drop-in headers.
// The reference runs one MPI rank per partition (`mpirun -n N`); there is no MPI in this image, so
// the ranks are the threads of this process, one per GPU (fus::run_ranks, include/fus/
// dolfinx_shim.hpp), and everything between the comments "rank body" is what a rank of the
// reference's driver does: create_box on the communicator, tags, coefficient functions, the
// global minimum of the mesh size for the time step (BM7-SC1/main.cpp:72-78), the solver, rk4.
// The ghost exchanges of the time loop run between the GPUs (fused peer transport over NVLink).
//
//   g++ -std=c++20 -O2 -pthread -Iinclude examples/linear_box_mp.cpp -Lfenicsx-fus_b200/lib
//       -lfus_b200 -Wl,-rpath,$PWD/fenicsx-fus_b200/lib -o examples/linear_box_mp
//   ./linear_box_mp px py pz cells_per_direction steps [model=linear|lossy|westervelt] [out_prefix]
// With out_prefix every rank writes its owned dofs as (int64 global index, double u) pairs to
// <out_prefix>.<rank>.bin -- what tests/test_gpu_multi.py compares with the single-domain oracle.
#include <fus/Linear.hpp>
#include <fus/Lossy.hpp>
#include <fus/Westervelt.hpp>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

using T = double;

int main(int argc, char* argv[]) {
  if (argc < 6) {
    std::fprintf(stderr, "usage: %s px py pz cells_per_direction steps [model] [out_prefix]\n", argv[0]);
    return 2;
  }
  const std::array<int, 3> pgrid{std::atoi(argv[1]), std::atoi(argv[2]), std::atoi(argv[3])};
  const std::size_t n = std::atoi(argv[4]);
  const int nsteps = std::atoi(argv[5]);
  const std::string kind = argc > 6 ? argv[6] : "linear";
  const std::string out = argc > 7 ? argv[7] : "";
  const int nranks = pgrid[0] * pgrid[1] * pgrid[2];

  fus::run_ranks(nranks, pgrid, [&](fus::Comm& comm) {
    // ---- rank body ---------------------------------------------------------------------------
    const T sourceFrequency = 0.5e6, sourceAmplitude = 60000, period = 1 / sourceFrequency;
    const T speedOfSound = 1500, density = 1000;
    const T domainLength = 0.12 * n / 54.0;
    constexpr int degreeOfBasis = 4;

    auto mesh = std::make_shared<mesh::Mesh<T>>(mesh::create_box<T>(
        comm, {{{0.0, 0.0, 0.0}, {domainLength, domainLength, domainLength}}}, {n, n, n},
        mesh::CellType::hexahedron));
    auto mt_facet = std::make_shared<mesh::MeshTags<std::int32_t>>(mesh::box_facet_tags(*mesh));
    auto element = basix::create_element<T>(basix::element::family::P, basix::cell::type::hexahedron,
                                            degreeOfBasis, basix::element::lagrange_variant::gll_warped,
                                            basix::element::dpc_variant::unset, false);
    auto V_DG = std::make_shared<fem::FunctionSpace<T>>(
        fem::create_functionspace(mesh, basix::FiniteElement<T>(0)));
    auto c0 = std::make_shared<fem::Function<T>>(V_DG);
    auto rho0 = std::make_shared<fem::Function<T>>(V_DG);
    auto delta0 = std::make_shared<fem::Function<T>>(V_DG);
    auto beta0 = std::make_shared<fem::Function<T>>(V_DG);
    std::span<T> c0_ = c0->x()->mutable_array();
    std::fill(c0_.begin(), c0_.end(), speedOfSound);
    std::span<T> rho0_ = rho0->x()->mutable_array();
    std::fill(rho0_.begin(), rho0_.end(), density);
    const T w0 = 2 * M_PI * sourceFrequency;
    std::span<T> d_ = delta0->x()->mutable_array();
    std::fill(d_.begin(), d_.end(), 2 * 5.0 * speedOfSound * speedOfSound * speedOfSound / w0 / w0);
    std::span<T> b_ = beta0->x()->mutable_array();
    std::fill(b_.begin(), b_.end(), 3.5);

    // Temporal parameters: the minimum cell size over all ranks (BM7-SC1/main.cpp:72-78, :112-118)
    const int tdim = mesh->topology()->dim();
    const int ncl = mesh->topology()->index_map(tdim)->size_local();
    std::vector<int> cells(ncl);
    std::iota(cells.begin(), cells.end(), 0);
    std::vector<T> hc = mesh::h(*mesh, cells, tdim);
    const T hmin_local = *std::min_element(hc.begin(), hc.end());
    const T meshSizeMinGlobal = comm.allreduce(hmin_local, 0); // MPI_Allreduce(..., MPI_MIN, ...)
    const T CFL = kind == "linear" ? 0.65 : 0.2; // all-facet absorbing term: see DESIGN.md section 6
    T timeStepSize = CFL * meshSizeMinGlobal / (speedOfSound * degreeOfBasis * degreeOfBasis);
    const int stepPerPeriod = period / timeStepSize + 1;
    timeStepSize = period / stepPerPeriod;
    const T startTime = 0.0, finalTime = startTime + (nsteps - 0.5) * timeStepSize;

    std::shared_ptr<fus::detail::SpectralModel3D<T, degreeOfBasis>> model;
    if (kind == "linear")
      model = std::make_shared<LinearSpectral3D<T, degreeOfBasis>>(
          element, mesh, mt_facet, c0, rho0, sourceFrequency, sourceAmplitude, speedOfSound);
    else if (kind == "lossy")
      model = std::make_shared<LossySpectral3D<T, degreeOfBasis>>(
          element, mesh, mt_facet, c0, rho0, delta0, sourceFrequency, sourceAmplitude, speedOfSound);
    else
      model = std::make_shared<WesterveltSpectral3D<T, degreeOfBasis>>(
          element, mesh, mt_facet, c0, rho0, delta0, beta0, sourceFrequency, sourceAmplitude,
          speedOfSound);
    if (comm.rank() == 0) {
      std::printf("Ranks: %d\nDegrees of freedom: %lld\n", comm.size(), (long long)model->number_of_dofs());
      std::printf("Time step size: %.17g\n", timeStepSize);
    }
    model->init();
    comm.barrier();
    auto t0 = std::chrono::steady_clock::now();
    model->rk4(startTime, finalTime, timeStepSize);
    comm.barrier();
    const double el = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

    // owned part of the solution: norms through the host collectives, values to disk
    auto u = model->u_sol()->x()->array();
    auto im = model->u_sol()->function_space()->dofmap()->index_map;
    const std::int32_t nowned = im->size_local();
    double s2 = 0.0;
    for (std::int32_t i = 0; i < nowned; ++i)
      s2 += u[i] * u[i];
    const double u_l2 = std::sqrt(comm.allreduce(s2, 2));
    if (comm.rank() == 0) {
      std::printf("Number of steps: %d\n", model->number_of_steps());
      std::printf("Solve time: %g\nTime per step: %g\n", el, el / model->number_of_steps());
      std::printf("u_l2: %.17g\n", u_l2);
    }
    if (!out.empty()) {
      std::vector<std::int32_t> loc(nowned);
      std::iota(loc.begin(), loc.end(), 0);
      std::vector<std::int64_t> glob(nowned);
      im->local_to_global(loc, glob);
      const std::string fn = out + "." + std::to_string(comm.rank()) + ".bin";
      std::FILE* f = std::fopen(fn.c_str(), "wb");
      if (!f)
        throw std::runtime_error("cannot write " + fn);
      for (std::int32_t i = 0; i < nowned; ++i) {
        std::fwrite(&glob[i], sizeof(std::int64_t), 1, f);
        std::fwrite(&u[i], sizeof(double), 1, f);
      }
      std::fclose(f);
    }
    // ---- end of the rank body ------------------------------------------------------------------
  });
  return 0;
}
