// fus/spectral_op.hpp -- drop-in for cpp/fenicsx-sf/common/spectral_op.hpp of the reference.
//
// Same class names, template parameters and call signatures:
//   MassSpectral3D<T,P>(std::shared_ptr<fem::FunctionSpace<T>>& V)            (spectral_op.hpp:29-32)
//   StiffnessSpectral3D<T,P>(std::shared_ptr<fem::FunctionSpace<T>>& V)       (spectral_op.hpp:132-135)
//   void operator()(const la::Vector<T,Alloc>& x, std::span<T> coeffs, la::Vector<T,Alloc>& y)
//                                                                             (:69-70, :173-174)
// with the same semantics: y += A(coeffs) x over the local cells, no communication, the caller
// zero-fills y and refreshes the ghosts of x.  The arithmetic runs in the CUDA library through the
// C ABI (fus_b200.h).  T = double runs the FP64 kernels; T = float (the reference's test_operators3d
// and its float timing runs instantiate the operators with it) runs the FP32 instantiation of the
// hexahedral operators (fus_*_apply_f32_host: float copies of the cell data on the device), and is
// widened to FP64 only where no FP32 kernel exists (2-D, lean contexts).  On a partitioned mesh
// (mesh::create_box(comm, ...)) the context also sets up the halo exchange between the GPUs, and
// la::Vector::scatter_fwd / scatter_rev of host vectors go through it.
// Construction uploads the cell data once (the reference
// precomputes G / detJ in its constructor too); each call moves x, coeffs and y across PCIe --
// the solver classes (fus/Linear.hpp, ...) keep everything resident instead.
#pragma once

#include "dolfinx_shim.hpp"

namespace fus::detail {
/// One device context per function space, shared by the operators built on it.
template <typename T>
class SpaceContext {
public:
  explicit SpaceContext(const dolfinx::fem::FunctionSpace<T>& V, int device = -1) {
    static_assert(std::is_floating_point_v<T>, "real scalar types only");
    auto mesh = V.mesh();
    if (device < 0) // rank r of a partitioned run drives GPU r
      device = mesh->comm().size() > 1 ? mesh->comm().rank() : 0;
    auto dm = V.dofmap()->map();
    auto im = V.dofmap()->index_map;
    auto xd = mesh->geometry().dofmap();
    auto xT = mesh->geometry().x();
    const std::vector<double> x(xT.begin(), xT.end()); // the device path is FP64
    auto create = mesh->topology()->dim() == 2 ? fus_ctx_create_from_mesh_2d
                                               : fus_ctx_create_from_mesh;
    check(create(V.degree(), (std::int64_t)dm.extent(0), im->size_local() + im->num_ghosts(),
                 im->size_local(), dm.data_handle(), (std::int64_t)x.size() / 3, x.data(),
                 xd.data_handle(), device, &_ctx),
          "fus_ctx_create_from_mesh");
    if (mesh->comm().size() > 1)
      connect(V, device);
  }
  ~SpaceContext() { fus_ctx_destroy(_ctx); }
  SpaceContext(const SpaceContext&) = delete;
  SpaceContext& operator=(const SpaceContext&) = delete;
  fus_ctx* get() const { return _ctx; }

private:
  /// Partitioned mesh: what DOLFINx sets up inside the IndexMap's scatterer.  NCCL communicator
  /// (the unique id travels through the ranks' shared world), then the fused peer transport: every
  /// rank publishes its mailbox, the neighbours map it (the ranks are threads of one process, so a
  /// device pointer and peer access are all it takes).  Host vectors scatter through this context.
  void connect(const dolfinx::fem::FunctionSpace<T>& V, int device) {
    const fus::Comm& comm = V.mesh()->comm();
    auto halo = V.dofmap()->halo;
    if (!halo)
      throw std::runtime_error("SpaceContext: a partitioned mesh needs the halo lists of its dofmap");
    auto& W = comm.world();
    if (comm.rank() == 0)
      check(fus_comm_unique_id(W.nccl_id.data()), "fus_comm_unique_id");
    comm.barrier();
    check(fus_halo_setup(_ctx, comm.rank(), comm.size(), W.nccl_id.data(), (int)halo->neigh.size(),
                         halo->neigh.data(), halo->send_off.data(), halo->send_idx.data(),
                         halo->recv_off.data(), halo->recv_idx.data(), halo->ninterface_cells),
          "fus_halo_setup");
    auto& me = W.peers[comm.rank()];
    me.layout.assign(6, 0);
    me.device = device;
    me.neigh = halo->neigh;
    me.soff = halo->send_off;
    me.roff = halo->recv_off;
    check(fus_halo_peer_export(_ctx, nullptr, me.layout.data(), &me.mailbox), "fus_halo_peer_export");
    comm.barrier();
    const std::size_t nn = halo->neigh.size();
    std::vector<void*> bases(std::max<std::size_t>(nn, 1));
    std::vector<int> devs(std::max<std::size_t>(nn, 1));
    std::vector<std::int64_t> off(6 * std::max<std::size_t>(nn, 1));
    for (std::size_t k = 0; k < nn; ++k) {
      const auto& q = W.peers[halo->neigh[k]];
      const int j = (int)(std::find(q.neigh.begin(), q.neigh.end(), comm.rank()) - q.neigh.begin());
      check(fus_halo_peer_offsets(q.layout.data(), q.soff.data(), q.roff.data(), j, off.data() + 6 * k),
            "fus_halo_peer_offsets");
      bases[k] = q.mailbox;
      devs[k] = q.device;
    }
    // the context stays on NCCL if the fused transport cannot be used (FUS_ERR_UNSUPPORTED)
    const int rc = fus_halo_peer_connect_local(_ctx, bases.data(), devs.data(), off.data());
    if (rc != FUS_OK && rc != FUS_ERR_UNSUPPORTED)
      check(rc, "fus_halo_peer_connect_local");
    comm.barrier();
    fus_ctx* c = _ctx;
    const std::int64_t n = V.dofmap()->index_map->size_local() + V.dofmap()->index_map->num_ghosts();
    V.dofmap()->index_map->set_scatter([c, n](double* xh, bool forward) {
      void* d = nullptr;
      check(fus_dev_alloc(c, sizeof(double) * n, &d), "fus_dev_alloc");
      check(fus_dev_upload(c, d, xh, sizeof(double) * n), "fus_dev_upload");
      check(forward ? fus_scatter_fwd_dev(c, (double*)d) : fus_scatter_rev_dev(c, (double*)d),
            "fus_scatter_*_dev");
      check(fus_dev_download(c, xh, d, sizeof(double) * n), "fus_dev_download");
      check(fus_dev_free(c, d), "fus_dev_free");
    });
  }
  fus_ctx* _ctx = nullptr;
};

/// y += A(coeffs) x through a `*_apply_host` entry point: T = double runs the FP64 kernels, T = float
/// the FP32 instantiation (float copies of the cell data on the device, fus_*_apply_f32_host), as
/// the reference's float runs do (tests/test_operators3d/main.cpp:13).
template <typename T>
void apply_host(int (*entry)(fus_ctx*, const double*, const double*, double*),
                int (*entry_f32)(fus_ctx*, const float*, const float*, float*), const char* what,
                fus_ctx* ctx, std::span<const T> x, std::span<T> coeffs, std::span<T> y) {
  if constexpr (std::is_same_v<T, double>) {
    check(entry(ctx, x.data(), coeffs.data(), y.data()), what);
  } else if constexpr (std::is_same_v<T, float>) {
    const int rc = entry_f32(ctx, x.data(), coeffs.data(), y.data());
    if (rc != FUS_ERR_UNSUPPORTED) { // 2-D and lean contexts have no FP32 kernels: widen instead
      check(rc, what);
      return;
    }
    const std::vector<double> xd(x.begin(), x.end()), cd(coeffs.begin(), coeffs.end());
    std::vector<double> yd(y.begin(), y.end());
    check(entry(ctx, xd.data(), cd.data(), yd.data()), what);
    std::transform(yd.begin(), yd.end(), y.begin(), [](double v) { return (T)v; });
  } else {
    const std::vector<double> xd(x.begin(), x.end()), cd(coeffs.begin(), coeffs.end());
    std::vector<double> yd(y.begin(), y.end());
    check(entry(ctx, xd.data(), cd.data(), yd.data()), what);
    std::transform(yd.begin(), yd.end(), y.begin(), [](double v) { return (T)v; });
  }
}
} // namespace fus::detail

using namespace dolfinx;

/// 3D Spectral Mass operator (spectral_op.hpp:28-107)
template <typename T, int P>
class MassSpectral3D {
public:
  MassSpectral3D(std::shared_ptr<fem::FunctionSpace<T>>& V)
      : _ctx(std::make_shared<fus::detail::SpaceContext<T>>(*V)) {
    static_assert(P >= 1 && P <= 7, "supported degrees: 1..7");
    if (V->degree() != P)
      throw std::runtime_error("MassSpectral3D: function space degree != P");
  }

  /// Operator y += M x
  template <typename Alloc>
  void operator()(const la::Vector<T, Alloc>& x, std::span<T> coeffs, la::Vector<T, Alloc>& y) {
    fus::detail::apply_host<T>(fus_mass_apply_host, fus_mass_apply_f32_host, "fus_mass_apply_host",
                               _ctx->get(), x.array(), coeffs, y.mutable_array());
  }

private:
  std::shared_ptr<fus::detail::SpaceContext<T>> _ctx;
};

/// 3D Spectral Stiffness operator (spectral_op.hpp:132-284)
template <typename T, int P>
class StiffnessSpectral3D {
public:
  StiffnessSpectral3D(std::shared_ptr<fem::FunctionSpace<T>>& V)
      : _ctx(std::make_shared<fus::detail::SpaceContext<T>>(*V)) {
    static_assert(P >= 1 && P <= 7, "supported degrees: 1..7");
    if (V->degree() != P)
      throw std::runtime_error("StiffnessSpectral3D: function space degree != P");
  }

  /// Operator y += K x
  template <typename Alloc>
  void operator()(const la::Vector<T, Alloc>& x, std::span<T> coeffs, la::Vector<T, Alloc>& y) {
    fus::detail::apply_host<T>(fus_stiffness_apply_host, fus_stiffness_apply_f32_host,
                               "fus_stiffness_apply_host", _ctx->get(), x.array(), coeffs,
                               y.mutable_array());
  }

private:
  std::shared_ptr<fus::detail::SpaceContext<T>> _ctx;
};

/// 2D Spectral Mass operator (cpp/fenicsx-sf-naive/common/spectral_op.hpp:28-107): the same
/// device path on a quadrilateral function space
template <typename T, int P>
class MassSpectral2D : public MassSpectral3D<T, P> {
public:
  MassSpectral2D(std::shared_ptr<fem::FunctionSpace<T>>& V) : MassSpectral3D<T, P>(V) {
    if (V->mesh()->topology()->dim() != 2)
      throw std::runtime_error("MassSpectral2D: quadrilateral mesh expected");
  }
};

/// 2D Spectral Stiffness operator (cpp/fenicsx-sf-naive/common/spectral_op.hpp:226-359)
template <typename T, int P>
class StiffnessSpectral2D : public StiffnessSpectral3D<T, P> {
public:
  StiffnessSpectral2D(std::shared_ptr<fem::FunctionSpace<T>>& V) : StiffnessSpectral3D<T, P>(V) {
    if (V->mesh()->topology()->dim() != 2)
      throw std::runtime_error("StiffnessSpectral2D: quadrilateral mesh expected");
  }
};
