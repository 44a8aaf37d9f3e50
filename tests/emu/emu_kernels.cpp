// emu_kernels.cpp -- TEST INFRASTRUCTURE: the kernels of fus_kernels.cuh compiled for the host
// through tests/emu/simt_emu.hpp, behind a small C interface for tests/test_kernel_emulation.py.
// Launch geometry follows fus_capi.cu (same block sizes, shared-memory sizes and cells per block);
// the grid is capped low so that the grid-stride loops and their tails are exercised.
#include "simt_emu.hpp"

#include "fus_halo_kernels.cuh"
#include "fus_kernels.cuh"

using namespace fus;

namespace {
template <int N>
DMat<N> make_dmat(const double* dphi, const double* pts, const double* wts) {
  DMat<N> D;
  std::memcpy(D.d, dphi, sizeof(double) * N * N);
  std::memcpy(D.w, wts, sizeof(double) * N);
  std::memcpy(D.x, pts, sizeof(double) * N);
  return D;
}

// reference layout G[c][q][6] -> device layout G2[cell][i0][p][t] (g_to_device_layout_kernel)
template <int N>
std::vector<double2> device_layout(const double* G, long long ncells) {
  constexpr int NN = N * N, Nd = N * NN;
  std::vector<double2> out((size_t)ncells * 3 * Nd);
  gridDim.x = 1;
  blockDim.x = 1;
  fus_emu::launch(1, 1, 0, [&] { g_to_device_layout_kernel<N>(G, ncells, out.data()); });
  return out;
}

template <int N>
int stiffness_n(int variant, int geom, const double* x, const double* x2, double* y,
                const int32_t* dofmap, const double* G, const double* aux, const double* coeff,
                const double* coeff2, long long ncells, const double* dphi, const double* pts,
                const double* wts, int max_blocks, long long cb, long long ce) {
  const DMat<N> D = make_dmat<N>(dphi, pts, wts);
  const bool fuse = x2 != nullptr;
  if (variant >= 3) { // stiffness_variant 3..6: line kernel, streamed G, kernel GEOM 4..7
    geom = variant + 1;
    variant = 2;
    if (geom == 7 && !LineCfg<N>::RING_FITS)
      geom = 6; // as fus_capi.cu: the ring of the largest degree does not fit in shared memory
  }
  std::vector<double2> G2;
  const double2* gptr = nullptr;
  if (geom == 0 || geom >= 4) {
    G2 = device_layout<N>(G, ncells);
    gptr = G2.data();
  } else {
    gptr = reinterpret_cast<const double2*>(aux); // Ghat[cell][3] double2, or the cell-map coefficients
  }
  auto blocks_for = [&](int cpb) {
    return (unsigned)std::max<long long>(1, std::min<long long>((ce - cb + cpb - 1) / cpb, max_blocks));
  };
  if (variant == 1) {
    auto run = [&](auto kern) {
      fus_emu::launch(blocks_for(1), N * N * N, 0, [&] { kern(x, x2, y, dofmap, gptr, coeff, coeff2, cb, ce, D); });
    };
    if (fuse)
      run(stiffness_point_kernel<N, true>);
    else
      run(stiffness_point_kernel<N, false>);
    return 0;
  }
  if (variant == 0) {
    using C = ColCfg<N>;
    auto run = [&](auto kern) {
      fus_emu::launch(blocks_for(C::CPB), C::THREADS, C::SMEM_BYTES,
                      [&] { kern(x, x2, y, dofmap, gptr, coeff, coeff2, cb, ce, D); });
    };
    if (fuse)
      run(stiffness_col_kernel<N, true>);
    else
      run(stiffness_col_kernel<N, false>);
    return 0;
  }
  using L = LineCfg<N>;
  auto run = [&](auto kern) {
    fus_emu::launch(blocks_for(L::CPB), L::THREADS, L::SMEM_BYTES,
                    [&] { kern(x, x2, y, dofmap, gptr, coeff, coeff2, cb, ce, D); });
  };
  if (geom == 0)
    fuse ? run(stiffness_line_kernel<N, true, 0>) : run(stiffness_line_kernel<N, false, 0>);
  else if (geom == 1)
    fuse ? run(stiffness_line_kernel<N, true, 1>) : run(stiffness_line_kernel<N, false, 1>);
  else if (geom == 2)
    fuse ? run(stiffness_line_kernel<N, true, 2>) : run(stiffness_line_kernel<N, false, 2>);
  else if (geom == 3)
    fuse ? run(stiffness_line_kernel<N, true, 3>) : run(stiffness_line_kernel<N, false, 3>);
  else if (geom == 4)
    fuse ? run(stiffness_line_kernel<N, true, 4>) : run(stiffness_line_kernel<N, false, 4>);
  else if (geom == 5)
    fuse ? run(stiffness_line_kernel<N, true, 5>) : run(stiffness_line_kernel<N, false, 5>);
  else if (geom == 6)
    fuse ? run(stiffness_line_kernel<N, true, 6>) : run(stiffness_line_kernel<N, false, 6>);
  else if constexpr (LineCfg<N>::RING_FITS) {
    auto run_ring = [&](auto kern) {
      fus_emu::launch(blocks_for(L::CPB), L::THREADS, L::SMEM_BYTES_RING,
                      [&] { kern(x, x2, y, dofmap, gptr, coeff, coeff2, cb, ce, D); });
    };
    fuse ? run_ring(stiffness_line_kernel<N, true, 7>) : run_ring(stiffness_line_kernel<N, false, 7>);
  }
  return 0;
}

template <int N>
int quad_n(const double* x, const double* x2, double* y, const int32_t* dofmap, const double* Gq,
           const double* coeff, const double* coeff2, long long ncells, const double* dphi,
           const double* pts, const double* wts, int max_blocks) {
  using Q = QuadCfg<N>;
  const DMat<N> D = make_dmat<N>(dphi, pts, wts);
  const unsigned blocks
      = (unsigned)std::max<long long>(1, std::min<long long>((ncells + Q::CPB - 1) / Q::CPB, max_blocks));
  if (x2)
    fus_emu::launch(blocks, Q::THREADS, 0, [&] {
      stiffness_quad_kernel<N, true>(x, x2, y, dofmap, Gq, coeff, coeff2, 0, ncells, D);
    });
  else
    fus_emu::launch(blocks, Q::THREADS, 0, [&] {
      stiffness_quad_kernel<N, false>(x, x2, y, dofmap, Gq, coeff, coeff2, 0, ncells, D);
    });
  return 0;
}

template <int N>
int tri_n(const double* xg, const int32_t* xdofmap, long long ncells, double* coeffs,
          const double* x, double* y, const int32_t* dofmap, const double* coeff,
          const double* pts, const double* wts) {
  fus_emu::launch((unsigned)((ncells + 127) / 128), 128, 0,
                  [&] { tri_coeff_kernel(xg, xdofmap, ncells, coeffs); });
  Rule1D<N> R;
  std::memcpy(R.pts, pts, sizeof(double) * N);
  std::memcpy(R.wts, wts, sizeof(double) * N);
  const long long np = ncells * N * N * N;
  fus_emu::launch(2, 256, 0, [&] { mass_tri_kernel<N>(x, y, dofmap, coeffs, coeff, 0, np, R); });
  return 0;
}

// FP32 instantiation: float copies of G (narrow_kernel on the device-layout array, as fus_capi.cu
// does), then stiffness_line_kernel<N,false,0,float> and mass_kernel_f32
template <int N>
int f32_n(const float* x, float* y, float* ym, const int32_t* dofmap, const double* G,
          const double* detJ, const float* coeff, long long ncells, const double* dphi,
          const double* pts, const double* wts, int max_blocks) {
  using L = LineCfg<N>;
  constexpr int Nd = N * N * N;
  const std::vector<double2> G2 = device_layout<N>(G, ncells);
  std::vector<float> G2f((size_t)ncells * Nd * 6), dJf((size_t)ncells * Nd);
  fus_emu::launch(2, 256, 0, [&] {
    narrow_kernel(reinterpret_cast<const double*>(G2.data()), G2f.data(), (long long)G2f.size());
  });
  fus_emu::launch(2, 256, 0, [&] { narrow_kernel(detJ, dJf.data(), (long long)dJf.size()); });
  DMatT<float, N> D;
  for (int i = 0; i < N * N; ++i)
    D.d[i] = (float)dphi[i];
  for (int i = 0; i < N; ++i) {
    D.w[i] = (float)wts[i];
    D.x[i] = (float)pts[i];
  }
  const unsigned blocks
      = (unsigned)std::max<long long>(1, std::min<long long>((ncells + L::CPB - 1) / L::CPB, max_blocks));
  fus_emu::launch(blocks, L::THREADS, L::SMEM_BYTES, [&] {
    stiffness_line_kernel<N, false, 0, float>(x, nullptr, y, dofmap,
                                              reinterpret_cast<const float2*>(G2f.data()), coeff,
                                              nullptr, 0, ncells, D);
  });
  fus_emu::launch(3, 256, 0, [&] {
    mass_kernel_f32(x, ym, dofmap, dJf.data(), coeff, ncells * Nd, Nd);
  });
  return 0;
}

template <int N>
int geometry_n(const double* xg, const int32_t* xdofmap, long long ncells, double* G, double* detJ,
               const double* pts, const double* wts, double* Ghat, int* all_affine) {
  constexpr int Nd = N * N * N;
  Rule1D<N> R;
  std::memcpy(R.pts, pts, sizeof(double) * N);
  std::memcpy(R.wts, wts, sizeof(double) * N);
  std::vector<double2> G2((size_t)ncells * 3 * Nd);
  fus_emu::launch(3, 128, 0, [&] { geometry_kernel<N>(xg, xdofmap, ncells, G2.data(), detJ, R); });
  fus_emu::launch(2, 256, 0, [&] { g_from_device_layout_kernel<N>(G2.data(), ncells, G); });
  DMat<N> D{};
  std::memcpy(D.w, wts, sizeof(double) * N);
  *all_affine = 1;
  fus_emu::launch((unsigned)((ncells + 127) / 128), 128, 0, [&] {
    affine_detect_kernel<N>(G2.data(), ncells, 1e-13, reinterpret_cast<double2*>(Ghat), all_affine, D);
  });
  return 0;
}

template <int N>
int geometry_quad_n(const double* xg, const int32_t* xdofmap, long long ncells, double* Gq,
                    double* detJ, const double* pts, const double* wts) {
  Rule1D<N> R;
  std::memcpy(R.pts, pts, sizeof(double) * N);
  std::memcpy(R.wts, wts, sizeof(double) * N);
  fus_emu::launch(2, 128, 0, [&] { geometry_quad_kernel<N>(xg, xdofmap, ncells, Gq, detJ, R); });
  return 0;
}

#define EMU_DISPATCH(N, fn, ...)                                                                   \
  switch (N) {                                                                                     \
  case 2: return fn<2>(__VA_ARGS__);                                                               \
  case 3: return fn<3>(__VA_ARGS__);                                                               \
  case 4: return fn<4>(__VA_ARGS__);                                                               \
  case 5: return fn<5>(__VA_ARGS__);                                                               \
  case 6: return fn<6>(__VA_ARGS__);                                                               \
  case 7: return fn<7>(__VA_ARGS__);                                                               \
  case 8: return fn<8>(__VA_ARGS__);                                                               \
  }                                                                                                \
  return -1
} // namespace

extern "C" {

// y += K x with the production kernels: variant 0 column, 1 point, 2 line, 3/4/5 line with the
// experimental software pipelines (as option stiffness_variant); geom 0 streamed G
// (reference layout in, re-laid-out here), 1 affine (aux = Ghat[cell][6]), 2 trilinear (aux =
// FUS_TRI_STRIDE doubles per cell).  x2/coeff2 non-NULL selects the fused two-vector gather.
// Only cells [cell_begin, cell_end) are applied: the split launches of a partitioned stage.
int emu_stiffness(int N, int variant, int geom, const double* x, const double* x2, double* y,
                  const int32_t* dofmap, const double* G, const double* aux, const double* coeff,
                  const double* coeff2, long long ncells, const double* dphi, const double* pts,
                  const double* wts, int max_blocks, long long cell_begin, long long cell_end) {
  EMU_DISPATCH(N, stiffness_n, variant, geom, x, x2, y, dofmap, G, aux, coeff, coeff2, ncells, dphi,
               pts, wts, max_blocks, cell_begin, cell_end);
}

int emu_stiffness_quad(int N, const double* x, const double* x2, double* y, const int32_t* dofmap,
                       const double* Gq, const double* coeff, const double* coeff2, long long ncells,
                       const double* dphi, const double* pts, const double* wts, int max_blocks) {
  EMU_DISPATCH(N, quad_n, x, x2, y, dofmap, Gq, coeff, coeff2, ncells, dphi, pts, wts, max_blocks);
}

// tri_coeff_kernel (coeffs out) followed by mass_tri_kernel (y += M x)
int emu_tri_coeffs_and_mass(int N, const double* xg, const int32_t* xdofmap, long long ncells,
                            double* coeffs, const double* x, double* y, const int32_t* dofmap,
                            const double* coeff, const double* pts, const double* wts) {
  EMU_DISPATCH(N, tri_n, xg, xdofmap, ncells, coeffs, x, y, dofmap, coeff, pts, wts);
}

// geometry_kernel -> G2 -> g_from_device_layout_kernel (reference layout out), then
// affine_detect_kernel on the same G2 (Ghat[ncells][6] and the all-affine flag out)
int emu_geometry(int N, const double* xg, const int32_t* xdofmap, long long ncells, double* G,
                 double* detJ, const double* pts, const double* wts, double* Ghat, int* all_affine) {
  EMU_DISPATCH(N, geometry_n, xg, xdofmap, ncells, G, detJ, pts, wts, Ghat, all_affine);
}

int emu_geometry_quad(int N, const double* xg, const int32_t* xdofmap, long long ncells, double* Gq,
                      double* detJ, const double* pts, const double* wts) {
  EMU_DISPATCH(N, geometry_quad_n, xg, xdofmap, ncells, Gq, detJ, pts, wts);
}

// y += K x and ym += M x in FP32 (float data in and out; G / detJ given in FP64 reference layouts)
int emu_operators_f32(int N, const float* x, float* y, float* ym, const int32_t* dofmap,
                      const double* G, const double* detJ, const float* coeff, long long ncells,
                      const double* dphi, const double* pts, const double* wts, int max_blocks) {
  EMU_DISPATCH(N, f32_n, x, y, ym, dofmap, G, detJ, coeff, ncells, dphi, pts, wts, max_blocks);
}

int emu_mass(const double* x, double* y, const int32_t* dofmap, const double* detJ,
             const double* coeff, long long npoints, int Nd) {
  fus_emu::launch(3, 256, 0, [&] { mass_kernel(x, y, dofmap, detJ, coeff, npoints, Nd); });
  return 0;
}

// one fused RK4 stage epilogue; vectors are updated in place.  The boundary terms of the NEXT stage
// are seeded into b from the compacted list (nb entries, bchunk = first entry per chunk of
// kStageChunk dofs) with the source scalars (g_next, dg_next); nb == 0 zero-fills b.
int emu_rk4_stage(int stage, int westervelt, double* b, const double* m, const double* dnl,
                  double* u0, double* v0, double* ua, double* va, double* un, double* vn,
                  long long nowned, long long ntotal, double dt, long long nb, const int32_t* bidx,
                  const double* bsrc, const double* bdsrc, const double* babs,
                  const long long* bchunk, double g_next, double dg_next, int hints, int grid) {
  StageArgs A;
  A.b = b, A.m = m, A.dnl = dnl, A.u0 = u0, A.v0 = v0, A.ua = ua, A.va = va, A.un = un, A.vn = vn;
  A.nowned = nowned, A.ntotal = ntotal;
  stage_coefficients(A, stage, dt);
  int step_ctr = 0;
  unsigned done = 0;
  double table[10];
  for (int r = 0; r < 5; ++r)
    table[2 * r] = g_next, table[2 * r + 1] = dg_next;
  A.step_ctr = &step_ctr, A.done_ctr = &done;
  A.nb = nb, A.bidx = bidx, A.bsrc = bsrc, A.bdsrc = bdsrc, A.babs = babs, A.bchunk = bchunk;
  A.src_table = table;
  auto go = [&](auto kern) { fus_emu::launch((unsigned)grid, kStageThreads, 0, [&] { kern(A); }); };
  switch (stage * 4 + (westervelt ? 2 : 0) + (hints ? 1 : 0)) {
  case 0: go(rk4_stage_kernel<0, false, false>); break;
  case 1: go(rk4_stage_kernel<0, false, true>); break;
  case 2: go(rk4_stage_kernel<0, true, false>); break;
  case 3: go(rk4_stage_kernel<0, true, true>); break;
  case 4: go(rk4_stage_kernel<1, false, false>); break;
  case 5: go(rk4_stage_kernel<1, false, true>); break;
  case 6: go(rk4_stage_kernel<1, true, false>); break;
  case 7: go(rk4_stage_kernel<1, true, true>); break;
  case 8: go(rk4_stage_kernel<2, false, false>); break;
  case 9: go(rk4_stage_kernel<2, false, true>); break;
  case 10: go(rk4_stage_kernel<2, true, false>); break;
  case 11: go(rk4_stage_kernel<2, true, true>); break;
  case 12: go(rk4_stage_kernel<3, false, false>); break;
  case 13: go(rk4_stage_kernel<3, false, true>); break;
  case 14: go(rk4_stage_kernel<3, true, false>); break;
  case 15: go(rk4_stage_kernel<3, true, true>); break;
  default: return -1;
  }
  if (stage == 3 && (step_ctr != 1 || done != 0))
    return -2; // the last block advances the step counter exactly once
  return kStageChunk;
}

int emu_boundary(double* b, const double* v, const int32_t* bidx, const double* bsrc,
                 const double* bdsrc, const double* babs, long long nb, double g, double dg) {
  fus_emu::launch((unsigned)((nb + 255) / 256), 256, 0, [&] {
    boundary_kernel(b, v, bidx, bsrc, bdsrc, babs, nb, g, dg, nullptr, nullptr, 0);
  });
  return 0;
}

// ---- halo kernels (fus_halo_kernels.cuh) ----------------------------------------------------------
// NCCL transport: pack the interface values of one or two vectors / unpack (insert or add)
int emu_halo_pack(const double* a, const double* b, const int32_t* idx, const int64_t* off,
                  int nneigh, double* buf, long long n, int nv) {
  fus_emu::launch((unsigned)((n + 255) / 256), 256, 0,
                  [&] { halo_pack_kernel(a, b, idx, off, nneigh, buf, n, nv); });
  return 0;
}
int emu_halo_unpack(int add, double* a, double* b, const int32_t* idx, const int64_t* off,
                    int nneigh, const double* buf, long long n, int nv) {
  fus_emu::launch((unsigned)((n + 255) / 256), 256, 0, [&] {
    if (add)
      halo_unpack_kernel<true>(a, b, idx, off, nneigh, buf, n, nv);
    else
      halo_unpack_kernel<false>(a, b, idx, off, nneigh, buf, n, nv);
  });
  return 0;
}

// peer-direct transport: one rank's put into its neighbours' mailboxes (dst[4*k..] = addresses of
// {forward data, forward flag, reverse data, reverse flag} inside neighbour k's mailbox), and one
// rank's wait + unpack from its own mailbox
int emu_peer_put(const double* a, const double* b, const int32_t* idx, const int64_t* off,
                 int nneigh, long long n, int nv, const unsigned long long* dst, int forward,
                 unsigned* counter, unsigned long long* epoch_ctr, int lightfence) {
  PeerTable tab;
  std::memset(&tab, 0, sizeof(tab));
  for (int k = 0; k < nneigh; ++k) {
    tab.fwd_dst[k] = reinterpret_cast<double*>(dst[4 * k + 0]);
    tab.fwd_flag[k] = reinterpret_cast<unsigned long long*>(dst[4 * k + 1]);
    tab.rev_dst[k] = reinterpret_cast<double*>(dst[4 * k + 2]);
    tab.rev_flag[k] = reinterpret_cast<unsigned long long*>(dst[4 * k + 3]);
  }
  fus_emu::launch((unsigned)std::max<long long>(1, (n + 255) / 256), 256, 0, [&] {
    peer_put_kernel(a, b, idx, off, nneigh, n, nv, &tab, forward, counter, epoch_ctr, lightfence);
  });
  return 0;
}
int emu_peer_wait(int add, double* a, double* b, const int32_t* idx, const int64_t* off, int nneigh,
                  long long n, int nv, const double* mbox_data, const unsigned long long* flags,
                  const unsigned long long* epoch_ctr, int* error) {
  fus_emu::launch((unsigned)((n + 255) / 256), 256, 0, [&] {
    if (add)
      peer_wait_kernel<true>(a, b, idx, off, nneigh, n, nv, mbox_data, flags, epoch_ctr, error);
    else
      peer_wait_kernel<false>(a, b, idx, off, nneigh, n, nv, mbox_data, flags, epoch_ctr, error);
  });
  return 0;
}

} // extern "C"
