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
                      [&] { kern(x, x2, y, dofmap, gptr, coeff, coeff2, cb, ce, D, HaloLaunch{}); });
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
                    [&] { kern(x, x2, y, dofmap, gptr, coeff, coeff2, cb, ce, D, HaloLaunch{}); });
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
                      [&] { kern(x, x2, y, dofmap, gptr, coeff, coeff2, cb, ce, D, HaloLaunch{}); });
    };
    fuse ? run_ring(stiffness_line_kernel<N, true, 7>) : run_ring(stiffness_line_kernel<N, false, 7>);
  }
  return 0;
}

// the HALO instantiations of the line kernel (fused peer transport): one launch over all cells
template <int N>
int stiffness_halo_n(int geom, const double* x, const double* x2, double* y, const int32_t* dofmap,
                     const double* G, const double* coeff, const double* coeff2, long long ncells,
                     const double* dphi, const double* pts, const double* wts, int max_blocks,
                     FusedHalo* H, long long ninterface) {
  using L = LineCfg<N>;
  const DMat<N> D = make_dmat<N>(dphi, pts, wts);
  const std::vector<double2> G2 = device_layout<N>(G, ncells);
  // as assemble_rhs (fus_capi.cu) issues a fused stage: the interface cells [0, ninterface) with
  // the HALO kernel (or its bookkeeping alone when there are none), then the rest with the plain one
  if (ninterface == 0) {
    fus_emu::launch(1, 32, 0, [&] { halo_operator_skipped_kernel(H); });
  }
  auto blocks_of = [&](long long nc) {
    return (unsigned)std::max<long long>(1, std::min<long long>((nc + L::CPB - 1) / L::CPB, max_blocks));
  };
  HaloLaunch HL; // as halo_fused_launch (fus_halo.cu) builds it
  HL.H = H, HL.nown = H->nowned, HL.mbu = H->fwd_u - H->nowned, HL.mbv = H->fwd_v - H->nowned;
  auto run = [&](auto kern) {
    if (ninterface > 0)
      fus_emu::launch(blocks_of(ninterface), L::THREADS, L::SMEM_BYTES, [&] {
        kern(x, x2, y, dofmap, G2.data(), coeff, coeff2, 0, ninterface, D, HL);
      });
  };
  const bool fuse = x2 != nullptr;
  if (geom == 0)
    fuse ? run(stiffness_line_kernel<N, true, 0, double, true>)
         : run(stiffness_line_kernel<N, false, 0, double, true>);
  else if (geom == 4)
    fuse ? run(stiffness_line_kernel<N, true, 4, double, true>)
         : run(stiffness_line_kernel<N, false, 4, double, true>);
  else if (geom == 6)
    fuse ? run(stiffness_line_kernel<N, true, 6, double, true>)
         : run(stiffness_line_kernel<N, false, 6, double, true>);
  else
    return -1;
  if (ncells > ninterface) {
    auto rest = [&](auto kern) {
      fus_emu::launch(blocks_of(ncells - ninterface), L::THREADS, L::SMEM_BYTES, [&] {
        kern(x, x2, y, dofmap, G2.data(), coeff, coeff2, ninterface, ncells, D, HaloLaunch{});
      });
    };
    fuse ? rest(stiffness_line_kernel<N, true, 0>) : rest(stiffness_line_kernel<N, false, 0>);
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
                                              nullptr, 0, ncells, D, HaloLaunch{});
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
                  const long long* bchunk, double g_next, double dg_next, int hints, int grid,
                  void* fused) {
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
  A.halo = static_cast<FusedHalo*>(fused);
  A.halo_defer_wait = 1; // see emu_fused_forward_landed
  auto go = [&](auto kern) { fus_emu::launch((unsigned)grid, kStageThreads, 0, [&] { kern(A); }); };
  if (fused) {
    switch (stage * 2 + (westervelt ? 1 : 0)) {
    case 0: go(rk4_stage_kernel<0, false, false, true>); break;
    case 1: go(rk4_stage_kernel<0, true, false, true>); break;
    case 2: go(rk4_stage_kernel<1, false, false, true>); break;
    case 3: go(rk4_stage_kernel<1, true, false, true>); break;
    case 4: go(rk4_stage_kernel<2, false, false, true>); break;
    case 5: go(rk4_stage_kernel<2, true, false, true>); break;
    case 6: go(rk4_stage_kernel<3, false, false, true>); break;
    case 7: go(rk4_stage_kernel<3, true, false, true>); break;
    }
    return kStageChunk;
  }
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

// ---- fused peer transport (fus_halo_kernels.cuh): one emulated rank's state ------------------------
struct EmuFused {
  FusedHalo H;
  std::vector<int64_t> soff, roff;
  std::vector<int32_t> sidx, spos_off, spos;
  std::vector<signed char> spos_nb;
  unsigned long long seq[SEQ_COUNT];
  unsigned int ctr[CTR_COUNT];
  int error;
};

// mailbox: this rank's own, laid out by fus_halo_mailbox_layout (layout6).  Returns NULL when the
// numbering does not have the shape the fused kernels need.
void* emu_fused_create(long long nowned, int nneigh, const int64_t* send_off, const int32_t* sidx,
                       const int64_t* recv_off, const int32_t* ridx, char* mailbox,
                       const int64_t* layout6, double timeout_s) {
  EmuFused* e = new EmuFused();
  const int64_t nsend = nneigh ? send_off[nneigh] : 0, nrecv = nneigh ? recv_off[nneigh] : 0;
  e->soff.assign(send_off, send_off + nneigh + 1);
  e->roff.assign(recv_off, recv_off + nneigh + 1);
  e->sidx.assign(sidx, sidx + nsend);
  int64_t nshared = 0;
  const std::string why = fused_halo_lists(nowned, nneigh, send_off, sidx, ridx, nrecv, &nshared,
                                           e->spos_off, e->spos, e->spos_nb);
  if (!why.empty()) {
    delete e;
    return nullptr;
  }
  std::memset(e->seq, 0, sizeof(e->seq));
  std::memset(e->ctr, 0, sizeof(e->ctr));
  e->error = 0;
  FusedHalo& F = e->H;
  std::memset(&F, 0, sizeof(F));
  F.nneigh = nneigh, F.nowned = nowned, F.nghost = nrecv, F.nshared = nshared, F.nsend = nsend;
  F.fwd_u = (const double*)mailbox;
  F.fwd_v = (const double*)(mailbox + layout6[0]);
  F.rev = (const double*)(mailbox + layout6[1]);
  F.fwd_flag = (const unsigned long long*)(mailbox + layout6[2]);
  F.rev_flag = (const unsigned long long*)(mailbox + layout6[3]);
  F.ready_flag = (const unsigned long long*)(mailbox + layout6[4]);
  F.soff = e->soff.data(), F.roff = e->roff.data(), F.sidx = e->sidx.data();
  F.spos_off = e->spos_off.data(), F.spos = e->spos.data(), F.spos_nb = e->spos_nb.data();
  F.seq = e->seq, F.ctr = e->ctr, F.error = &e->error;
  F.timeout_ns = (unsigned long long)(timeout_s * 1e9);
  return e;
}
// neighbour k's mailbox and the byte offsets of this rank's runs inside it (fus_halo_peer_offsets)
int emu_fused_connect(void* h, int k, char* base, const int64_t* off6) {
  FusedHalo& F = static_cast<EmuFused*>(h)->H;
  F.r_fwd_u[k] = (double*)(base + off6[0]);
  F.r_fwd_v[k] = (double*)(base + off6[1]);
  F.r_rev[k] = (double*)(base + off6[2]);
  F.r_fwd_flag[k] = (unsigned long long*)(base + off6[3]);
  F.r_rev_flag[k] = (unsigned long long*)(base + off6[4]);
  F.r_ready_flag[k] = (unsigned long long*)(base + off6[5]);
  return 0;
}
long long emu_fused_nshared(void* h) { return static_cast<EmuFused*>(h)->H.nshared; }
int emu_fused_state(void* h, unsigned long long* seq, unsigned int* ctr) {
  EmuFused* e = static_cast<EmuFused*>(h);
  std::memcpy(seq, e->seq, sizeof(e->seq));
  std::memcpy(ctr, e->ctr, sizeof(e->ctr));
  return e->error;
}
void emu_fused_destroy(void* h) { delete static_cast<EmuFused*>(h); }
int emu_fused_ready(void* h, int phase) {
  FusedHalo* H = &static_cast<EmuFused*>(h)->H;
  fus_emu::launch(1, 32, 0, [&] { halo_ready_kernel(H, phase); });
  return 0;
}
// the closing wait of the entry put and of the epilogues runs as a kernel of its own here: ranks
// are emulated one after another, so a rank cannot wait for a neighbour that has not run yet
int emu_fused_entry_put(void* h, const double* u, const double* v) {
  FusedHalo* H = &static_cast<EmuFused*>(h)->H;
  fus_emu::launch((unsigned)std::max<long long>(1, (H->nsend + 255) / 256), 256, 0,
                  [&] { halo_entry_put_kernel(H, u, v, 1); });
  return 0;
}
int emu_fused_forward_landed(void* h) {
  FusedHalo* H = &static_cast<EmuFused*>(h)->H;
  fus_emu::launch(1, 32, 0, [&] { halo_forward_landed_kernel(H); });
  return 0;
}
int emu_fused_exit(void* h, double* u, double* v) {
  FusedHalo* H = &static_cast<EmuFused*>(h)->H;
  fus_emu::launch((unsigned)std::max<long long>(1, (H->nghost + 255) / 256), 256, 0,
                  [&] { halo_exit_unpack_kernel(H, u, v); });
  return 0;
}
// y += K x over all local cells as a fused stage issues it: the interface cells with the HALO line
// kernel (geom 0, 4 or 6; ghost values from the mailbox, ghost partial sums of y shipped to the
// owners), the rest with the plain kernel.
int emu_stiffness_fused(int N, int geom, const double* x, const double* x2, double* y,
                        const int32_t* dofmap, const double* G, const double* coeff,
                        const double* coeff2, long long ncells, const double* dphi,
                        const double* pts, const double* wts, int max_blocks, void* h,
                        long long ninterface) {
  FusedHalo* H = &static_cast<EmuFused*>(h)->H;
  EMU_DISPATCH(N, stiffness_halo_n, geom, x, x2, y, dofmap, G, coeff, coeff2, ncells, dphi, pts, wts,
               max_blocks, H, ninterface);
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

} // extern "C"
