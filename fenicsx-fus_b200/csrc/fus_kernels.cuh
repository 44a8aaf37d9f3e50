// fus_kernels.cuh -- hand-written sm_100a kernels of the sum-factorised operator + RK4 path.
//
// FP64 on the CUDA cores (no tensor cores: the 1-D contractions are (P+1)x(P+1) and the path is
// HBM-bound).  References are to cpp/fenicsx-sf/common/ of adeebkor/fenicsx-fus.
//
// Device layout of the geometric factor (DESIGN.md "data layout"):
//   G2[cell][i0][p][t]  (double2),  p = 0..2 holding (G00,G01) (G02,G11) (G12,G22),
//   t = i1*N + i2, i.e. for one cell and one i0-level the N*N points of a component pair are
//   contiguous: a thread column (i1,i2) streams its 3*N double2 with 16-byte coalesced loads.
#pragma once
// FUS_HOST_EMULATION: defined only by tests/emu (a SIMT emulator that runs these kernels on host
// threads so that their indexing, staging and barrier logic is exercised by the CPU test suite).
// The guarded alternatives replace inline PTX and CUDA-only declarations; device builds never see
// them.
#include "fus_halo_kernels.cuh"
#include "fus_trilinear.hpp"

#ifndef FUS_HOST_EMULATION
#include <cuda_runtime.h>
#endif

#include <cstdint>

namespace fus {

// Scalar type of an operator instantiation: double everywhere in the solvers; float exists for
// the operator classes only (the reference's float runs).
template <typename T>
struct Vec2;
template <>
struct Vec2<double> {
  using type = double2;
};
template <>
struct Vec2<float> {
  using type = float2;
};
template <typename T>
__device__ __forceinline__ typename Vec2<T>::type make_v2(T a, T b) {
  typename Vec2<T>::type v;
  v.x = a;
  v.y = b;
  return v;
}

template <typename T, int N>
struct DMatT {
  T d[N * N]; // d[q*N+k] = phi_k'(xi_q)
  T w[N];     // 1-D GLL weights (only read by the compressed-geometry kernels)
  T x[N];     // 1-D GLL points  (only read by the trilinear-geometry kernel)
};
template <int N>
using DMat = DMatT<double, N>;

template <int N>
struct Rule1D {
  double pts[N];
  double wts[N];
};

// Streaming 16-byte load for data that is used exactly once (G): read-only path, do not keep
// the line in L1 so that L1 stays available for the gathered dofs.
__device__ __forceinline__ double2 ld_stream(const double2* p) {
#ifdef FUS_HOST_EMULATION
  return *p;
#else
  double2 v;
  asm("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
  return v;
#endif
}
__device__ __forceinline__ float2 ld_stream(const float2* p) {
#ifdef FUS_HOST_EMULATION
  return *p;
#else
  float2 v;
  asm("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
#endif
}

// ------------------------------------------------------------------------------------------------
// TMA bulk copies into a shared-memory ring (line kernel GEOM 7): one cp.async.bulk per cell brings
// the cell's whole block of G (48*N^3 contiguous bytes) into a ring stage and signals an mbarrier
// with the byte count; consumers wait on the barrier's phase.  No registers and no scoreboard are
// tied up while the copy is in flight.  Waits are bounded: a protocol error gives wrong numbers (which
// the parity checks see), never a hung device.
// ------------------------------------------------------------------------------------------------
#ifdef FUS_HOST_EMULATION
__device__ __forceinline__ void ring_bar_init(unsigned long long* bar) {
  std::atomic_ref<unsigned long long>(*bar).store(0ull, std::memory_order_release);
}
__device__ __forceinline__ void ring_bar_init_fence() {}
// the emulated copy is synchronous; the barrier word counts completed phases
__device__ __forceinline__ void ring_issue(void* dst, const void* src, unsigned bytes,
                                           unsigned long long* bar) {
  std::memcpy(dst, src, bytes);
  std::atomic_ref<unsigned long long>(*bar).fetch_add(1ull, std::memory_order_release);
}
__device__ __forceinline__ unsigned long long l2_policy_evict_first() { return 0ull; }
__device__ __forceinline__ void ring_issue_hint(void* dst, const void* src, unsigned bytes,
                                                unsigned long long* bar, unsigned long long) {
  ring_issue(dst, src, bytes, bar);
}
__device__ __forceinline__ bool ring_wait(unsigned long long* bar, unsigned phase_index) {
  for (long long spin = 0; spin < (1ll << 34); ++spin) {
    if (std::atomic_ref<unsigned long long>(*bar).load(std::memory_order_acquire) > phase_index)
      return true;
    std::this_thread::yield();
  }
  return false;
}
#else
__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void ring_bar_init(unsigned long long* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void ring_bar_init_fence() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// arm the stage's barrier with the byte count, then start the copy that will complete it; the
// proxy fence orders the generic-proxy reads of the previous contents before the async-proxy write
__device__ __forceinline__ void ring_issue(void* dst, const void* src, unsigned bytes,
                                           unsigned long long* bar) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// the same with an L2 eviction policy on the source lines (createpolicy): for a stream that is read
// once and must not push the vectors out of L2
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
  unsigned long long pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void ring_issue_hint(void* dst, const void* src, unsigned bytes,
                                                unsigned long long* bar, unsigned long long pol) {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
               "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
               : "memory");
}
__device__ __forceinline__ bool ring_wait(unsigned long long* bar, unsigned phase_index) {
  const unsigned addr = smem_u32(bar), parity = phase_index & 1u;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (int spin = 0;; ++spin) { // try_wait suspends for a while by itself
    unsigned done;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(addr), "r"(parity)
                 : "memory");
    if (done)
      return true;
    if ((spin & 15) == 15) { // a copy lands within microseconds: give up after 20 ms
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 20000000ull)
        return false;
    }
  }
}
#endif

// ------------------------------------------------------------------------------------------------
// Stiffness operator, "column" kernel (kernels (1)(2)(3) of the north star in one pass):
//   y[dof] += sum_cells B^T (coeff_c G_c) B x[dof]        StiffnessSpectral3D::operator(),
//                                                         spectral_op.hpp:173-243
// One thread owns the column (i1,i2) of a cell and keeps the N values along i0 in registers:
//   - gather  x[dofmap]                                   (:185-186)
//   - direction-0 contraction in registers with dphi as immediate constant-bank operands
//   - direction-1/2 contractions through shared memory    (:194-210)
//   - G transform with G streamed straight to registers   (:113-130, :213-214)
//   - transposed contractions                             (:221-238)
//   - scatter-add with FP64 RED atomics                   (:240-241)
// The geometric factors of the NEXT cell are loaded into the registers the current cell has just
// consumed (level by level), and the next cell's dofs are gathered as soon as the current ones
// have been staged, so a warp always has one whole cell (~7 KB at P=4) of HBM requests in flight.
// For N*N <= 32 a cell (or several) lives inside one warp and only __syncwarp is needed.
// FUSE2: x := coeff[c]*x + coeff2[c]*x2 at gather time and the transform coefficient is 1 --
// the lossy model's K(-1/rho) u + K(-delta/rho c^2) v in a single application (Lossy.hpp:230-232).
// ------------------------------------------------------------------------------------------------
template <int N>
struct ColCfg {
  static constexpr int NN = N * N;
  static constexpr bool WARP = (NN <= 32);
  static constexpr int CPW = WARP ? 32 / NN : 0; // cells per warp
  static constexpr int WPB = 4;                  // warps per block in WARP mode
  static constexpr int CPB = WARP ? CPW * WPB : (N == 6 ? 8 : 4); // cells per block
  static constexpr int THREADS = WARP ? 32 * WPB : ((CPB * NN + 31) / 32) * 32;
  static constexpr int NS = N | 1;                  // padded row length (doubles), odd
  static constexpr int PL = N * NS;                 // plane stride
  static constexpr int CS = ((N * PL + 7) / 8) * 8 + 8; // cell stride of one buffer (doubles)
  static constexpr int SMEM_BYTES = CPB * 3 * CS * (int)sizeof(double);
};

template <int N, bool FUSE2>
__global__ void __launch_bounds__(ColCfg<N>::THREADS)
    stiffness_col_kernel(const double* __restrict__ x, const double* __restrict__ x2,
                         double* __restrict__ y, const int32_t* __restrict__ dofmap,
                         const double2* __restrict__ G2, const double* __restrict__ coeff,
                         const double* __restrict__ coeff2, long long cell_begin,
                         long long cell_end, const __grid_constant__ DMat<N> D,
                         const HaloLaunch = HaloLaunch{}) {
  using C = ColCfg<N>;
  constexpr int NN = C::NN, NS = C::NS, PL = C::PL;
#ifdef FUS_HOST_EMULATION
  double* smem = fus_emu::dynamic_shared();
#else
  extern __shared__ double smem[];
#endif

  const int tid = threadIdx.x;
  int slot, t;
  bool lane_ok;
  if constexpr (C::WARP) {
    const int lane = tid & 31, w = tid >> 5;
    const int cw = lane / NN;
    t = lane - cw * NN;
    lane_ok = cw < C::CPW;
    slot = w * C::CPW + (lane_ok ? cw : 0);
  } else {
    slot = tid / NN;
    t = tid - slot * NN;
    lane_ok = slot < C::CPB;
    if (!lane_ok)
      slot = 0;
  }
  const int i1 = t / N, i2 = t - i1 * N;
  double* xs = smem + slot * (3 * C::CS);
  double* s1 = xs + C::CS;
  double* s2 = s1 + C::CS;

  // thread-dependent rows/columns of dphi (direction 1 and 2), kept in registers
  double D1[N], D2[N], DT1[N], DT2[N];
#pragma unroll
  for (int k = 0; k < N; ++k) {
    D1[k] = D.d[i1 * N + k];
    D2[k] = D.d[i2 * N + k];
    DT1[k] = D.d[k * N + i1];
    DT2[k] = D.d[k * N + i2];
  }

  const long long stride = (long long)gridDim.x * C::CPB;
  const long long ncell = cell_end - cell_begin;
  const int niter = (int)((ncell + stride - 1) / stride); // uniform over the block
  long long c = cell_begin + (long long)blockIdx.x * C::CPB + slot;

  int idx[N], idxn[N];
  double xv[N];
  double2 g[N][3];
  double cf = 0.0;
#pragma unroll
  for (int k = 0; k < N; ++k) {
    idx[k] = 0;
    idxn[k] = 0;
    xv[k] = 0.0;
#pragma unroll
    for (int p = 0; p < 3; ++p)
      g[k][p] = make_double2(0.0, 0.0);
  }

  // ---- prologue: everything for the first cell, dof indices for the second ----
  bool valid = lane_ok && (c < cell_end);
  if (valid) {
    const int32_t* dm = dofmap + c * (N * NN) + t;
#pragma unroll
    for (int k = 0; k < N; ++k)
      idx[k] = __ldg(dm + k * NN);
    if constexpr (FUSE2) {
      const double ca = __ldg(coeff + c), cb = __ldg(coeff2 + c);
#pragma unroll
      for (int k = 0; k < N; ++k)
        xv[k] = ca * __ldg(x + idx[k]) + cb * __ldg(x2 + idx[k]);
      cf = 1.0;
    } else {
#pragma unroll
      for (int k = 0; k < N; ++k)
        xv[k] = __ldg(x + idx[k]);
      cf = __ldg(coeff + c);
    }
    const double2* gp = G2 + c * (3 * N * NN) + t;
#pragma unroll
    for (int k = 0; k < N; ++k)
#pragma unroll
      for (int p = 0; p < 3; ++p)
        g[k][p] = ld_stream(gp + (k * 3 + p) * NN);
  }
  long long cn = c + stride;
  bool validn = lane_ok && (cn < cell_end);
  if (validn) {
    const int32_t* dm = dofmap + cn * (N * NN) + t;
#pragma unroll
    for (int k = 0; k < N; ++k)
      idxn[k] = __ldg(dm + k * NN);
  }

  for (int it = 0; it < niter; ++it) {
    // (a) stage the column, direction-0 contraction in registers
    if (lane_ok) {
#pragma unroll
      for (int k = 0; k < N; ++k)
        xs[k * PL + i1 * NS + i2] = xv[k];
    }
    double f0[N];
#pragma unroll
    for (int q = 0; q < N; ++q) {
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k)
        s = fma(D.d[q * N + k], xv[k], s);
      f0[q] = s;
    }
    const double cfc = cf;
    // gather the next cell's dofs now that xv has been consumed
    if (validn) {
      if constexpr (FUSE2) {
        const double ca = __ldg(coeff + cn), cb = __ldg(coeff2 + cn);
#pragma unroll
        for (int k = 0; k < N; ++k)
          xv[k] = ca * __ldg(x + idxn[k]) + cb * __ldg(x2 + idxn[k]);
      } else {
#pragma unroll
        for (int k = 0; k < N; ++k)
          xv[k] = __ldg(x + idxn[k]);
        cf = __ldg(coeff + cn);
      }
    }
    if constexpr (C::WARP)
      __syncwarp();
    else
      __syncthreads();

    // (b) per i0-level: directions 1,2 from shared memory, G transform, transposed direction 0
    double yv[N];
#pragma unroll
    for (int m = 0; m < N; ++m)
      yv[m] = 0.0;
    const double2* gpn = G2 + cn * (3 * N * NN) + t;
#pragma unroll
    for (int i0 = 0; i0 < N; ++i0) {
      double f1 = 0.0, f2 = 0.0;
      if (lane_ok) { // padding lanes must not load either: aliased addresses cost bank conflicts
#pragma unroll
        for (int k = 0; k < N; ++k) {
          f1 = fma(D1[k], xs[i0 * PL + k * NS + i2], f1);
          f2 = fma(D2[k], xs[i0 * PL + i1 * NS + k], f2);
        }
      }
      const double2 ga = g[i0][0], gb = g[i0][1], gc = g[i0][2];
      // stiffness::transform (spectral_op.hpp:113-130)
      const double t0 = cfc * (ga.x * f0[i0] + ga.y * f1 + gb.x * f2);
      const double t1 = cfc * (ga.y * f0[i0] + gb.y * f1 + gc.x * f2);
      const double t2 = cfc * (gb.x * f0[i0] + gc.x * f1 + gc.y * f2);
      // refill this level's registers with the next cell's factors
      if (validn) {
#pragma unroll
        for (int p = 0; p < 3; ++p)
          g[i0][p] = ld_stream(gpn + (i0 * 3 + p) * NN);
      }
#pragma unroll
      for (int m = 0; m < N; ++m)
        yv[m] = fma(D.d[i0 * N + m], t0, yv[m]);
      if (lane_ok) { // padding lanes alias slot 0 and must not write
        s1[i0 * PL + i1 * NS + i2] = t1;
        s2[i0 * PL + i1 * NS + i2] = t2;
      }
    }
    if constexpr (C::WARP)
      __syncwarp();
    else
      __syncthreads();

    // (c) transposed directions 1,2 and scatter-add
#pragma unroll
    for (int j0 = 0; j0 < N; ++j0) {
      if (valid) {
        double acc = yv[j0];
#pragma unroll
        for (int q = 0; q < N; ++q) {
          acc = fma(DT1[q], s1[j0 * PL + q * NS + i2], acc);
          acc = fma(DT2[q], s2[j0 * PL + i1 * NS + q], acc);
        }
        atomicAdd(y + idx[j0], acc);
      }
    }

    // rotate the pipeline
#pragma unroll
    for (int k = 0; k < N; ++k)
      idx[k] = idxn[k];
    valid = validn;
    c = cn;
    cn += stride;
    validn = lane_ok && (cn < cell_end);
    if (validn) {
      const int32_t* dm = dofmap + cn * (N * NN) + t;
#pragma unroll
      for (int k = 0; k < N; ++k)
        idxn[k] = __ldg(dm + k * NN);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// The tail of a HALO launch of the line kernel (fused peer transport, fus_halo_kernels.cuh).
// A HALO launch covers only the cells that touch a dof shared with a neighbour (stored first,
// [0, ninterface)); the rest of the mesh runs through the plain kernel in a second launch.  The
// exchange code sits AFTER the cell loop, so the loop is the plain kernel's (anything with waits or
// calls in front of or inside the loop cost ~150-300 B of spills per thread in the loop itself --
// measured with ptxas -v -- hence also: no forward wait here, the epilogue does it).
// A "group" is the set of warps that works on a cell slot: one warp (WARP) or GTHREADS/32 warps
// synchronising on the named barrier group + 1.
// ------------------------------------------------------------------------------------------------
template <bool WARP, int GTHREADS>
__device__ __forceinline__ void halo_group_sync(int group) {
  if constexpr (WARP)
    __syncwarp();
  else
#ifdef FUS_HOST_EMULATION
    fus_emu::named_barrier(group + 1, GTHREADS);
#else
    asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(GTHREADS) : "memory");
#endif
}

// After the last cell: the group reports in; once every group of the grid has (they all finish
// within a few microseconds of each other, and all are resident: the grid is sized by occupancy),
// the ghost part of y is final and the groups ship it to the owners' mailboxes, one 2 KB chunk at a
// time; the group that completes the last chunk raises the reverse flags.  The emulation runs
// blocks one after another, so there a group looks once and the last one ships everything.
template <bool WARP, int GTHREADS>
__device__ __forceinline__ void halo_group_report_and_ship(const FusedHalo* Hp, const double* y,
                                                        long long nown, int* word, int lg,
                                                        int group, unsigned int total) {
  const FusedHalo& H = *Hp;
  const long long nchunks = (H.nghost + kRevChunk - 1) / kRevChunk;
  __threadfence();
  halo_group_sync<WARP, GTHREADS>(group);
  if (lg == 0) {
    const unsigned int before = atomicAdd(H.ctr + CTR_GROUPS_PAST, 1u);
    if (before == total - 1 && nchunks == 0)
      halo_raise(H, false); // nothing to send: the exchange number advances all the same
    int all = ld_acquire_gpu(H.ctr + CTR_GROUPS_PAST) >= total ? 1 : 0;
#ifndef FUS_HOST_EMULATION
    if (!all && nchunks > 0) {
      const unsigned long long t0 = halo_time_ns();
      while (!(all = ld_acquire_gpu(H.ctr + CTR_GROUPS_PAST) >= total ? 1 : 0)) {
        __nanosleep(100);
        if (halo_time_ns() - t0 > H.timeout_ns || *(volatile int*)H.error) {
          atomicExch(H.error, 1);
          break;
        }
      }
    }
#endif
    *word = all;
  }
  halo_group_sync<WARP, GTHREADS>(group);
  const bool all_past = *word != 0;
  halo_group_sync<WARP, GTHREADS>(group);
  if (!all_past)
    return;
  for (;;) {
    if (lg == 0)
      *word = (int)atomicAdd(H.ctr + CTR_NEXT_CHUNK, 1u);
    halo_group_sync<WARP, GTHREADS>(group);
    const long long ch = *word;
    halo_group_sync<WARP, GTHREADS>(group);
    if (ch >= nchunks)
      break;
    const long long e1 = (ch + 1) * kRevChunk < H.nghost ? (ch + 1) * kRevChunk : H.nghost;
    for (long long e = ch * kRevChunk + lg; e < e1; e += GTHREADS) {
      int kq = 0;
      while (kq + 1 < H.nneigh && e >= H.roff[kq + 1])
        ++kq;
      H.r_rev[kq][e - H.roff[kq]] = __ldcg(y + nown + e);
    }
    __threadfence_system();
    halo_group_sync<WARP, GTHREADS>(group);
    if (lg == 0)
      *word = (long long)(atomicAdd(H.ctr + CTR_CHUNKS_DONE, 1u) + 1u) == nchunks ? 1 : 0;
    halo_group_sync<WARP, GTHREADS>(group);
    const bool last = *word != 0;
    halo_group_sync<WARP, GTHREADS>(group);
    if (last) { // one lane per neighbour raises its flag
      __shared__ unsigned long long raise_word;
      halo_raise_parallel(H, false, lg, &raise_word, [group] { halo_group_sync<WARP, GTHREADS>(group); });
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Stiffness operator, "line" kernel: same work split as the column kernel (N*N threads per cell, G
// streamed to registers with whole-cell look-ahead, RED scatter) but every contraction runs in
// registers.  Thread (a,b) owns one line of the cell in each of three layouts
//     A: (k,a,b) along i0      B: (a,k,b) along i1      C: (a,b,k) along i2
// and the tensors are re-laid-out through shared memory between the phases, so a thread moves
// 7N stores + 8N loads of distinct data per cell instead of N + 2N stores and 4N*N broadcast
// loads (each 64-bit shared access costs two wavefronts whether or not lanes share addresses,
// and L1TEX wavefronts are what bounds the column kernel).  dphi is a constant-bank operand in
// all six contractions, which also frees the 4N registers of per-thread dphi rows.
// Buffer strides are chosen per access-pattern pair (conflict-free for N = 5):
//     Sx (written A, read B and C), S1 (A and B only), S2 (A and C only).
// GPF = look-ahead depth of the G stream in i0-levels (a divisor of N; N = whole cell).
// ------------------------------------------------------------------------------------------------
template <int N>
struct LineCfg {
  static constexpr int NN = N * N;
  static constexpr bool WARP = (NN <= 32);
  // A "group" of GW warps works on GC cells and synchronises on its own: a warp (__syncwarp) when a
  // cell fits in one, else a named barrier per group, so groups of a block never wait on each other.
  static constexpr int GW = WARP ? 1 : (N == 6 ? 3 : 2);
  static constexpr int GC = WARP ? 32 / NN : (N == 6 ? 2 : 1);
  static constexpr int GROUPS = WARP ? 4 : (N == 6 ? 2 : 4); // sized so the register file fills
  static constexpr int CPB = GROUPS * GC;
  static constexpr int THREADS = GROUPS * GW * 32;
  // strides in doubles: element (i0,i1,i2) of buffer Z at i0*Z_S0 + i1*Z_S1 + i2
  static constexpr int X_S1 = N, X_S0 = N * N;
  static constexpr int B1_S1 = N, B1_S0 = (N == 5) ? 37 : (N * N + ((N % 2) ? 0 : 1));
  static constexpr int B2_S1 = N, B2_S0 = N * N;
  static constexpr int X_SZ = N * X_S0, B1_SZ = N * B1_S0, B2_SZ = N * B2_S0;
  static constexpr int CS = ((X_SZ + B1_SZ + B2_SZ + 7) / 8) * 8 + 8; // all three buffers of a cell
  static constexpr int SMEM_BYTES = CPB * CS * (int)sizeof(double);
  static constexpr int GPF = (N <= 7) ? N : N / 2;
  // GEOM 7: G of a cell staged in shared memory by TMA bulk copies, RING_STAGES cells deep per slot
  static constexpr int RING_STAGES = 2;
  static constexpr int CELLG = 6 * N * NN; // doubles of G per cell (48*N^3 bytes, a multiple of 16)
  static constexpr int SMEM_BYTES_RING
      = SMEM_BYTES + CPB * RING_STAGES * (CELLG + 1) * (int)sizeof(double);
  static constexpr bool RING_FITS = SMEM_BYTES_RING <= 227 * 1024;
};

// GEOM selects where the geometric factors come from (option "geometry_mode", see DESIGN):
//   0  streamed: G2[cell][i0][p][t], 48 B per point (the reference's data, the headline path)
//   1  affine:   every cell is a parallelepiped, so G[c][q] = w_q * Ghat[c] exactly (J is constant
//                in a cell).  G2 points to Ghat (3 double2 per CELL); the quadrature weight is
//                rebuilt from the 1-D weights.
//   2  trilinear: G2 points to the monomial coefficients of the cell map (FUS_TRI_STRIDE doubles
//                per CELL, fus_trilinear.hpp) and |det J| w K K^T f is evaluated per point from
//                them: exact for every mesh with a degree-1 coordinate element, 192 B per cell
//                instead of 48 B per point, ~45 more FP64 operations per point.
//   3  the same code as 2 compiled under a 128-register cap (4 blocks/SM for P <= 4 instead of 3:
//                16 warps/SM at the price of a few spilled values) -- an occupancy experiment that
//                bench.py's child sweep measures next to mode 2.
//   4, 5, 6  streamed G like 0 with a different software pipeline (option "stiffness_variant"
//                3, 4, 5; experiments for the stall the ncu source view of mode 0 shows).
//                ptxas puts every global load of the cell loop on ONE scoreboard, so the first
//                consumer of any loaded value waits for all loads in flight; in mode 0 that is a
//                register move of the carried coefficient at the loop end, a few instructions after
//                the dofmap rows of the next cell have been requested.
//                4 folds the coefficient into x when it is staged (K(c) x = K(1)(c x); no carried
//                copy, no move); 5 additionally prefetches the dofmap rows of the cell after next
//                into L2 at the loop end; 6 instead loads those rows into registers during phase 3
//                (with the G refills), so that every wait on the scoreboard finds the youngest load
//                at least two phases old; with FUSE2 it also keeps the two gathered vectors raw and
//                combines them when they are staged an iteration later (in mode 0 the multiply sits
//                right behind the loads).
//   7  pipeline 6 with the G stream taken off the scoreboard altogether: one TMA bulk copy per cell
//                (cp.async.bulk, 48*N^3 contiguous bytes) into a two-stage shared-memory ring per
//                cell slot, completion through an mbarrier, G read back with 16-byte shared loads.
//                No registers hold G in flight (12*N fewer live registers).  "stiffness_variant" 6.
// HALO (mesh partitioned, fused peer transport, fus_halo_kernels.cuh): the launch covers the cells
// that touch a dof shared with a neighbour.  Ghost values of the stage input are gathered straight
// from the mailbox the owners' epilogues write into (the epilogue BEFORE this launch has already
// waited for the owners' flags, so there is no wait here); after the last cell the groups ship the
// ghost part of y -- the partial sums the owners need -- into the owners' mailboxes and raise the
// reverse flags (halo_group_report_and_ship above, outside the cell loop).
template <int N, bool FUSE2, int GEOM = 0, typename T = double, bool HALO = false,
          bool REV = false>
__global__ void __launch_bounds__(LineCfg<N>::THREADS,
                                  (GEOM == 2 && N <= 5)
                                      ? 3
                                      : (((GEOM == 3 || GEOM == 5 || (GEOM == 6 && !FUSE2)
                                           || (HALO && GEOM != 6))
                                          && N <= 5)
                                             ? 4
                                             : ((GEOM == 6 && N == 6) ? 2 : 0)))
    stiffness_line_kernel(const T* __restrict__ x, const T* __restrict__ x2, T* __restrict__ y,
                          const int32_t* __restrict__ dofmap,
                          const typename Vec2<T>::type* __restrict__ G2,
                          const T* __restrict__ coeff, const T* __restrict__ coeff2,
                          long long cell_begin, long long cell_end,
                          const __grid_constant__ DMatT<T, N> D,
                          const __grid_constant__ HaloLaunch HL) {
  using C = LineCfg<N>;
  using V2 = typename Vec2<T>::type;
  static_assert(GEOM < 2 || sizeof(T) == sizeof(double),
                "the trilinear cell map is evaluated in FP64 only");
  constexpr int NN = C::NN, GPF = C::GPF;
  constexpr bool AFFINE = (GEOM == 1), TRI = (GEOM == 2 || GEOM == 3);
  constexpr bool STREAM = (GEOM == 0 || GEOM >= 4); // G read from memory per point
  constexpr bool RING = (GEOM == 7);                // ... through a TMA-fed shared-memory ring
  constexpr bool REGRING = STREAM && !RING;         // ... through the register ring g[][]
  constexpr bool CFX = (GEOM >= 4), DMPF = (GEOM == 5), DM2 = (GEOM == 6 || GEOM == 7);
  static_assert(!RING || sizeof(T) == sizeof(double), "the TMA ring is built for FP64 only");
  static_assert(!HALO || (sizeof(T) == sizeof(double) && STREAM), "fused halo: FP64, streamed G");
  constexpr int TQ = FUS_TRI_STRIDE / 2; // double2 per cell of trilinear coefficients
  static_assert(N % GPF == 0, "G look-ahead depth must divide N");
  static_assert(GEOM >= 0 && GEOM <= 7, "unknown geometry mode");
#ifdef FUS_HOST_EMULATION
  T* smem = reinterpret_cast<T*>(fus_emu::dynamic_shared());
#else
  extern __shared__ double smem_raw[];
  T* smem = reinterpret_cast<T*>(smem_raw);
#endif

  const int tid = threadIdx.x;
  const int group = tid / (C::GW * 32), lg = tid - group * (C::GW * 32);
  const int cg = lg / NN;
  const int t = lg - cg * NN;
  const bool lane_ok = cg < C::GC;
  const int slot = group * C::GC + (lane_ok ? cg : 0);
  const int a = t / N, b = t - a * N;
  T* Sx = smem + slot * C::CS;
  T* S1 = Sx + C::X_SZ;
  T* S2 = S1 + C::B1_SZ;
  // per-thread base offsets of the three access patterns
  T* SxA = Sx + a * C::X_S1 + b;   // + k*X_S0
  T* SxB = Sx + a * C::X_S0 + b;   // + k*X_S1
  T* SxC = Sx + a * C::X_S0 + b * C::X_S1; // + k
  T* S1A = S1 + a * C::B1_S1 + b;  // + k*B1_S0
  T* S1B = S1 + a * C::B1_S0 + b;  // + k*B1_S1
  T* S2A = S2 + a * C::B2_S1 + b;  // + k*B2_S0
  T* S2C = S2 + a * C::B2_S0 + b * C::B2_S1; // + k

  auto sync = [group] {
    if constexpr (C::WARP)
      __syncwarp();
    else
#ifdef FUS_HOST_EMULATION
      fus_emu::named_barrier(group + 1, C::GW * 32);
#else
      asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "n"(C::GW * 32) : "memory");
#endif
  };

  // RING: per cell slot, RING_STAGES stages of CELLG doubles behind the scratch buffers, then one
  // mbarrier per (slot, stage).  Thread t == 0 of a cell is its producer.
  constexpr unsigned RING_BYTES = (unsigned)(C::CELLG * sizeof(double));
  T* const ring = smem + C::CPB * C::CS + slot * (C::RING_STAGES * C::CELLG);
  unsigned long long* const ring_bar
      = reinterpret_cast<unsigned long long*>(smem + C::CPB * C::CS
                                              + C::CPB * C::RING_STAGES * C::CELLG)
        + slot * C::RING_STAGES;
  const bool producer = lane_ok && t == 0;
  bool ring_ok = true;
  if constexpr (RING) {
    if (producer)
      for (int sg = 0; sg < C::RING_STAGES; ++sg)
        ring_bar_init(ring_bar + sg);
    ring_bar_init_fence();
    __syncthreads();
  }

  // REV: the launch walks its cells from the last to the first.  Inside the RK4 loop the epilogue
  // before this kernel has written the stage input and the zeroed b in ascending dof order, so their
  // TAILS are what is still in L2 when this kernel starts, and the epilogue after it starts at the
  // dofs this kernel touched last.  A compile-time flag: as a run-time one it cost the cell loops
  // of P = 5 and P = 7 registers they do not have (spills inside the loop, -10 %).  Only the
  // instantiations the RK4 loop uses exist (fus_capi.cu); the direction is a hint, never needed.
  const long long ncell = cell_end - cell_begin;
  const long long pass = (long long)gridDim.x * C::CPB;
  const int niter = (int)((ncell + pass - 1) / pass);
  const long long stride = REV ? -pass : pass;
  auto in_range = [&](long long cell) { return REV ? cell >= cell_begin : cell < cell_end; };
  int* h_word = nullptr; // one word per group for broadcasts of the group leader's findings
  if constexpr (HALO) {
    __shared__ int h_words[C::GROUPS];
    h_word = h_words + group;
    if (*(volatile int*)HL.H->error)
      return; // an earlier wait timed out: the run is being aborted
    if (blockIdx.x == 0 && tid == 0) { // state for the NEXT epilogue; nothing here reads it
      HL.H->seq[SEQ_REV_EXPECT] += 1ull;
      HL.H->ctr[CTR_SHARED_DONE] = 0u;
      HL.H->ctr[CTR_EPI_NEXT] = 0u;
    }
  }
  // owned entries from the vector, ghost entries from the mailbox
  auto gx = [&](int i) -> T {
    if constexpr (HALO)
      return __ldg((i < HL.nown ? x : reinterpret_cast<const T*>(HL.mbu)) + i);
    else
      return __ldg(x + i);
  };
  auto gx2 = [&](int i) -> T {
    if constexpr (HALO)
      return __ldg((i < HL.nown ? x2 : reinterpret_cast<const T*>(HL.mbv)) + i);
    else
      return __ldg(x2 + i);
  };
  long long c = REV ? cell_end - 1 - ((long long)blockIdx.x * C::CPB + slot)
                    : cell_begin + (long long)blockIdx.x * C::CPB + slot;

  int idx[N], idxn[N];
  int idxnn[DM2 ? N : 1]; // DM2: dofmap rows of the cell after next
  constexpr bool RAW2 = DM2 && FUSE2; // second gathered vector kept raw until it is staged
  T xb[RAW2 ? N : 1];
  T can = T(0), cbn = T(0);
  T xv[N];
  V2 g[REGRING ? GPF : 1][3];
  V2 gh[3], ghn[3]; // AFFINE: Ghat of the current and of the next cell
  // TRI: the pieces of J on this thread's line (xi1,xi2) = (x[a],x[b]) for the current cell.  The
  // 192 B of a cell are read by all its threads at the same addresses, so there is nothing to keep
  // in flight: the next cell's two lines are prefetched into L2 and loaded when it becomes current.
  TriLine tl;
  const T wab = STREAM ? T(0) : D.w[a] * D.w[b];
  const double xia = TRI ? (double)D.x[a] : 0.0, xib = TRI ? (double)D.x[b] : 0.0;
  T cf = T(0);
  if constexpr (TRI) {
#pragma unroll
    for (int i = 0; i < 3; ++i) { // identity map: padding lanes stay finite
      tl.j0[i] = (i == 0), tl.a0[i] = (i == 1), tl.b0[i] = (i == 2);
      tl.da[i] = tl.db[i] = 0.0;
    }
  }
  // coefficients of cell `cell` -> line pieces
  auto tri_setup = [&](long long cell) {
    if constexpr (TRI) {
      double cq[FUS_TRI_STRIDE];
#pragma unroll
      for (int k = 0; k < TQ; ++k) {
        const double2 v = __ldg(G2 + cell * TQ + k);
        cq[2 * k] = v.x;
        cq[2 * k + 1] = v.y;
      }
      tri_line_setup(cq, xia, xib, tl);
    }
  };
  auto tri_prefetch = [&](long long cell) {
    const V2* q = G2 + cell * TQ;
#ifdef FUS_HOST_EMULATION
    (void)q;
#else
    asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(q + TQ - 1));
#endif
  };
#pragma unroll
  for (int p = 0; p < 3; ++p)
    gh[p] = ghn[p] = make_v2<T>(T(0), T(0));
#pragma unroll
  for (int k = 0; k < N; ++k) {
    idx[k] = 0;
    idxn[k] = 0;
    xv[k] = T(0);
  }
#pragma unroll
  for (int k = 0; k < (DM2 ? N : 1); ++k)
    idxnn[k] = 0;
#pragma unroll
  for (int k = 0; k < (RAW2 ? N : 1); ++k)
    xb[k] = T(0);
#pragma unroll
  for (int k = 0; k < (REGRING ? GPF : 1); ++k)
#pragma unroll
    for (int p = 0; p < 3; ++p)
      g[k][p] = make_v2<T>(T(0), T(0));

  bool valid = lane_ok && in_range(c);
  if (valid) {
    const int32_t* dm = dofmap + c * (N * NN) + t;
#pragma unroll
    for (int k = 0; k < N; ++k)
      idx[k] = __ldg(dm + k * NN);
    if constexpr (RAW2) {
      can = __ldg(coeff + c), cbn = __ldg(coeff2 + c);
#pragma unroll
      for (int k = 0; k < N; ++k) {
        xv[k] = gx(idx[k]);
        xb[k] = gx2(idx[k]);
      }
      cf = T(1);
    } else if constexpr (FUSE2) {
      const T ca = __ldg(coeff + c), cb = __ldg(coeff2 + c);
#pragma unroll
      for (int k = 0; k < N; ++k)
        xv[k] = ca * gx(idx[k]) + cb * gx2(idx[k]);
      cf = T(1);
    } else {
#pragma unroll
      for (int k = 0; k < N; ++k)
        xv[k] = gx(idx[k]);
      cf = __ldg(coeff + c);
    }
    if constexpr (AFFINE) {
#pragma unroll
      for (int p = 0; p < 3; ++p)
        gh[p] = __ldg(G2 + c * 3 + p);
    } else if constexpr (TRI) {
      tri_setup(c);
    } else if constexpr (RING) {
      if (producer)
        ring_issue(ring, G2 + c * (3 * N * NN), RING_BYTES, ring_bar);
    } else {
      const V2* gp = G2 + c * (3 * N * NN) + t;
#pragma unroll
      for (int k = 0; k < GPF; ++k)
#pragma unroll
        for (int p = 0; p < 3; ++p)
          g[k][p] = ld_stream(gp + (k * 3 + p) * NN);
    }
  }
  long long cn = c + stride;
  bool validn = lane_ok && in_range(cn);
  if (validn) {
    const int32_t* dm = dofmap + cn * (N * NN) + t;
#pragma unroll
    for (int k = 0; k < N; ++k)
      idxn[k] = __ldg(dm + k * NN);
    if constexpr (AFFINE) {
#pragma unroll
      for (int p = 0; p < 3; ++p)
        ghn[p] = __ldg(G2 + cn * 3 + p);
    }
    if constexpr (TRI)
      tri_prefetch(cn);
    if constexpr (RING) {
      if (producer)
        ring_issue(ring + C::CELLG, G2 + cn * (3 * N * NN), RING_BYTES, ring_bar + 1);
    }
  }

  for (int it = 0; it < niter; ++it) {
    // (1) stage x in layout A; direction-0 derivative in registers
    if constexpr (RAW2) { // gathered one iteration ago
#pragma unroll
      for (int k = 0; k < N; ++k)
        xv[k] = can * xv[k] + cbn * xb[k];
    } else if constexpr (CFX && !FUSE2) { // K(c) x = K(1) (c x): the coefficient is constant in a cell
#pragma unroll
      for (int k = 0; k < N; ++k)
        xv[k] = cf * xv[k];
    }
    if (lane_ok) {
#pragma unroll
      for (int k = 0; k < N; ++k)
        SxA[k * C::X_S0] = xv[k];
    }
    T f0[N];
#pragma unroll
    for (int q = 0; q < N; ++q) {
      T s = T(0);
#pragma unroll
      for (int k = 0; k < N; ++k)
        s = fma(D.d[q * N + k], xv[k], s);
      f0[q] = s;
    }
    const T cfc = CFX ? T(1) : cf;
    if (validn) {
      if constexpr (RAW2) {
        can = __ldg(coeff + cn), cbn = __ldg(coeff2 + cn);
#pragma unroll
        for (int k = 0; k < N; ++k) {
          xv[k] = gx(idxn[k]);
          xb[k] = gx2(idxn[k]);
        }
      } else if constexpr (FUSE2) {
        const T ca = __ldg(coeff + cn), cb = __ldg(coeff2 + cn);
#pragma unroll
        for (int k = 0; k < N; ++k)
          xv[k] = ca * gx(idxn[k]) + cb * gx2(idxn[k]);
      } else {
#pragma unroll
        for (int k = 0; k < N; ++k)
          xv[k] = gx(idxn[k]);
        cf = __ldg(coeff + cn);
      }
    }
    sync();

    // (2) direction-1 and direction-2 derivatives along the thread's own lines (layouts B, C).
    // Padding lanes are predicated off for loads too: their aliased addresses would add bank
    // conflicts (ncu: 3-4 wavefronts per LDS.64 instead of 2).
    if (lane_ok) {
      T l[N];
#pragma unroll
      for (int k = 0; k < N; ++k)
        l[k] = SxB[k * C::X_S1];
#pragma unroll
      for (int q = 0; q < N; ++q) {
        T s = T(0);
#pragma unroll
        for (int k = 0; k < N; ++k)
          s = fma(D.d[q * N + k], l[k], s);
        S1B[q * C::B1_S1] = s;
      }
#pragma unroll
      for (int k = 0; k < N; ++k)
        l[k] = SxC[k];
#pragma unroll
      for (int q = 0; q < N; ++q) {
        T s = T(0);
#pragma unroll
        for (int k = 0; k < N; ++k)
          s = fma(D.d[q * N + k], l[k], s);
        S2C[q] = s;
      }
    }
    sync();

    if constexpr (DM2) { // nothing consumes a global load between here and the end of the iteration
      const long long c2 = cn + stride;
      if (lane_ok && in_range(c2)) {
        const int32_t* dm = dofmap + c2 * (N * NN) + t;
#pragma unroll
        for (int k = 0; k < N; ++k)
          idxnn[k] = __ldg(dm + k * NN);
      }
    }

    // (3) back in layout A: G transform per level, transposed direction 0 in registers
    T yv[N];
#pragma unroll
    for (int m = 0; m < N; ++m)
      yv[m] = T(0);
    const V2* gpc = G2 + c * (3 * N * NN) + t;
    const V2* gpn = G2 + cn * (3 * N * NN) + t;
    // RING: this iteration's stage, filled by the copy issued two iterations ago
    const int stage = it % C::RING_STAGES;
    const V2* gring = reinterpret_cast<const V2*>(ring + stage * C::CELLG) + t;
    if constexpr (RING) { // after one time-out a thread stops waiting (wrong numbers, no stall)
      if (valid && ring_ok)
        ring_ok = ring_wait(ring_bar + stage, (unsigned)(it / C::RING_STAGES));
    }
#pragma unroll
    for (int i0 = 0; i0 < N; ++i0) {
      T f1 = T(0), f2 = T(0);
      if (lane_ok) {
        f1 = S1A[i0 * C::B1_S0];
        f2 = S2A[i0 * C::B2_S0];
      }
      T t0, t1, t2;
      if constexpr (TRI) {
        tri_transform(tl, D.x[i0], cfc * (D.w[i0] * wab), f0[i0], f1, f2, t0, t1, t2);
      } else {
        V2 ga, gb, gc;
        T scale;
        if constexpr (AFFINE) {
          ga = gh[0], gb = gh[1], gc = gh[2];
          scale = cfc * (D.w[i0] * wab);
        } else if constexpr (RING) {
          if (valid) {
            ga = gring[(i0 * 3 + 0) * NN], gb = gring[(i0 * 3 + 1) * NN],
            gc = gring[(i0 * 3 + 2) * NN];
          } else {
            ga = gb = gc = make_v2<T>(T(0), T(0));
          }
          scale = cfc;
        } else {
          ga = g[i0 % GPF][0], gb = g[i0 % GPF][1], gc = g[i0 % GPF][2];
          scale = cfc;
        }
        t0 = scale * (ga.x * f0[i0] + ga.y * f1 + gb.x * f2);
        t1 = scale * (ga.y * f0[i0] + gb.y * f1 + gc.x * f2);
        t2 = scale * (gb.x * f0[i0] + gc.x * f1 + gc.y * f2);
      }
      // refill the ring slot: level i0+GPF of this cell, or of the next cell once past the top
      if constexpr (!REGRING) {
      } else if (i0 + GPF < N) {
        if (valid) {
#pragma unroll
          for (int p = 0; p < 3; ++p)
            g[i0 % GPF][p] = ld_stream(gpc + ((i0 + GPF) * 3 + p) * NN);
        }
      } else if (validn) {
#pragma unroll
        for (int p = 0; p < 3; ++p)
          g[i0 % GPF][p] = ld_stream(gpn + ((i0 + GPF - N) * 3 + p) * NN);
      }
#pragma unroll
      for (int m = 0; m < N; ++m)
        yv[m] = fma(D.d[i0 * N + m], t0, yv[m]);
      if (lane_ok) { // same addresses this thread has just read
        S1A[i0 * C::B1_S0] = t1;
        S2A[i0 * C::B2_S0] = t2;
      }
    }
    sync();

    if constexpr (RING) { // every thread of the cell is past its reads of this stage: refill it
      const long long c2 = cn + stride; // the cell this slot works on two iterations from now
      if (producer && in_range(c2))
        ring_issue(ring + stage * C::CELLG, G2 + c2 * (3 * N * NN), RING_BYTES, ring_bar + stage);
    }

    // (4) transposed direction 1 and 2 along the thread's own lines, results back in place
    if (lane_ok) {
      T l[N];
#pragma unroll
      for (int q = 0; q < N; ++q)
        l[q] = S1B[q * C::B1_S1];
#pragma unroll
      for (int j = 0; j < N; ++j) {
        T s = T(0);
#pragma unroll
        for (int q = 0; q < N; ++q)
          s = fma(D.d[q * N + j], l[q], s);
        S1B[j * C::B1_S1] = s;
      }
#pragma unroll
      for (int q = 0; q < N; ++q)
        l[q] = S2C[q];
#pragma unroll
      for (int j = 0; j < N; ++j) {
        T s = T(0);
#pragma unroll
        for (int q = 0; q < N; ++q)
          s = fma(D.d[q * N + j], l[q], s);
        S2C[j] = s;
      }
    }
    sync();

    // (5) sum the three contributions in layout A and scatter-add
#pragma unroll
    for (int j0 = 0; j0 < N; ++j0) {
      if (valid) {
        const T acc = yv[j0] + S1A[j0 * C::B1_S0] + S2A[j0 * C::B2_S0];
        atomicAdd(y + idx[j0], acc);
      }
    }

#pragma unroll
    for (int k = 0; k < N; ++k)
      idx[k] = idxn[k];
    valid = validn;
    c = cn;
    cn += stride;
    validn = lane_ok && in_range(cn);
    if constexpr (AFFINE) {
#pragma unroll
      for (int p = 0; p < 3; ++p)
        gh[p] = ghn[p];
    }
    if constexpr (TRI) {
      if (valid)
        tri_setup(c);
    }
    if constexpr (DM2) { // loaded during phase (1) of this iteration
#pragma unroll
      for (int k = 0; k < N; ++k)
        idxn[k] = idxnn[k];
    } else if (validn) {
      const int32_t* dm = dofmap + cn * (N * NN) + t;
#pragma unroll
      for (int k = 0; k < N; ++k)
        idxn[k] = __ldg(dm + k * NN);
      if constexpr (AFFINE) {
#pragma unroll
        for (int p = 0; p < 3; ++p)
          ghn[p] = __ldg(G2 + cn * 3 + p);
      }
      if constexpr (TRI)
        tri_prefetch(cn);
    }
    if constexpr (DMPF) { // dofmap rows of the cell after next: N*N*N int32, touched line by line
      const long long c2 = cn + stride;
      if (lane_ok && in_range(c2) && t * 32 < N * NN) {
#ifndef FUS_HOST_EMULATION
        asm volatile("prefetch.global.L2 [%0];" ::"l"(dofmap + c2 * (N * NN) + t * 32));
#endif
      }
    }
  }
  if constexpr (HALO)
    halo_group_report_and_ship<C::WARP, C::GW * 32>(HL.H, reinterpret_cast<const double*>(y), HL.nown,
                                                    h_word, lg, group, gridDim.x * C::GROUPS);
}

// ------------------------------------------------------------------------------------------------
// Stiffness operator, "point" kernel: one thread per quadrature point, one cell per block pass.
// Deliberately plain; kept as the on-device cross-check of the column kernel.
// ------------------------------------------------------------------------------------------------
template <int N, bool FUSE2>
__global__ void __launch_bounds__(N* N* N)
    stiffness_point_kernel(const double* __restrict__ x, const double* __restrict__ x2,
                           double* __restrict__ y, const int32_t* __restrict__ dofmap,
                           const double2* __restrict__ G2, const double* __restrict__ coeff,
                           const double* __restrict__ coeff2, long long cell_begin,
                           long long cell_end, const __grid_constant__ DMat<N> D) {
  constexpr int NN = N * N, Nd = N * N * N;
  __shared__ double xs[Nd], s0[Nd], s1[Nd], s2[Nd];
  __shared__ double Ds[NN];
  const int q = threadIdx.x;
  const int i0 = q / NN, t = q - i0 * NN, i1 = t / N, i2 = t - i1 * N;
  if (q < NN)
    Ds[q] = D.d[q];
  for (long long c = cell_begin + blockIdx.x; c < cell_end; c += gridDim.x) {
    const int dof = dofmap[c * Nd + q];
    double cf;
    if constexpr (FUSE2) {
      xs[q] = coeff[c] * x[dof] + coeff2[c] * x2[dof];
      cf = 1.0;
    } else {
      xs[q] = x[dof];
      cf = coeff[c];
    }
    __syncthreads();
    double f0 = 0.0, f1 = 0.0, f2 = 0.0;
#pragma unroll
    for (int k = 0; k < N; ++k) {
      f0 = fma(Ds[i0 * N + k], xs[k * NN + i1 * N + i2], f0);
      f1 = fma(Ds[i1 * N + k], xs[i0 * NN + k * N + i2], f1);
      f2 = fma(Ds[i2 * N + k], xs[i0 * NN + i1 * N + k], f2);
    }
    const double2* gp = G2 + (c * N + i0) * (3 * NN) + t;
    const double2 ga = gp[0], gb = gp[NN], gc = gp[2 * NN];
    s0[q] = cf * (ga.x * f0 + ga.y * f1 + gb.x * f2);
    s1[q] = cf * (ga.y * f0 + gb.y * f1 + gc.x * f2);
    s2[q] = cf * (gb.x * f0 + gc.x * f1 + gc.y * f2);
    __syncthreads();
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < N; ++k) {
      acc = fma(Ds[k * N + i0], s0[k * NN + i1 * N + i2], acc);
      acc = fma(Ds[k * N + i1], s1[i0 * NN + k * N + i2], acc);
      acc = fma(Ds[k * N + i2], s2[i0 * NN + i1 * N + k], acc);
    }
    atomicAdd(y + dof, acc);
  }
}

// ------------------------------------------------------------------------------------------------
// Mass operator: y[dof] += coeff_c * detJ[c][q] * x[dof]   MassSpectral3D::operator(),
// spectral_op.hpp:69-86 with mass::transform :19-26.  One thread per quadrature point.
// ------------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256)
    mass_kernel(const double* __restrict__ x, double* __restrict__ y,
                const int32_t* __restrict__ dofmap, const double* __restrict__ detJ,
                const double* __restrict__ coeff, long long npoints, int Nd) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npoints; p += stride) {
    const long long c = p / Nd;
    const int dof = __ldg(dofmap + p);
    const double v = __ldg(coeff + c) * __ldg(x + dof) * __ldg(detJ + p);
    atomicAdd(y + dof, v);
  }
}

// FP32 instantiation of the operators (the reference's float runs,
// tests/test_operators3d/main.cpp:13): the stiffness operator is stiffness_line_kernel<N,FUSE2,0,
// float> on float copies of the cell data (24 B of G per point instead of 48), the mass operator
// the kernel below; the two conversion kernels make those copies once per context.
static __global__ void __launch_bounds__(256)
    mass_kernel_f32(const float* __restrict__ x, float* __restrict__ y,
                    const int32_t* __restrict__ dofmap, const float* __restrict__ detJ,
                    const float* __restrict__ coeff, long long npoints, int Nd) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npoints; p += stride) {
    const long long c = p / Nd;
    const int dof = __ldg(dofmap + p);
    atomicAdd(y + dof, __ldg(coeff + c) * __ldg(x + dof) * __ldg(detJ + p));
  }
}

static __global__ void __launch_bounds__(256)
    narrow_kernel(const double* __restrict__ in, float* __restrict__ out, long long n) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = (float)in[i];
}

// Mass operator of a lean context (no detJ array): |det J| w_q is rebuilt from the trilinear cell
// map, the same helpers as the geometry_mode = 2 stiffness kernel.  Only used at model creation
// (lumped mass) and by MassSpectral3D, so it is kept simple: one thread per point, the 21
// coefficients of a cell are broadcast loads shared by the lanes that work on it.
template <int N>
__global__ void __launch_bounds__(256)
    mass_tri_kernel(const double* __restrict__ x, double* __restrict__ y,
                    const int32_t* __restrict__ dofmap, const double* __restrict__ tri,
                    const double* __restrict__ coeff, long long cell_begin, long long npoints,
                    const __grid_constant__ Rule1D<N> R) {
  constexpr int NN = N * N, Nd = N * N * N;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npoints; p += stride) {
    const long long cl = p / Nd;
    const int q = (int)(p - cl * Nd);
    const int i0 = q / NN, t = q - i0 * NN, a = t / N, b = t - a * N;
    const long long c = cell_begin + cl;
    double cq[FUS_TRI_STRIDE];
#pragma unroll
    for (int k = 0; k < FUS_TRI_STRIDE / 2; ++k) {
      const double2 v = __ldg(reinterpret_cast<const double2*>(tri) + c * (FUS_TRI_STRIDE / 2) + k);
      cq[2 * k] = v.x;
      cq[2 * k + 1] = v.y;
    }
    TriLine L;
    tri_line_setup(cq, R.pts[a], R.pts[b], L);
    const double dj = tri_abs_det(L, R.pts[i0]) * (R.wts[i0] * (R.wts[a] * R.wts[b]));
    const int dof = __ldg(dofmap + c * Nd + q);
    atomicAdd(y + dof, __ldg(coeff + c) * __ldg(x + dof) * dj);
  }
}

// ------------------------------------------------------------------------------------------------
// Geometry setup on the device (runs once): per cell and quadrature point
//   J = sum_v x_v (x) grad phi_v^{Q1}(xi_q),  K = J^-1,  G = K K^T,
//   detJ[c][q] = |det J| w_q,  G2 <- |det J| w_q {G00,G01,G02,G11,G12,G22}
// compute_scaled_jacobian_determinant / compute_scaled_geometrical_factor, precompute.hpp:33-213.
// ------------------------------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(128)
    geometry_kernel(const double* __restrict__ xg, const int32_t* __restrict__ xdofmap,
                    long long ncells, double2* __restrict__ G2, double* __restrict__ detJ,
                    const __grid_constant__ Rule1D<N> R) {
  constexpr int NN = N * N, Nd = N * N * N;
  const long long total = ncells * Nd;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total;
       gid += stride) {
    const long long c = gid / Nd;
    const int q = (int)(gid - c * Nd);
    const int q0 = q / NN, t = q - q0 * NN, q1 = t / N, q2 = t - q1 * N;
    const double xi[3] = {R.pts[q0], R.pts[q1], R.pts[q2]};
    const double w = R.wts[q0] * R.wts[q1] * R.wts[q2];
    double J[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      const int a = v & 1, b = (v >> 1) & 1, cc = (v >> 2) & 1;
      const double l0 = a ? xi[0] : 1.0 - xi[0], l1 = b ? xi[1] : 1.0 - xi[1],
                   l2 = cc ? xi[2] : 1.0 - xi[2];
      const double d0 = a ? 1.0 : -1.0, d1 = b ? 1.0 : -1.0, d2 = cc ? 1.0 : -1.0;
      const double gr[3] = {d0 * l1 * l2, l0 * d1 * l2, l0 * l1 * d2};
      const double* X = xg + 3 * (long long)__ldg(xdofmap + 8 * c + v);
#pragma unroll
      for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
          J[i][j] += X[i] * gr[j];
    }
    const double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    const double c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    const double c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    const double det = J[0][0] * c00 + J[0][1] * c01 + J[0][2] * c02;
    const double dj = fabs(det) * w;
    if (detJ)
      detJ[gid] = dj;
    if (G2) {
      const double id = 1.0 / det;
      // K = adj(J)/det, row a of K
      const double K[3][3]
          = {{c00 * id, (J[0][2] * J[2][1] - J[0][1] * J[2][2]) * id,
              (J[0][1] * J[1][2] - J[0][2] * J[1][1]) * id},
             {c01 * id, (J[0][0] * J[2][2] - J[0][2] * J[2][0]) * id,
              (J[0][2] * J[1][0] - J[0][0] * J[1][2]) * id},
             {c02 * id, (J[0][1] * J[2][0] - J[0][0] * J[2][1]) * id,
              (J[0][0] * J[1][1] - J[0][1] * J[1][0]) * id}};
      double Gm[3][3];
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = a; b < 3; ++b)
          Gm[a][b] = K[a][0] * K[b][0] + K[a][1] * K[b][1] + K[a][2] * K[b][2];
      double2* o = G2 + (c * N + q0) * (3 * NN) + t;
      o[0] = make_double2(dj * Gm[0][0], dj * Gm[0][1]);
      o[NN] = make_double2(dj * Gm[0][2], dj * Gm[1][1]);
      o[2 * NN] = make_double2(dj * Gm[1][2], dj * Gm[2][2]);
    }
  }
}

// Monomial coefficients of the trilinear cell map (fus_trilinear.hpp), one thread per cell:
// what the trilinear-geometry operator reads instead of G.
static __global__ void __launch_bounds__(128)
    tri_coeff_kernel(const double* __restrict__ xg, const int32_t* __restrict__ xdofmap,
                     long long ncells, double* __restrict__ coeffs) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells)
    return;
  double X[8][3];
#pragma unroll
  for (int v = 0; v < 8; ++v) {
    const double* p = xg + 3 * (long long)__ldg(xdofmap + 8 * c + v);
    X[v][0] = p[0];
    X[v][1] = p[1];
    X[v][2] = p[2];
  }
  double out[FUS_TRI_STRIDE];
  tri_cell_coeffs(X, out);
#pragma unroll
  for (int k = 0; k < FUS_TRI_STRIDE; ++k)
    coeffs[c * FUS_TRI_STRIDE + k] = out[k];
}

// Affine-cell detection and compression: a cell is affine iff G[c][q]/w_q does not depend on q.
// One thread per cell; writes Ghat (from the first point) and clears *all_affine otherwise.
template <int N>
__global__ void __launch_bounds__(128)
    affine_detect_kernel(const double2* __restrict__ G2, long long ncells, double tol,
                         double2* __restrict__ Ghat, int* all_affine,
                         const __grid_constant__ DMat<N> D) {
  constexpr int NN = N * N;
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncells)
    return;
  const double2* gc = G2 + c * (3 * N * NN);
  const double w000 = D.w[0] * D.w[0] * D.w[0];
  double ref[6], mx = 0.0;
  for (int p = 0; p < 3; ++p) {
    const double2 v = gc[p * NN];
    ref[2 * p] = v.x / w000;
    ref[2 * p + 1] = v.y / w000;
  }
  for (int k = 0; k < 6; ++k)
    mx = fmax(mx, fabs(ref[k]));
  bool ok = true;
  for (int i0 = 0; i0 < N && ok; ++i0)
    for (int t = 0; t < NN && ok; ++t) {
      const double w = D.w[i0] * D.w[t / N] * D.w[t % N];
      for (int p = 0; p < 3; ++p) {
        const double2 v = gc[(i0 * 3 + p) * NN + t];
        if (fabs(v.x - w * ref[2 * p]) > tol * w * mx || fabs(v.y - w * ref[2 * p + 1]) > tol * w * mx)
          ok = false;
      }
    }
  for (int p = 0; p < 3; ++p)
    Ghat[c * 3 + p] = make_double2(ref[2 * p], ref[2 * p + 1]);
  if (!ok)
    atomicExch(all_affine, 0);
}

// Reference layout G[c][q][6] (a chunk of cells already on the device) -> G2, and back.
template <int N>
__global__ void __launch_bounds__(256)
    g_to_device_layout_kernel(const double* __restrict__ Gref, long long ncells_chunk,
                              double2* __restrict__ G2_chunk) {
  constexpr int NN = N * N, Nd = N * N * N;
  const long long total = ncells_chunk * Nd;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total;
       gid += stride) {
    const long long c = gid / Nd;
    const int q = (int)(gid - c * Nd);
    const int q0 = q / NN, t = q - q0 * NN;
    const double* s = Gref + gid * 6;
    double2* o = G2_chunk + (c * N + q0) * (3 * NN) + t;
    o[0] = make_double2(s[0], s[1]);
    o[NN] = make_double2(s[2], s[3]);
    o[2 * NN] = make_double2(s[4], s[5]);
  }
}

template <int N>
__global__ void __launch_bounds__(256)
    g_from_device_layout_kernel(const double2* __restrict__ G2_chunk, long long ncells_chunk,
                                double* __restrict__ Gref) {
  constexpr int NN = N * N, Nd = N * N * N;
  const long long total = ncells_chunk * Nd;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total;
       gid += stride) {
    const long long c = gid / Nd;
    const int q = (int)(gid - c * Nd);
    const int q0 = q / NN, t = q - q0 * NN;
    const double2* s = G2_chunk + (c * N + q0) * (3 * NN) + t;
    double* o = Gref + gid * 6;
    const double2 a = s[0], b = s[NN], cc = s[2 * NN];
    o[0] = a.x;
    o[1] = a.y;
    o[2] = b.x;
    o[3] = b.y;
    o[4] = cc.x;
    o[5] = cc.y;
  }
}

// ------------------------------------------------------------------------------------------------
// 2-D quadrilateral variant: StiffnessSpectral2D::operator(),
// cpp/fenicsx-sf-naive/common/spectral_op.hpp:275-318 with the 2-D stiffness::transform :196-208.
//   y[dof] += sum_cells B^T (coeff_c G_c) B x[dof],  G symmetric 2x2 per point
// One thread per point (i0,i1), whole cells per block, the two contractions of each phase read
// the cell through shared memory.  2-D problems are a few million dofs at most, so the kernel is
// written for clarity, not tuned like the hexahedral ones.  Device layout of the geometric
// factor: Gq[cell][p][q], p = 0..2 <-> {G00, G01, G11}, q = i0*N + i1 (coalesced per component).
// The mass operator, the boundary terms and the RK4 epilogue are dimension-independent and shared.
// ------------------------------------------------------------------------------------------------
template <int N>
struct QuadCfg {
  static constexpr int NN = N * N;
  static constexpr int CPB = (128 / NN) > 0 ? (128 / NN) : 1; // cells per block
  static constexpr int THREADS = ((CPB * NN + 31) / 32) * 32;
};

template <int N, bool FUSE2>
__global__ void __launch_bounds__(QuadCfg<N>::THREADS)
    stiffness_quad_kernel(const double* __restrict__ x, const double* __restrict__ x2,
                          double* __restrict__ y, const int32_t* __restrict__ dofmap,
                          const double* __restrict__ Gq, const double* __restrict__ coeff,
                          const double* __restrict__ coeff2, long long cell_begin,
                          long long cell_end, const __grid_constant__ DMat<N> D) {
  using C = QuadCfg<N>;
  constexpr int NN = C::NN;
  __shared__ double xs[C::CPB][NN], t0s[C::CPB][NN], t1s[C::CPB][NN];
  __shared__ double Ds[NN];
  const int tid = threadIdx.x;
  const int slot = tid / NN, t = tid - slot * NN;
  const bool lane_ok = slot < C::CPB;
  const int i0 = t / N, i1 = t - i0 * N;
  if (tid < NN)
    Ds[tid] = D.d[tid];
  const long long stride = (long long)gridDim.x * C::CPB;
  for (long long base = cell_begin + (long long)blockIdx.x * C::CPB; base < cell_end;
       base += stride) { // uniform over the block: every thread reaches every barrier
    const long long c = base + slot;
    const bool valid = lane_ok && c < cell_end;
    int dof = 0;
    double cf = 0.0;
    if (valid) {
      dof = __ldg(dofmap + c * NN + t);
      if constexpr (FUSE2) {
        xs[slot][t] = __ldg(coeff + c) * __ldg(x + dof) + __ldg(coeff2 + c) * __ldg(x2 + dof);
        cf = 1.0;
      } else {
        xs[slot][t] = __ldg(x + dof);
        cf = __ldg(coeff + c);
      }
    }
    __syncthreads();
    if (valid) {
      double d0 = 0.0, d1 = 0.0; // derivatives along reference directions 0 (slow) and 1 (fast)
#pragma unroll
      for (int k = 0; k < N; ++k) {
        d0 = fma(Ds[i0 * N + k], xs[slot][k * N + i1], d0);
        d1 = fma(Ds[i1 * N + k], xs[slot][i0 * N + k], d1);
      }
      const double* g = Gq + c * (3 * NN) + t;
      const double g00 = __ldg(g), g01 = __ldg(g + NN), g11 = __ldg(g + 2 * NN);
      t0s[slot][t] = cf * (g00 * d0 + g01 * d1);
      t1s[slot][t] = cf * (g01 * d0 + g11 * d1);
    }
    __syncthreads();
    if (valid) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < N; ++k) {
        acc = fma(Ds[k * N + i0], t0s[slot][k * N + i1], acc);
        acc = fma(Ds[k * N + i1], t1s[slot][i0 * N + k], acc);
      }
      atomicAdd(y + dof, acc);
    }
    __syncthreads(); // the next pass overwrites xs / t0s / t1s
  }
}

// Geometry of bilinear quadrilaterals on the device (runs once):
//   Gq <- |det J| w_q {G00, G01, G11},  detJ[c][q] = |det J| w_q
// (cpp/fenicsx-sf-naive/common/precompute.hpp:33-213 with gdim == 2).  xg is padded to 3
// coordinates per vertex as in DOLFINx; xdofmap has 4 vertices per cell, v = a + 2b.
template <int N>
__global__ void __launch_bounds__(128)
    geometry_quad_kernel(const double* __restrict__ xg, const int32_t* __restrict__ xdofmap,
                         long long ncells, double* __restrict__ Gq, double* __restrict__ detJ,
                         const __grid_constant__ Rule1D<N> R) {
  constexpr int NN = N * N;
  const long long total = ncells * NN;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x; gid < total;
       gid += stride) {
    const long long c = gid / NN;
    const int q = (int)(gid - c * NN);
    const int q0 = q / N, q1 = q - q0 * N;
    const double xi0 = R.pts[q0], xi1 = R.pts[q1];
    double J[2][2] = {{0, 0}, {0, 0}};
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int a = v & 1, b = v >> 1;
      const double l0 = a ? xi0 : 1.0 - xi0, l1 = b ? xi1 : 1.0 - xi1;
      const double g0 = (a ? 1.0 : -1.0) * l1, g1 = l0 * (b ? 1.0 : -1.0);
      const double* X = xg + 3 * (long long)__ldg(xdofmap + 4 * c + v);
      J[0][0] += X[0] * g0;
      J[0][1] += X[0] * g1;
      J[1][0] += X[1] * g0;
      J[1][1] += X[1] * g1;
    }
    const double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    const double dj = fabs(det) * (R.wts[q0] * R.wts[q1]);
    if (detJ)
      detJ[gid] = dj;
    if (Gq) {
      const double id = 1.0 / det;
      const double k00 = J[1][1] * id, k01 = -J[0][1] * id, k10 = -J[1][0] * id, k11 = J[0][0] * id;
      double* o = Gq + c * (3 * NN) + q;
      o[0] = dj * (k00 * k00 + k01 * k01);
      o[NN] = dj * (k00 * k10 + k01 * k11);
      o[2 * NN] = dj * (k10 * k10 + k11 * k11);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Boundary terms of the right-hand side over the compacted list of boundary dofs:
//   b[d] += g * src[k] + dg * dsrc[k] - absb[k] * v[d]
// -- fem::assemble_vector(b_, *L) with the collocated `ds` forms (Linear.hpp:205, forms.py).
// ------------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256)
    boundary_kernel(double* __restrict__ b, const double* __restrict__ v,
                    const int32_t* __restrict__ bidx, const double* __restrict__ bsrc,
                    const double* __restrict__ bdsrc, const double* __restrict__ babs,
                    long long nb, double g, double dg, const double* __restrict__ src_table,
                    const int* __restrict__ step_ctr, int stage) {
  // Inside fus_model_rk4 the source scalars come from a table computed on the host for every
  // (step, stage) with the reference's own time arithmetic, indexed by a device step counter, so
  // that one captured CUDA graph can replay any step.
  if (src_table) {
    const int s = *step_ctr;
    g = src_table[(s * 4 + stage) * 2];
    dg = src_table[(s * 4 + stage) * 2 + 1];
  }
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < nb) {
    const int d = bidx[k];
    b[d] += g * bsrc[k] + dg * bdsrc[k] - babs[k] * v[d];
  }
}

// ------------------------------------------------------------------------------------------------
// 32-byte global accesses with an L2 eviction priority (sm_100: LDG/STG.E.256 carry EFL2 / ELL2
// directly).  Streaming vectors that are not needed again soon are marked evict-first so that the
// lines the NEXT kernel wants (b, the next stage input) stay in the 126 MB L2.
// ------------------------------------------------------------------------------------------------
struct D4 {
  double v[4];
};
enum L2Pol { L2_NORMAL = 0, L2_FIRST = 1, L2_LAST = 2 };

template <int POL>
__device__ __forceinline__ D4 ld4(const double* p) {
  D4 r;
#ifdef FUS_HOST_EMULATION
  for (int k = 0; k < 4; ++k)
    r.v[k] = p[k];
#else
  if constexpr (POL == L2_FIRST)
    asm volatile("ld.global.L1::no_allocate.L2::evict_first.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3])
                 : "l"(p)
                 : "memory");
  else if constexpr (POL == L2_LAST)
    asm volatile("ld.global.L1::no_allocate.L2::evict_last.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3])
                 : "l"(p)
                 : "memory");
  else
    asm volatile("ld.global.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                 : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3])
                 : "l"(p)
                 : "memory");
#endif
  return r;
}

template <int POL>
__device__ __forceinline__ void st4(double* p, const D4& r) {
#ifdef FUS_HOST_EMULATION
  for (int k = 0; k < 4; ++k)
    p[k] = r.v[k];
#else
  if constexpr (POL == L2_FIRST)
    asm volatile("st.global.L2::evict_first.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(r.v[0]),
                 "d"(r.v[1]), "d"(r.v[2]), "d"(r.v[3])
                 : "memory");
  else if constexpr (POL == L2_LAST)
    asm volatile("st.global.L2::evict_last.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(r.v[0]),
                 "d"(r.v[1]), "d"(r.v[2]), "d"(r.v[3])
                 : "memory");
  else
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(r.v[0]), "d"(r.v[1]),
                 "d"(r.v[2]), "d"(r.v[3])
                 : "memory");
#endif
}

// ------------------------------------------------------------------------------------------------
// Fused RK4 stage epilogue (kernels (4)(5) of the north star).  After the operator has
// accumulated b = K(un[,vn]) + boundary terms, one pass does, per owned dof,
//     kv     = b / m                                   (Linear.hpp:212-221; Westervelt: m = m0 - dnl*un,
//                                                       b += dnl*vn^2, Westervelt.hpp:249-265)
//     ku     = vn                                      (f0, Linear.hpp:171-174)
//     un'    = u0 + a'*dt*ku ; vn' = v0 + a'*dt*kv     (:279-283 of the NEXT stage)
//     b      = boundary terms of the NEXT stage        (:203,205 of the NEXT stage; 0 off the boundary)
// and the solution update u += b_i dt ku, v += b_i dt kv (:293-294) WITHOUT a read-modify-write of
// the accumulators in every stage.  Because ku_i = vn_i and the stage inputs are themselves
// vn_1 = v0 + a_1 dt kv_0, vn_3 = v0 + a_3 dt kv_2, the classical tableau gives
//     u_new = [u0 + b_0 dt v0 + b_1 dt vn_1 + b_2 dt vn_2]            + b_3 dt vn_3
//     v_new = [(1 - b_0/a_1 - b_2/a_3) v0 + (b_0/a_1) vn_1 + b_1 dt kv_1] + (b_2/a_3) vn_3 + b_3 dt kv_3
// where both brackets are known in the stage-1 epilogue (vn_2 is produced there) and the rest in
// the stage-3 epilogue.  So ua, va are WRITTEN once (stage 1) and READ once (stage 3); the
// arithmetic differs from the reference's running sums by rounding only (b_0/a_1 = b_2/a_3 = 1/3).
// STAGE 0: the stage input is (u0,v0) itself.  STAGE 3: the new state goes straight into (u0,v0),
// the next step's stage-0 input.  Vector passes per stage: 7 / 10 / 8 / 8 (was 9 / 12 / 12 / 8).
//
// Boundary terms.  The collocated `ds` forms reduce to g*src[d] + dg*dsrc[d] - absb[d]*v[d] on
// boundary dofs (fem::assemble_vector(b_, *L), Linear.hpp:205).  v of the next stage is produced
// right here, so instead of zero-filling b and launching a boundary kernel, the block seeds b with
// the next stage's boundary terms: the compacted, dof-sorted boundary list is indexed per chunk of
// kStageChunk dofs (bchunk).  The source scalars come from the host-computed table of
// fus_model_rk4, row (step, stage) with the step counter on the device.
// ------------------------------------------------------------------------------------------------
constexpr int kStageThreads = 256;
constexpr int kStageVec = 4;                              // doubles per access (32 bytes)
constexpr int kStageChunk = kStageThreads * kStageVec;    // dofs per block pass

struct StageArgs {
  double* b;         // rhs accumulator; seeded for the next stage on exit (owned + ghosts)
  const double* m;   // lumped mass (m0 for Westervelt)
  const double* dnl; // Westervelt: D^(2 beta / rho^2 c^4) assembled; else nullptr
  double* u0;        // state at step start (stage-3 output)
  double* v0;
  double* ua; // partial sums of the update: written by stage 1, read by stage 3
  double* va;
  double* un; // stage inputs of stages 1..3 (read as ku = vn, rewritten for the next stage)
  double* vn;
  long long nowned;
  long long ntotal;
  double a_next_dt; // a_{i+1} * dt
  double bw_dt;     // b_i * dt
  double bw0_dt;    // stage 1: b_0 * dt
  double bw2_dt;    // stage 1: b_2 * dt
  double c_v0;      // stage 1: 1 - b_0/a_1 - b_2/a_3
  double c_vn;      // stage 1: b_0/a_1 ; stage 3: b_2/a_3
  int* step_ctr;    // advanced by the stage-3 epilogue (index into the source-scalar table)
  unsigned int* done_ctr; // stage 3: blocks finished (the last one advances step_ctr)
  // boundary terms of the next stage (nb == 0: b is zero-filled)
  long long nb;
  const int32_t* bidx;
  const double* bsrc;
  const double* bdsrc;
  const double* babs;
  const long long* bchunk;    // [nchunks + 1] first boundary entry of every chunk of kStageChunk dofs
  const double* src_table;    // (g, dg) per (step, stage)
  const FusedHalo* halo;      // fused exchange (HALO instantiations), else nullptr
  int halo_defer_wait;        // tests/emu only: the closing wait runs as a kernel of its own
};

// Classical RK4 tableau (Linear.hpp:263-265) -> the coefficients of stage `i` with step size dt
inline void stage_coefficients(StageArgs& A, int i, double dt) {
  const double a_runge[4] = {0.0, 0.5, 0.5, 1.0};
  const double b_runge[4] = {1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0};
  A.bw_dt = dt * b_runge[i];
  A.a_next_dt = (i < 3) ? dt * a_runge[i + 1] : 0.0;
  A.bw0_dt = dt * b_runge[0];
  A.bw2_dt = dt * b_runge[2];
  A.c_vn = (i == 1) ? b_runge[0] / a_runge[1] : b_runge[2] / a_runge[3];
  A.c_v0 = 1.0 - b_runge[0] / a_runge[1] - b_runge[2] / a_runge[3];
}

// one dof of one stage
template <int STAGE, bool WESTERVELT>
__device__ __forceinline__ void stage_dof(const StageArgs& A, double b, double m, double dnl,
                                          double u0, double v0, double un, double vn, double ua,
                                          double va, double& o_un, double& o_vn, double& o_ua,
                                          double& o_va) {
  // STAGE 0: un == u0, vn == v0 are passed by the caller
  if constexpr (WESTERVELT) {
    m = m - dnl * un;
    b = b + dnl * (vn * vn);
  }
  const double kv = b / m;
  if constexpr (STAGE < 3) {
    o_un = fma(vn, A.a_next_dt, u0);
    o_vn = fma(kv, A.a_next_dt, v0);
  }
  if constexpr (STAGE == 1) {
    o_ua = fma(o_vn, A.bw2_dt, fma(vn, A.bw_dt, fma(v0, A.bw0_dt, u0)));
    o_va = fma(kv, A.bw_dt, fma(vn, A.c_vn, A.c_v0 * v0));
  }
  if constexpr (STAGE == 3) {
    o_un = fma(vn, A.bw_dt, ua);                    // u_new
    o_vn = fma(kv, A.bw_dt, fma(vn, A.c_vn, va));   // v_new
  }
}

// HALO (mesh partitioned, fused peer transport; see fus_halo_kernels.cuh): the chunks that hold the
// dofs shared with neighbours, [0, nshared), come first.  Their blocks wait for the neighbours'
// partial sums of b, add them in a fixed order (send-list order: reproducible, unlike atomics), and
// store the next stage input straight into the neighbours' mailboxes; the block that finishes the
// last shared chunk raises the forward flags while the other ~99 % of the kernel is still running.
template <int STAGE, bool WESTERVELT, bool HINTS = false, bool HALO = false>
__global__ void __launch_bounds__(kStageThreads) rk4_stage_kernel(const StageArgs A) {
  constexpr int STREAM = HINTS ? L2_FIRST : L2_NORMAL; // vectors not needed by the next kernel
  const int tid = threadIdx.x;
  const long long nchunks = (A.ntotal + kStageChunk - 1) / kStageChunk;
  long long nshared = 0, shared_chunks = 0;
  if constexpr (HALO) {
    if (*(volatile int*)A.halo->error)
      return; // an earlier wait timed out: the run is being aborted
    nshared = A.halo->nshared;
    shared_chunks = (nshared + kStageChunk - 1) / kStageChunk;
    if (blockIdx.x == 0 && tid == 0) { // state for the NEXT operator; nothing here reads it
      for (int q = CTR_GROUPS_PAST; q <= CTR_CHUNKS_DONE; ++q)
        A.halo->ctr[q] = 0u;
    }
  }
  __shared__ int s_word;
  __shared__ unsigned long long s_raise;
  auto block_sync = [] { __syncthreads(); };
  bool rev_waited = false;
  double g = 0.0, dg = 0.0;
  if (A.nb) {
    const int s = *A.step_ctr; // stage 3 advances it only after every block has read it
    const int row = (STAGE < 3) ? (s * 4 + STAGE + 1) : ((s + 1) * 4);
    g = A.src_table[2 * row];
    dg = A.src_table[2 * row + 1];
  }
  double* const vnext = (STAGE < 3) ? A.vn : A.v0;
  // HALO: every block starts with chunk blockIdx.x and then takes the next free one from a counter:
  // the blocks that begin with a shared chunk (scalar work, remote stores, flags) are slower there
  // and must not be left with as many private chunks as everybody else.
  // The next chunk is claimed while the current one is being processed (one atomic per 1 024 dofs,
  // its latency hidden behind the chunk's own loads) and handed to the block through one of two
  // shared words, so a chunk ends with a single barrier.
  __shared__ int s_next[2];
  int parity = 0;
  long long ch = blockIdx.x;
  while (ch < nchunks) {
    if constexpr (HALO) {
      if (tid == 0)
        s_next[parity] = (int)atomicAdd(A.halo->ctr + CTR_EPI_NEXT, 1u);
      if (ch < shared_chunks && !rev_waited) { // uniform over the block
        rev_waited = true;
        if (!halo_wait_parallel(*A.halo, false, A.halo->seq[SEQ_REV_EXPECT], tid, &s_word, block_sync))
          return; // timed out: the run is being aborted
      }
    }
    const long long i = ch * kStageChunk + (long long)tid * kStageVec;
    if (i + kStageVec <= A.nowned && (!HALO || i >= nshared)) {
      D4 b = ld4<L2_NORMAL>(A.b + i), m = ld4<STREAM>(A.m + i);
      D4 u0, v0, un, vn, ua, va, dnl;
      if constexpr (STAGE < 3) {
        u0 = ld4<STREAM>(A.u0 + i);
        v0 = ld4<STREAM>(A.v0 + i);
      }
      if constexpr (STAGE == 0) {
        un = u0;
        vn = v0;
      } else {
        vn = ld4<L2_NORMAL>(A.vn + i);
        if constexpr (WESTERVELT)
          un = ld4<L2_NORMAL>(A.un + i);
      }
      if constexpr (STAGE == 3) {
        ua = ld4<STREAM>(A.ua + i);
        va = ld4<STREAM>(A.va + i);
      }
      if constexpr (WESTERVELT)
        dnl = ld4<STREAM>(A.dnl + i);
      D4 oun, ovn, oua, ova, zero;
#pragma unroll
      for (int k = 0; k < kStageVec; ++k) {
        stage_dof<STAGE, WESTERVELT>(A, b.v[k], m.v[k], WESTERVELT ? dnl.v[k] : 0.0,
                                     STAGE < 3 ? u0.v[k] : 0.0, STAGE < 3 ? v0.v[k] : 0.0,
                                     (WESTERVELT || STAGE == 0) ? un.v[k] : 0.0, vn.v[k],
                                     STAGE == 3 ? ua.v[k] : 0.0, STAGE == 3 ? va.v[k] : 0.0,
                                     oun.v[k], ovn.v[k], oua.v[k], ova.v[k]);
        zero.v[k] = 0.0;
      }
      if constexpr (STAGE < 3) {
        st4<L2_NORMAL>(A.un + i, oun);
        st4<L2_NORMAL>(A.vn + i, ovn);
      } else {
        st4<L2_NORMAL>(A.u0 + i, oun);
        st4<L2_NORMAL>(A.v0 + i, ovn);
      }
      if constexpr (STAGE == 1) {
        st4<STREAM>(A.ua + i, oua);
        st4<STREAM>(A.va + i, ova);
      }
      st4<L2_NORMAL>(A.b + i, zero);
    } else {
      // the chunk that straddles the owned / ghost boundary, and the ghost chunks: entry by entry
#pragma unroll
      for (int k = 0; k < kStageVec; ++k) {
        const long long j = i + k;
        if (j < A.nowned) {
          const double u0 = (STAGE < 3) ? A.u0[j] : 0.0, v0 = (STAGE < 3) ? A.v0[j] : 0.0;
          const double vn = (STAGE == 0) ? v0 : A.vn[j];
          const double un = (STAGE == 0) ? u0 : (WESTERVELT ? A.un[j] : 0.0);
          double oun = 0.0, ovn = 0.0, oua = 0.0, ova = 0.0;
          double bj = A.b[j];
          if constexpr (HALO) {
            if (j < nshared) { // ghost -> owner: the neighbours' partial sums, in send-list order
              const FusedHalo& H = *A.halo;
              for (int e = H.spos_off[j]; e < H.spos_off[j + 1]; ++e)
                bj += __ldcg(H.rev + H.spos[e]);
            }
          }
          stage_dof<STAGE, WESTERVELT>(A, bj, A.m[j], WESTERVELT ? A.dnl[j] : 0.0, u0, v0, un,
                                       vn, STAGE == 3 ? A.ua[j] : 0.0, STAGE == 3 ? A.va[j] : 0.0,
                                       oun, ovn, oua, ova);
          if constexpr (HALO) {
            if (j < nshared) { // owner -> ghost: the next stage input into the neighbours' mailboxes
              const FusedHalo& H = *A.halo;
              for (int e = H.spos_off[j]; e < H.spos_off[j + 1]; ++e) {
                const int pos = H.spos[e], k = H.spos_nb[e];
                const long long slot = pos - H.soff[k];
                H.r_fwd_u[k][slot] = oun;
                H.r_fwd_v[k][slot] = ovn;
              }
            }
          }
          if constexpr (STAGE < 3) {
            A.un[j] = oun;
            A.vn[j] = ovn;
          } else {
            A.u0[j] = oun;
            A.v0[j] = ovn;
          }
          if constexpr (STAGE == 1) {
            A.ua[j] = oua;
            A.va[j] = ova;
          }
          A.b[j] = 0.0;
        } else if (j < A.ntotal) {
          A.b[j] = 0.0; // ghost partial sums have been sent to their owners
        }
      }
    }
    if constexpr (HALO) {
      if (ch < shared_chunks) { // uniform over the block
        __threadfence_system(); // this thread's stores into the neighbours' mailboxes
        __syncthreads();
        if (tid == 0) {
          const unsigned int done = atomicAdd(A.halo->ctr + CTR_SHARED_DONE, 1u) + 1u;
          s_word = (long long)done == shared_chunks ? 1 : 0;
          if (s_word)
            __threadfence();
        }
        __syncthreads();
        const bool last = s_word != 0;
        __syncthreads();
        if (last) // one thread per neighbour raises its forward flag
          halo_raise_parallel(*A.halo, true, tid, &s_raise, block_sync);
      }
    }
    if (A.nb) { // uniform over the block
      const long long kb = A.bchunk[ch], ke = A.bchunk[ch + 1];
      if (ke > kb) {
        __syncthreads(); // the block's own stores of v' and of the zeros are ordered before this
        for (long long k = kb + tid; k < ke; k += kStageThreads) {
          const int d = A.bidx[k];
          A.b[d] = g * A.bsrc[k] + dg * A.bdsrc[k] - A.babs[k] * __ldcg(vnext + d);
        }
      }
    }
    if constexpr (HALO) { // next chunk: first come, first served
      __syncthreads();
      ch = (long long)gridDim.x + s_next[parity];
      parity ^= 1;
    } else {
      ch += gridDim.x;
    }
  }
  if constexpr (HALO) {
    if (blockIdx.x == 0 && tid == 0 && shared_chunks == 0)
      halo_raise(*A.halo, true); // nothing to send: the exchange number advances all the same
    if (blockIdx.x == 0) {
      // The next operator gathers ghost values from the mailbox without waiting: this kernel does
      // not end before the neighbours' forward data of this exchange have landed.  Their epilogues
      // send first thing, as this one did above, so the wait is over long before the private part
      // of this kernel is -- and every rank raises its own flags before it waits: no cycle.
      if (!A.halo_defer_wait) {
        __syncthreads();
        halo_wait_parallel(*A.halo, true, A.halo->seq[SEQ_FWD_EXPECT] + 1ull, tid, &s_word, block_sync);
        if (tid == 0)
          A.halo->seq[SEQ_FWD_EXPECT] += 1ull;
      }
    }
  }
  if constexpr (STAGE == 3) {
    if (A.step_ctr && A.done_ctr) { // the last block to finish advances the step counter
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        if (atomicAdd(A.done_ctr, 1u) == gridDim.x - 1) {
          *A.done_ctr = 0u;
          *A.step_ctr += 1;
        }
      }
    }
  }
}

// out = b / m (Westervelt: with the solution-dependent terms) -- single f1 evaluation for tests
template <bool WESTERVELT>
__global__ void __launch_bounds__(256)
    f1_finish_kernel(const double* __restrict__ b, const double* __restrict__ m,
                     const double* __restrict__ dnl, const double* __restrict__ un,
                     const double* __restrict__ vn, double* __restrict__ out, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    double bb = b[i], mm = m[i];
    if constexpr (WESTERVELT) {
      mm = mm - dnl[i] * un[i];
      bb = bb + dnl[i] * (vn[i] * vn[i]);
    }
    out[i] = bb / mm;
  }
}

// y[i] += a[i]
static __global__ void __launch_bounds__(256)
    add_kernel(double* __restrict__ y, const double* __restrict__ a, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    y[i] += a[i];
}

// y[i] = value
static __global__ void __launch_bounds__(256)
    fill_kernel(double* __restrict__ y, double value, long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    y[i] = value;
}

} // namespace fus
