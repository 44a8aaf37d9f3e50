// Stiffness operator, "cell" kernel for P = 2 (N = 3): one THREAD owns a whole cell.
//   y[dof] += sum_cells B^T (coeff_c G_c) B x[dof]        StiffnessSpectral3D::operator(),
//                                                         spectral_op.hpp:173-243
// OPT-IN (option "stiffness_variant" 7), NOT the default: measured 0.358-0.364 ms against the
// column kernel's 0.347 ms on the 107^3-cell box of the degree sweep (profiles/r2l_r2s_cell_kernel_*).
// Why it was built: at P = 2 the column kernel is bound by the L1TEX data pipe (93 % in
// profiles/r2p_ncu_full_stiffness_P2.json), and 30 of its ~60 wavefronts per cell are the
// shared-memory exchanges between the 9 threads of a cell.  With 27 dofs a cell fits the registers
// of one thread (27 inputs + 27 outputs), so nothing is exchanged at all: the derivatives at a point
// are 3 x 3 FMAs on the thread's own values (sum_factorisation.hpp:43-86 collapses to that for
// N = 3), the transposed contractions likewise.
// For the global accesses to coalesce the cell data are kept in blocks of 32 cells, lane-minor:
//   Gt  double2 [block][q = i0*9 + t][p][lane]          (the entries of G2, regrouped)
//   dmt int32   [block][q][lane]
// made once from G2 / tensor_dofmap by transpose_cells_kernel (the cost: a second copy of G and of
// the dofmap for P = 2 contexts, 52 B per point).
// G does not pass through registers on its way in.  The first version streamed it with 16-byte
// loads into a register ring; ptxas tracks all of a loop's global loads on one scoreboard, so each
// use of a ring slot waited for the refill issued just before it and the ring was one deep in
// effect whatever its size (3 or 9 points: 0.339-0.361 ms at 255 registers, 0.417 at 168 with
// spills).  Now every warp owns a shared-memory ring of kCellStages rows (a row = 3 points x 3 x
// 32 lanes x 16 B = 4.5 KB, contiguous in Gt) that lane 0 keeps full with cp.async.bulk copies
// completing on mbarriers -- no registers and no scoreboard are tied up while the bytes travel, the
// ring runs on across cell boundaries, and the copies carry an L2 evict-first policy (G is read
// once; without it the stream pushed x and y out of L2 between the touches of neighbouring cells:
// 2.25 GB of DRAM traffic and 0.392 ms instead of 2.00 GB and 0.364).
// What the measurements say about the rest (ncu captures in profiles/, timing variants of the same
// run): L1TEX is down to 58 %, DRAM traffic is at the algorithmic bytes, the FP64 pipe at 16 % --
// and the kernel without its scatter runs at the copy peak (0.266 ms) while the scatter adds 0.1 ms
// however it is shaped (27 lane-per-sector REDs per cell, or 18 paired ones after the in-warp face
// hand-over below; dof indices re-read or kept in registers).  The column kernel pays the same for
// its REDs but hides it behind nine times as many threads.  Left in the library as a measured
// alternative and as the starting point should FP64 REDs get cheaper.
#pragma once

namespace fus {

constexpr int kCellLanes = 32;   // cells per block of the transposed arrays (= a warp)
constexpr int kCellThreads = 128;
constexpr int kCellWarps = kCellThreads / 32;
constexpr int kCellStages = 3;                          // rows in flight per warp (9 % stages == 0)
constexpr int kCellRowV2 = 3 * 3 * kCellLanes;          // double2 per row: 3 points x 3 x 32 lanes
constexpr int kCellRowBytes = kCellRowV2 * 16;
constexpr int kCellSmemBytes
    = kCellWarps * kCellStages * kCellRowBytes + kCellWarps * kCellStages * 8;

static __global__ void transpose_cells_kernel(const double2* __restrict__ G2,
                                              const int32_t* __restrict__ dofmap,
                                              double2* __restrict__ Gt, int32_t* __restrict__ dmt,
                                              long long ncells, int Nd) {
  // one thread per (padded cell, entry); entries beyond ncells are zero / dof 0
  const int NN = Nd / 3; // N = 3: 9 threads' worth of entries per level
  const long long nblk = (ncells + kCellLanes - 1) / kCellLanes;
  const long long total = nblk * 3 * Nd * kCellLanes;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int lane = (int)(i % kCellLanes);
    const long long be = i / kCellLanes;
    const int e = (int)(be % (3 * Nd)); // destination entry q*3 + p
    const int q = e / 3, p = e - 3 * q;
    const int i0 = q / NN, t = q - i0 * NN;
    const long long blk = be / (3 * Nd);
    const long long cell = blk * kCellLanes + lane;
    Gt[i] = cell < ncells ? G2[cell * (3 * Nd) + (i0 * 3 + p) * NN + t] : make_double2(0.0, 0.0);
    if (p == 0)
      dmt[(blk * Nd + q) * kCellLanes + lane] = cell < ncells ? dofmap[cell * Nd + q] : 0;
  }
}

template <bool FUSE2>
__global__ void __launch_bounds__(kCellThreads, 3)
    stiffness_cell_kernel(const double* __restrict__ x, const double* __restrict__ x2,
                          double* __restrict__ y, const int32_t* __restrict__ dmt,
                          const double2* __restrict__ Gt, const double* __restrict__ coeff,
                          const double* __restrict__ coeff2, long long cell_begin,
                          long long cell_end, const __grid_constant__ DMat<3> D, int reverse) {
  constexpr int N = 3, Nd = 27, W = kCellLanes, S = kCellStages, ROWS = 9;
  static_assert(ROWS % S == 0, "static ring slots");
#ifdef FUS_HOST_EMULATION
  double* smem = fus_emu::dynamic_shared();
#else
  extern __shared__ __align__(128) double smem[];
#endif
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  double2* const ring = reinterpret_cast<double2*>(smem) + wib * (S * kCellRowV2);
  unsigned long long* const bar
      = reinterpret_cast<unsigned long long*>(smem + kCellWarps * S * kCellRowV2 * 2) + wib * S;

  const long long warp = (long long)blockIdx.x * kCellWarps + wib;
  const long long nwarps = (long long)gridDim.x * kCellWarps;
  const long long blk_first = cell_begin / W, blk_last = (cell_end - 1) / W; // inclusive
  const long long nblk = blk_last - blk_first + 1;
  if (warp >= nblk)
    return;
  const long long step = reverse ? -nwarps : nwarps;
  const long long blk0 = reverse ? blk_last - warp : blk_first + warp;
  const int niter = (int)((nblk - warp + nwarps - 1) / nwarps);
  const int nrows = niter * ROWS;

  // row n of this warp's stream: block blk0 + (n / 9) * step, rows are contiguous inside a block
  auto row_src = [&](int n) {
    const long long b = blk0 + (long long)(n / ROWS) * step;
    return Gt + (b * ROWS + (n % ROWS)) * kCellRowV2;
  };
  if (lane == 0) {
    for (int sg = 0; sg < S; ++sg)
      ring_bar_init(bar + sg);
  }
  ring_bar_init_fence();
  __syncwarp();
  // G is read once: its lines are marked evict-first in L2 so that they do not push out x and y,
  // which the neighbouring cells of the same sweep still need
  const unsigned long long pol = l2_policy_evict_first();
  auto issue = [&](int sg, int n) {
    ring_issue_hint(ring + sg * kCellRowV2, row_src(n), kCellRowBytes, bar + sg, pol);
  };
  if (lane == 0) {
    for (int sg = 0; sg < S && sg < nrows; ++sg)
      issue(sg, sg);
  }
  bool ring_ok = true;

  long long blk = blk0;
  for (int it = 0; it < niter; ++it, blk += step) {
    const long long cell = blk * W + lane;
    const bool valid = cell >= cell_begin && cell < cell_end;
    const long long cs = valid ? cell : cell_begin; // padding lanes compute, write nothing
    const int32_t* dm = dmt + blk * (Nd * W) + lane;
#ifndef FUS_HOST_EMULATION
    if (it + 1 < niter && lane < Nd) // the next block's 27 rows of dof indices: into L2 meanwhile
      asm volatile("prefetch.global.L2 [%0];" ::"l"(dmt + (blk + step) * (Nd * W) + lane * W));
#endif

    double u[Nd], w[Nd];
    int di[Nd]; // kept for the scatter
    double cf;
#pragma unroll
    for (int q = 0; q < Nd; ++q)
      di[q] = __ldg(dm + q * W);
    if constexpr (FUSE2) {
      const double ca = __ldg(coeff + cs), cb = __ldg(coeff2 + cs);
#pragma unroll
      for (int q = 0; q < Nd; ++q)
        u[q] = ca * __ldg(x + di[q]) + cb * __ldg(x2 + di[q]);
      cf = 1.0;
    } else {
#pragma unroll
      for (int q = 0; q < Nd; ++q)
        u[q] = __ldg(x + di[q]);
      cf = __ldg(coeff + cs);
    }
#pragma unroll
    for (int q = 0; q < Nd; ++q)
      w[q] = 0.0;

#pragma unroll
    for (int r = 0; r < ROWS; ++r) {
      const int sg = r % S; // static: 9 % S == 0
      const int n = it * ROWS + r;
      if (ring_ok) // after one time-out the warp stops waiting (wrong numbers, no stall)
        ring_ok = ring_wait(bar + sg, (unsigned)(n / S));
      const double2* gs = ring + sg * kCellRowV2 + lane;
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int q = r * 3 + j;
        const int i0 = q / 9, i1 = (q / 3) % 3, i2 = q % 3;
        const double2 ga = gs[(j * 3 + 0) * W], gb = gs[(j * 3 + 1) * W], gc = gs[(j * 3 + 2) * W];
        double d0 = 0.0, d1 = 0.0, d2 = 0.0;
#pragma unroll
        for (int k = 0; k < N; ++k) {
          d0 = fma(D.d[i0 * N + k], u[(k * 3 + i1) * 3 + i2], d0);
          d1 = fma(D.d[i1 * N + k], u[(i0 * 3 + k) * 3 + i2], d1);
          d2 = fma(D.d[i2 * N + k], u[(i0 * 3 + i1) * 3 + k], d2);
        }
        // stiffness::transform (spectral_op.hpp:113-130)
        const double t0 = cf * (ga.x * d0 + ga.y * d1 + gb.x * d2);
        const double t1 = cf * (ga.y * d0 + gb.y * d1 + gc.x * d2);
        const double t2 = cf * (gb.x * d0 + gc.x * d1 + gc.y * d2);
#pragma unroll
        for (int k = 0; k < N; ++k) {
          w[(k * 3 + i1) * 3 + i2] = fma(D.d[i0 * N + k], t0, w[(k * 3 + i1) * 3 + i2]);
          w[(i0 * 3 + k) * 3 + i2] = fma(D.d[i1 * N + k], t1, w[(i0 * 3 + k) * 3 + i2]);
          w[(i0 * 3 + i1) * 3 + k] = fma(D.d[i2 * N + k], t2, w[(i0 * 3 + i1) * 3 + k]);
        }
      }
      __syncwarp(); // every lane is past its reads of this row: refill the slot
      if (lane == 0 && n + S < nrows)
        issue(sg, n + S);
    }
    {
      // ---- scatter-add.  Lane-per-cell REDs put every lane in a sector of its own (27
      // instructions x 32 sector requests per block of cells).  Two warp-level steps cut the
      // request count to a third, for any mesh and numbering (they compare dof indices, they
      // assume nothing) -- measured: worth 1 % here, kept because it never costs:
      //  (a) the cell of the next lane usually is the neighbour across the face i2 = 2, whose dofs
      //      are that cell's i2 = 0 dofs: where the indices agree the value is handed over by
      //      shuffle and added there in registers (9 of 27 values never leave the warp);
      //  (b) the dofs (i0,i1,0) and (i0,i1,1) of a cell are neighbours in memory under a cell-
      //      blocked numbering: lane pairs swap one value so that an even/odd pair of lanes updates
      //      the two dofs of ONE cell in the same instruction -- one sector request carries both.
      constexpr unsigned FULL = 0xffffffffu;
      const bool even = (lane & 1) == 0;
      const bool up_ok = valid && lane < 31 && cell + 1 < cell_end; // the next lane scatters too
      const bool pvalid = __shfl_xor_sync(FULL, valid ? 1 : 0, 1) != 0;
#pragma unroll
      for (int f = 0; f < 9; ++f) { // f = i0*3 + i1
        const int q0 = f * 3, q1 = f * 3 + 1, q2 = f * 3 + 2;
        const int d0 = di[q0], d1 = di[q1], d2 = di[q2];
        const int nb0 = __shfl_down_sync(FULL, d0, 1);
        const bool give = up_ok && nb0 == d2;
        const double got = __shfl_up_sync(FULL, give ? w[q2] : 0.0, 1);
        if (lane > 0)
          w[q0] += got;
        if (valid && !give)
          atomicAdd(y + d2, w[q2]);
        const double rv = __shfl_xor_sync(FULL, even ? w[q1] : w[q0], 1);
        const int rd = __shfl_xor_sync(FULL, even ? d1 : d0, 1);
        // first instruction: the even lane's cell (its dofs 0 and 1); second: the odd lane's
        if (even ? valid : pvalid)
          atomicAdd(y + (even ? d0 : rd), even ? w[q0] : rv);
        if (even ? pvalid : valid)
          atomicAdd(y + (even ? rd : d1), even ? rv : w[q1]);
      }
    }
  }
}

} // namespace fus
