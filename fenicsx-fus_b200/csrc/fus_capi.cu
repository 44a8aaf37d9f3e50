// fus_capi.cu -- C ABI (include/fus_b200.h) on top of the sm_100a kernels in fus_kernels.cuh.
//
// One fus_ctx per GPU/process; all launches of a context go to one stream and are
// stream-ordered.  There is no CPU fallback: without a usable device every compute entry
// point returns FUS_ERR_CUDA.
#include "fus_halo.hpp"
#include "fus_internal.hpp"
#include "fus_kernels.cuh"
#include "fus_cell_kernel.cuh"

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

namespace fus {

static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
}

#define FUS_CUDA(call)                                                                             \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) {                                                                      \
      fus::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));       \
      return FUS_ERR_CUDA;                                                                         \
    }                                                                                              \
  } while (0)

#define FUS_TRY(call)                                                                              \
  do {                                                                                             \
    int r__ = (call);                                                                              \
    if (r__ != FUS_OK)                                                                             \
      return r__;                                                                                  \
  } while (0)

#define FUS_LAUNCHED()                                                                             \
  do {                                                                                             \
    fus::g_launches.fetch_add(1, std::memory_order_relaxed);                                       \
    FUS_CUDA(cudaGetLastError());                                                                  \
  } while (0)

static inline int grid_for(long long n, int block, int max_blocks) {
  long long g = (n + block - 1) / block;
  if (g < 1)
    g = 1;
  if (g > max_blocks)
    g = max_blocks;
  return (int)g;
}

} // namespace fus

using namespace fus;

struct fus_ctx {
  int dim = 3;              // 3: hexahedra (cpp/fenicsx-sf), 2: quadrilaterals (cpp/fenicsx-sf-naive)
  int P = 0, N = 0, Nd = 0;
  int64_t ncells = 0, ndofs = 0, nowned = 0;
  int device = 0;
  int num_sms = 148;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int32_t* d_dofmap = nullptr;
  double2* d_G2 = nullptr;
  double* d_Gq = nullptr;   // dim == 2: Gq[cell][3][N*N]
  double* d_detJ = nullptr;
  double dphi[64];
  double pts[8], wts[8];    // 1-D GLL points and weights
  // optional compressed geometry (option "geometry_mode"): 1 = one Ghat per affine cell,
  // 2 = trilinear map coefficients per cell, G rebuilt in the kernel (fus_trilinear.hpp)
  double2* d_Ghat = nullptr;
  double* d_tri = nullptr;
  // float copies of G2 / detJ for the FP32 operator entry points, made on first use
  float* d_G2f = nullptr;
  float* d_detJf = nullptr;
  // cell data in blocks of 32 cells, lane-minor, for the cell-per-thread kernel (P = 2), made on
  // first use (fus_cell_kernel.cuh)
  double2* d_Gt = nullptr;
  int32_t* d_dmt = nullptr;
  int geom_active = 0;      // what the stiffness operator currently uses: 0 streamed G, 1, 2, 3
  bool lean = false;        // neither G nor detJ exist on the device: always mode 2
  int live_models = 0;      // fus_model objects that still point at this context
  // -1 auto (column kernel for P <= 3, line kernel for P >= 4: measured crossover, see
  // profiles/), 0 column kernel, 1 point kernel, 2 line kernel, 3 / 4 / 5 / 6 line kernel with the
  // software pipelines (kernel GEOM 4 / 5 / 6 / 7), 7 cell-per-thread kernel (P = 2 only)
  int variant = -1;
  int col_blocks_per_sm = 0;
  int reserve_sms = 0;      // SMs left free for the halo kernels while cells overlap with them
  int halo_reserve = 4;     // reserve_sms inside a partitioned stage, NCCL side-stream mode
  int l2_persist = 0;       // keep the rhs accumulator b resident in L2 during rk4 (option)
  int stage_hints = 0;      // epilogue: streaming vectors marked L2 evict-first (option)
  int reverse_op = 1;       // RK4 loop: the operator walks the cells backwards (option "reverse_operator")
  int reverse_cells = 0;    // what the next stiffness launch does (set by assemble_rhs)
  int use_graph = 1;        // replay RK4 steps from a captured CUDA graph (option "use_graph")
  long long config_epoch = 0; // bumped by anything that changes what a step launches
  Halo* halo = nullptr;
  // optional per-kernel event timing (bench.py roofline): family 0 stiffness, 1 stage, 2 boundary
  bool profile = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events[3];
  size_t prof_used[3] = {0, 0, 0};
};

namespace {
// cudaFuncSetAttribute and the occupancy query are per device: a process that drives several GPUs
// (one context each) must configure every kernel on every device it launches it on.
// Contexts on different devices may be driven from different host threads, so the first-use
// configuration is done under a lock and published through an atomic flag.
constexpr int kMaxDevices = 64;
struct KernelCfg {
  std::mutex mu;
  std::atomic<bool> configured[kMaxDevices] = {};
  int blocks_plain[kMaxDevices] = {}, blocks_fuse[kMaxDevices] = {};
};

// RAII event bracket around one launch; a no-op unless profiling is on.
struct ProfScope {
  fus_ctx* c;
  int fam;
  cudaStream_t st;
  cudaEvent_t stop = nullptr;
  ProfScope(fus_ctx* c_, int fam_, cudaStream_t st_) : c(c_), fam(fam_), st(st_) {
    if (!c->profile)
      return;
    auto& ev = c->prof_events[fam];
    if (c->prof_used[fam] == ev.size()) {
      cudaEvent_t a, b;
      if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess)
        return;
      ev.push_back({a, b});
    }
    auto& pr = ev[c->prof_used[fam]++];
    cudaEventRecord(pr.first, st);
    stop = pr.second;
  }
  ~ProfScope() {
    if (stop)
      cudaEventRecord(stop, st);
  }
};
} // namespace

namespace {
// device scratch that is released on every exit path
template <typename T>
struct DevPtr {
  T* p = nullptr;
  DevPtr() = default;
  DevPtr(const DevPtr&) = delete;
  DevPtr& operator=(const DevPtr&) = delete;
  ~DevPtr() { cudaFree(p); }
};
} // namespace

struct fus_model {
  fus_ctx* ctx = nullptr;
  int kind = 0;
  double freq = 0, p0 = 0, s0 = 0, w0 = 0, period = 0, window_length = 4.0;
  // per-cell operator coefficients
  double *d_lin = nullptr, *d_att = nullptr;
  // per-dof
  double *d_m = nullptr, *d_dnl = nullptr;
  // compacted boundary lists
  int64_t nb = 0;
  int32_t* d_bidx = nullptr;
  double *d_bsrc = nullptr, *d_bdsrc = nullptr, *d_babs = nullptr;
  long long* d_bchunk = nullptr; // first boundary entry of every epilogue chunk (kStageChunk dofs)
  unsigned int* d_done = nullptr; // stage-3 epilogue: blocks finished
  // state (u0,v0 double as u_n,v_n) and work vectors
  double *d_u0 = nullptr, *d_v0 = nullptr, *d_ua = nullptr, *d_va = nullptr, *d_un = nullptr,
         *d_vn = nullptr, *d_b = nullptr;
  // rk4: per-(step,stage) source scalars computed on the host, step counter on the device
  double* d_src = nullptr;
  size_t src_cap = 0;
  int* d_stepctr = nullptr;
  // one RK4 step captured as a CUDA graph (replayed while dt, stream and halo mode stay the same)
  cudaGraphExec_t step_graph = nullptr;
  double graph_dt = 0.0;
  cudaStream_t graph_stream = nullptr;
  int graph_halo_mode = -2;
  long long graph_epoch = -1;
  long long graph_launches = 0;
  bool use_graph = true;
};

namespace {

int select_device(fus_ctx* c) {
  FUS_CUDA(cudaSetDevice(c->device));
  return FUS_OK;
}

// ---- per-degree dispatch ---------------------------------------------------------------------

// Kernel per degree, from the measured sweep of every variant at every degree on a B200
// (profiles/r2a_variant_sweep.jsonl, one application on ~10 M dofs, fraction of the measured HBM
// peak): P<=3 column kernel (0.82 / 0.94), P=4 line kernel (0.91), P=5 and P=6 the line kernel
// with the dofmap rows loaded next to the G refills (0.89 / 0.81; the default pipeline gives
// 0.83 / 0.70), P=7 the pipeline with the coefficient folded into x (0.76 vs 0.73).
int resolved_variant(const fus_ctx* c) {
  const int N = c->N;
  int v = (c->variant >= 0) ? c->variant : (N <= 4 ? 0 : (N == 5 ? 2 : (N <= 7 ? 5 : 3)));
  if (v == 7 && (N != 3 || c->geom_active != 0 || c->dim != 3))
    v = 0; // the cell-per-thread kernel exists for P = 2 with streamed G only
  return v;
}

// Lane-minor copies of G and of the dofmap for the cell-per-thread kernel.  Allocates: must not
// run while the stream is being captured (fus_model_rk4 calls it before it captures).
int ensure_cell_layout(fus_ctx* c) {
  if (c->d_Gt || resolved_variant(c) != 7 || !c->d_G2)
    return FUS_OK;
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  FUS_CUDA(cudaStreamIsCapturing(c->stream, &cap));
  if (cap != cudaStreamCaptureStatusNone) {
    set_error("cell-per-thread kernel: its cell data must be made before a stream capture");
    return FUS_ERR_STATE;
  }
  const long long nblk = (c->ncells + kCellLanes - 1) / kCellLanes;
  FUS_CUDA(cudaMalloc(&c->d_Gt, sizeof(double2) * 3 * c->Nd * kCellLanes * nblk));
  FUS_CUDA(cudaMalloc(&c->d_dmt, sizeof(int32_t) * c->Nd * kCellLanes * nblk));
  transpose_cells_kernel<<<grid_for(nblk * 3 * c->Nd * kCellLanes, 256, c->num_sms * 8), 256, 0,
                           c->stream>>>(c->d_G2, c->d_dofmap, c->d_Gt, c->d_dmt, c->ncells, c->Nd);
  FUS_LAUNCHED();
  FUS_CUDA(cudaStreamSynchronize(c->stream)); // once; later launches may use another stream
  return FUS_OK;
}

template <int N>
int launch_stiffness_n(fus_ctx* c, const double* x, const double* x2, const double* coeff,
                       const double* coeff2, double* y, long long cb, long long ce,
                       cudaStream_t st, const FusedHalo* fh) {
  if (ce <= cb)
    return FUS_OK;
  if (ce - cb >= (1ll << 30)) { // the kernels count their walk through the cells in 32 bits
    set_error("stiffness operator: more than 2^30 cells in one launch");
    return FUS_ERR_UNSUPPORTED;
  }
  DMat<N> D;
  std::memcpy(D.d, c->dphi, sizeof(double) * N * N);
  std::memcpy(D.w, c->wts, sizeof(double) * N);
  std::memcpy(D.x, c->pts, sizeof(double) * N);
  const bool fuse = (x2 != nullptr);
  if (fh) { // partitioned mesh, fused peer transport: the cells that touch shared dofs
    using L = LineCfg<N>;
    const int v = (c->variant >= 0) ? c->variant : (N <= 5 ? 2 : (N <= 7 ? 5 : 3));
    auto go = [&](auto kern_plain, auto kern_fuse, KernelCfg& cfg) -> int {
      std::atomic<bool>& configured = cfg.configured[c->device];
      int &bps_plain = cfg.blocks_plain[c->device], &bps_fuse = cfg.blocks_fuse[c->device];
      if (!configured.load(std::memory_order_acquire)) {
        std::lock_guard<std::mutex> lock(cfg.mu);
        if (!configured.load(std::memory_order_relaxed)) {
          FUS_CUDA(cudaFuncSetAttribute(kern_plain, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        L::SMEM_BYTES));
          FUS_CUDA(cudaFuncSetAttribute(kern_fuse, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        L::SMEM_BYTES));
          FUS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps_plain, kern_plain, L::THREADS,
                                                                L::SMEM_BYTES));
          FUS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps_fuse, kern_fuse, L::THREADS,
                                                                L::SMEM_BYTES));
          if (bps_plain < 1 || bps_fuse < 1) {
            set_error("stiffness kernel <N=%d> does not fit on an SM", N);
            return FUS_ERR_CUDA;
          }
          configured.store(true, std::memory_order_release);
        }
      }
      const int bps = fuse ? bps_fuse : bps_plain;
      const long long want = (ce - cb + L::CPB - 1) / L::CPB;
      const int blocks = (int)std::min<long long>(want, (long long)c->num_sms * bps);
      const HaloLaunch HL = halo_fused_launch(c->halo);
      ProfScope prof(c, 0, st);
      if (fuse)
        kern_fuse<<<blocks, L::THREADS, L::SMEM_BYTES, st>>>(x, x2, y, c->d_dofmap, c->d_G2, coeff,
                                                             coeff2, cb, ce, D, HL);
      else
        kern_plain<<<blocks, L::THREADS, L::SMEM_BYTES, st>>>(x, x2, y, c->d_dofmap, c->d_G2, coeff,
                                                              coeff2, cb, ce, D, HL);
      FUS_LAUNCHED();
      return FUS_OK;
    };
    if (v >= 5) { // dofmap rows loaded next to the G refills (kernel GEOM 6)
      static KernelCfg cfg;
      return go(stiffness_line_kernel<N, false, 6, double, true>,
                stiffness_line_kernel<N, true, 6, double, true>, cfg);
    }
    if (v >= 3) { // coefficient folded into x (kernel GEOM 4)
      static KernelCfg cfg;
      return go(stiffness_line_kernel<N, false, 4, double, true>,
                stiffness_line_kernel<N, true, 4, double, true>, cfg);
    }
    static KernelCfg cfg;
    return go(stiffness_line_kernel<N, false, 0, double, true>,
              stiffness_line_kernel<N, true, 0, double, true>, cfg);
  }
  const int variant = resolved_variant(c);
  if constexpr (N == 3) {
    if (variant == 7) { // a thread per cell, cell data in lane-minor blocks of 32 cells
      FUS_TRY(ensure_cell_layout(c));
      auto go = [&](auto kern_plain, auto kern_fuse, KernelCfg& cfg) -> int {
        std::atomic<bool>& configured = cfg.configured[c->device];
        int& bps = cfg.blocks_plain[c->device];
        if (!configured.load(std::memory_order_acquire)) {
          std::lock_guard<std::mutex> lock(cfg.mu);
          int bf = 0;
          FUS_CUDA(cudaFuncSetAttribute(kern_plain, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        kCellSmemBytes));
          FUS_CUDA(cudaFuncSetAttribute(kern_fuse, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        kCellSmemBytes));
          FUS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, kern_plain, kCellThreads,
                                                                kCellSmemBytes));
          FUS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bf, kern_fuse, kCellThreads,
                                                                kCellSmemBytes));
          bps = std::min(bps, bf);
          if (bps < 1) {
            set_error("cell-per-thread stiffness kernel does not fit on an SM");
            return FUS_ERR_CUDA;
          }
          configured.store(true, std::memory_order_release);
        }
        ProfScope prof(c, 0, st);
        const long long nblk = (ce - 1) / kCellLanes - cb / kCellLanes + 1;
        const long long want = (nblk * 32 + kCellThreads - 1) / kCellThreads;
        const int sms = std::max(1, c->num_sms - c->reserve_sms);
        const int per_sm = c->col_blocks_per_sm > 0 ? std::min(bps, c->col_blocks_per_sm) : bps;
        const int blocks = (int)std::min<long long>(want, (long long)sms * per_sm);
        if (fuse)
          kern_fuse<<<blocks, kCellThreads, kCellSmemBytes, st>>>(x, x2, y, c->d_dmt, c->d_Gt, coeff, coeff2, cb,
                                                     ce, D, c->reverse_cells ? 1 : 0);
        else
          kern_plain<<<blocks, kCellThreads, kCellSmemBytes, st>>>(x, x2, y, c->d_dmt, c->d_Gt, coeff, coeff2, cb,
                                                      ce, D, c->reverse_cells ? 1 : 0);
        FUS_LAUNCHED();
        return FUS_OK;
      };
      static KernelCfg cfg;
      return go(stiffness_cell_kernel<false>, stiffness_cell_kernel<true>, cfg);
    }
  }
  if (variant == 1 && c->geom_active == 0) {
    ProfScope prof(c, 0, st);
    const int blocks = (int)std::min<long long>(ce - cb, (long long)c->num_sms * 16);
    if (fuse)
      stiffness_point_kernel<N, true><<<blocks, N * N * N, 0, st>>>(
          x, x2, y, c->d_dofmap, c->d_G2, coeff, coeff2, cb, ce, D);
    else
      stiffness_point_kernel<N, false><<<blocks, N * N * N, 0, st>>>(
          x, x2, y, c->d_dofmap, c->d_G2, coeff, coeff2, cb, ce, D);
    FUS_LAUNCHED();
    return FUS_OK;
  }
  // variant 0: column kernel, variant 2: line kernel (same launch geometry rules)
  const double2* Gptr = c->d_G2;
  auto launch = [&](auto kern_plain, auto kern_fuse, int threads, int smem_bytes, int cpb,
                    KernelCfg& cfg) -> int {
    std::atomic<bool>& configured = cfg.configured[c->device];
    int &bps_plain = cfg.blocks_plain[c->device], &bps_fuse = cfg.blocks_fuse[c->device];
    if (!configured.load(std::memory_order_acquire)) {
      std::lock_guard<std::mutex> lock(cfg.mu);
      if (!configured.load(std::memory_order_relaxed)) {
        FUS_CUDA(cudaFuncSetAttribute(kern_plain, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      smem_bytes));
        FUS_CUDA(cudaFuncSetAttribute(kern_fuse, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      smem_bytes));
        FUS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps_plain, kern_plain, threads,
                                                              smem_bytes));
        FUS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps_fuse, kern_fuse, threads,
                                                              smem_bytes));
        if (bps_plain < 1 || bps_fuse < 1) {
          set_error("stiffness kernel <N=%d> does not fit on an SM", N);
          return FUS_ERR_CUDA;
        }
        configured.store(true, std::memory_order_release);
      }
    }
    ProfScope prof(c, 0, st);
    int bps = fuse ? bps_fuse : bps_plain;
    if (c->col_blocks_per_sm > 0)
      bps = std::min(bps, c->col_blocks_per_sm);
    const long long want = (ce - cb + cpb - 1) / cpb;
    const int sms = std::max(1, c->num_sms - c->reserve_sms);
    const int blocks = (int)std::min<long long>(want, (long long)sms * bps);
    const HaloLaunch HL{};
    if (fuse)
      kern_fuse<<<blocks, threads, smem_bytes, st>>>(x, x2, y, c->d_dofmap, Gptr, coeff, coeff2,
                                                     cb, ce, D, HL);
    else
      kern_plain<<<blocks, threads, smem_bytes, st>>>(x, x2, y, c->d_dofmap, Gptr, coeff, coeff2,
                                                      cb, ce, D, HL);
    FUS_LAUNCHED();
    return FUS_OK;
  };
  if (c->geom_active == 1) { // all cells are parallelepipeds: Ghat per cell instead of G per point
    using L = LineCfg<N>;
    static KernelCfg cfg;
    Gptr = c->d_Ghat;
    return launch(stiffness_line_kernel<N, false, 1>, stiffness_line_kernel<N, true, 1>,
                  L::THREADS, L::SMEM_BYTES, L::CPB, cfg);
  }
  if (c->geom_active == 2) { // trilinear cells: G rebuilt per point from 192 B per cell
    using L = LineCfg<N>;
    static KernelCfg cfg;
    Gptr = reinterpret_cast<const double2*>(c->d_tri);
    return launch(stiffness_line_kernel<N, false, 2>, stiffness_line_kernel<N, true, 2>,
                  L::THREADS, L::SMEM_BYTES, L::CPB, cfg);
  }
  if (c->geom_active == 3) { // same, compiled under a 128-register cap (occupancy experiment)
    using L = LineCfg<N>;
    static KernelCfg cfg;
    Gptr = reinterpret_cast<const double2*>(c->d_tri);
    return launch(stiffness_line_kernel<N, false, 3>, stiffness_line_kernel<N, true, 3>,
                  L::THREADS, L::SMEM_BYTES, L::CPB, cfg);
  }
  // Inside the RK4 loop (reverse_cells, set by assemble_rhs) the operator walks the cells backwards,
  // for the L2 reuse with the epilogues on either side.  The direction is a template flag of the
  // line kernel, instantiated for the kernels the library picks at P = 4, 5, 6; any other choice
  // walks forward (it is a hint).  P = 7 stays forward: its kernel has no register to spare.
  if (c->reverse_cells) {
    using L = LineCfg<N>;
    if constexpr (N == 5) {
      if (variant == 2) {
        static KernelCfg cfg;
        return launch(stiffness_line_kernel<N, false, 0, double, false, true>,
                      stiffness_line_kernel<N, true, 0, double, false, true>, L::THREADS,
                      L::SMEM_BYTES, L::CPB, cfg);
      }
    }
    if constexpr (N == 6 || N == 7) {
      if (variant == 5) {
        static KernelCfg cfg;
        return launch(stiffness_line_kernel<N, false, 6, double, false, true>,
                      stiffness_line_kernel<N, true, 6, double, false, true>, L::THREADS,
                      L::SMEM_BYTES, L::CPB, cfg);
      }
    }
  }
  if (variant == 2) {
    using L = LineCfg<N>;
    static KernelCfg cfg;
    return launch(stiffness_line_kernel<N, false>, stiffness_line_kernel<N, true>, L::THREADS,
                  L::SMEM_BYTES, L::CPB, cfg);
  }
  if (variant == 3) { // pipeline experiment: coefficient of the current cell (kernel GEOM 4)
    using L = LineCfg<N>;
    static KernelCfg cfg;
    return launch(stiffness_line_kernel<N, false, 4>, stiffness_line_kernel<N, true, 4>, L::THREADS,
                  L::SMEM_BYTES, L::CPB, cfg);
  }
  if (variant == 4) { // ... and the dofmap of the cell after next prefetched into L2 (GEOM 5)
    using L = LineCfg<N>;
    static KernelCfg cfg;
    return launch(stiffness_line_kernel<N, false, 5>, stiffness_line_kernel<N, true, 5>, L::THREADS,
                  L::SMEM_BYTES, L::CPB, cfg);
  }
  if (variant == 5 || (variant == 6 && !LineCfg<N>::RING_FITS)) { // dofmap rows with the G refills
    using L = LineCfg<N>;
    static KernelCfg cfg;
    return launch(stiffness_line_kernel<N, false, 6>, stiffness_line_kernel<N, true, 6>, L::THREADS,
                  L::SMEM_BYTES, L::CPB, cfg);
  }
  if constexpr (LineCfg<N>::RING_FITS) {
    if (variant == 6) { // G through a TMA-fed shared-memory ring (GEOM 7)
      using L = LineCfg<N>;
      static KernelCfg cfg;
      return launch(stiffness_line_kernel<N, false, 7>, stiffness_line_kernel<N, true, 7>,
                    L::THREADS, L::SMEM_BYTES_RING, L::CPB, cfg);
    }
  }
  using C = ColCfg<N>;
  static KernelCfg cfg;
  return launch(stiffness_col_kernel<N, false>, stiffness_col_kernel<N, true>, C::THREADS,
                C::SMEM_BYTES, C::CPB, cfg);
}

template <int N>
int launch_stiffness_quad_n(fus_ctx* c, const double* x, const double* x2, const double* coeff,
                            const double* coeff2, double* y, long long cb, long long ce,
                            cudaStream_t st) {
  if (ce <= cb)
    return FUS_OK;
  using Q = QuadCfg<N>;
  DMat<N> D;
  std::memcpy(D.d, c->dphi, sizeof(double) * N * N);
  std::memcpy(D.w, c->wts, sizeof(double) * N);
  std::memcpy(D.x, c->pts, sizeof(double) * N);
  ProfScope prof(c, 0, st);
  const long long want = (ce - cb + Q::CPB - 1) / Q::CPB;
  const int blocks = (int)std::min<long long>(want, (long long)c->num_sms * 8);
  if (x2)
    stiffness_quad_kernel<N, true><<<blocks, Q::THREADS, 0, st>>>(x, x2, y, c->d_dofmap, c->d_Gq,
                                                                  coeff, coeff2, cb, ce, D);
  else
    stiffness_quad_kernel<N, false><<<blocks, Q::THREADS, 0, st>>>(x, x2, y, c->d_dofmap, c->d_Gq,
                                                                   coeff, coeff2, cb, ce, D);
  FUS_LAUNCHED();
  return FUS_OK;
}

int launch_stiffness(fus_ctx* c, const double* x, const double* x2, const double* coeff,
                     const double* coeff2, double* y, long long cb, long long ce,
                     cudaStream_t st, const FusedHalo* fh = nullptr) {
  if (c->dim == 2) {
    if (!c->d_Gq) {
      set_error("context was created without G: stiffness operator unavailable");
      return FUS_ERR_STATE;
    }
    switch (c->N) {
    case 2: return launch_stiffness_quad_n<2>(c, x, x2, coeff, coeff2, y, cb, ce, st);
    case 3: return launch_stiffness_quad_n<3>(c, x, x2, coeff, coeff2, y, cb, ce, st);
    case 4: return launch_stiffness_quad_n<4>(c, x, x2, coeff, coeff2, y, cb, ce, st);
    case 5: return launch_stiffness_quad_n<5>(c, x, x2, coeff, coeff2, y, cb, ce, st);
    case 6: return launch_stiffness_quad_n<6>(c, x, x2, coeff, coeff2, y, cb, ce, st);
    case 7: return launch_stiffness_quad_n<7>(c, x, x2, coeff, coeff2, y, cb, ce, st);
    case 8: return launch_stiffness_quad_n<8>(c, x, x2, coeff, coeff2, y, cb, ce, st);
    }
    set_error("unsupported degree P=%d", c->P);
    return FUS_ERR_UNSUPPORTED;
  }
  if (!c->d_G2 && c->geom_active < 2) {
    set_error("context was created without G: stiffness operator unavailable");
    return FUS_ERR_STATE;
  }
  switch (c->N) {
  case 2: return launch_stiffness_n<2>(c, x, x2, coeff, coeff2, y, cb, ce, st, fh);
  case 3: return launch_stiffness_n<3>(c, x, x2, coeff, coeff2, y, cb, ce, st, fh);
  case 4: return launch_stiffness_n<4>(c, x, x2, coeff, coeff2, y, cb, ce, st, fh);
  case 5: return launch_stiffness_n<5>(c, x, x2, coeff, coeff2, y, cb, ce, st, fh);
  case 6: return launch_stiffness_n<6>(c, x, x2, coeff, coeff2, y, cb, ce, st, fh);
  case 7: return launch_stiffness_n<7>(c, x, x2, coeff, coeff2, y, cb, ce, st, fh);
  case 8: return launch_stiffness_n<8>(c, x, x2, coeff, coeff2, y, cb, ce, st, fh);
  }
  set_error("unsupported degree P=%d", c->P);
  return FUS_ERR_UNSUPPORTED;
}

// ---- FP32 operators (float copies of the cell data; 3-D, streamed geometry only) ------------------
int ensure_f32(fus_ctx* c, bool want_G, bool want_detJ) {
  if (c->dim != 3 || c->lean) {
    set_error("the FP32 operators need a hexahedral context that stores G / detJ");
    return FUS_ERR_UNSUPPORTED;
  }
  const long long nent = c->ncells * c->Nd;
  if (want_G && !c->d_G2f) {
    if (!c->d_G2) {
      set_error("context was created without G: stiffness operator unavailable");
      return FUS_ERR_STATE;
    }
    FUS_CUDA(cudaMalloc(&c->d_G2f, sizeof(float) * 6 * nent));
    narrow_kernel<<<grid_for(6 * nent, 256, c->num_sms * 8), 256, 0, c->stream>>>(
        reinterpret_cast<const double*>(c->d_G2), c->d_G2f, 6 * nent); // same [cell][i0][p][t] order
    FUS_LAUNCHED();
  }
  if (want_detJ && !c->d_detJf) {
    if (!c->d_detJ) {
      set_error("context was created without detJ: mass operator unavailable");
      return FUS_ERR_STATE;
    }
    FUS_CUDA(cudaMalloc(&c->d_detJf, sizeof(float) * nent));
    narrow_kernel<<<grid_for(nent, 256, c->num_sms * 8), 256, 0, c->stream>>>(c->d_detJ,
                                                                              c->d_detJf, nent);
    FUS_LAUNCHED();
  }
  return FUS_OK;
}

template <int N>
int launch_stiffness_f32_n(fus_ctx* c, const float* x, const float* coeff, float* y) {
  using L = LineCfg<N>;
  static KernelCfg cfg;
  DMatT<float, N> D;
  for (int i = 0; i < N * N; ++i)
    D.d[i] = (float)c->dphi[i];
  for (int i = 0; i < N; ++i) {
    D.w[i] = (float)c->wts[i];
    D.x[i] = (float)c->pts[i];
  }
  auto kern = stiffness_line_kernel<N, false, 0, float>;
  if (!cfg.configured[c->device].load(std::memory_order_acquire)) {
    std::lock_guard<std::mutex> lock(cfg.mu);
    if (!cfg.configured[c->device].load(std::memory_order_relaxed)) {
      FUS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    L::SMEM_BYTES));
      FUS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cfg.blocks_plain[c->device], kern,
                                                            L::THREADS, L::SMEM_BYTES));
      if (cfg.blocks_plain[c->device] < 1) {
        set_error("FP32 stiffness kernel <N=%d> does not fit on an SM", N);
        return FUS_ERR_CUDA;
      }
      cfg.configured[c->device].store(true, std::memory_order_release);
    }
  }
  ProfScope prof(c, 0, c->stream);
  const long long want = (c->ncells + L::CPB - 1) / L::CPB;
  const int blocks
      = (int)std::min<long long>(want, (long long)c->num_sms * cfg.blocks_plain[c->device]);
  kern<<<blocks, L::THREADS, L::SMEM_BYTES, c->stream>>>(
      x, nullptr, y, c->d_dofmap, reinterpret_cast<const float2*>(c->d_G2f), coeff, nullptr, 0,
      c->ncells, D, HaloLaunch{});
  FUS_LAUNCHED();
  return FUS_OK;
}

template <int N>
int launch_mass_tri_n(fus_ctx* c, const double* x, const double* coeff, double* y, long long cb,
                      long long ce, cudaStream_t st) {
  Rule1D<N> R;
  std::memcpy(R.pts, c->pts, sizeof(double) * N);
  std::memcpy(R.wts, c->wts, sizeof(double) * N);
  const long long np = (ce - cb) * c->Nd;
  mass_tri_kernel<N><<<grid_for(np, 256, c->num_sms * 8), 256, 0, st>>>(x, y, c->d_dofmap, c->d_tri,
                                                                        coeff, cb, np, R);
  FUS_LAUNCHED();
  return FUS_OK;
}

int launch_mass(fus_ctx* c, const double* x, const double* coeff, double* y, long long cb,
                long long ce, cudaStream_t st) {
  if (c->lean && c->d_tri) { // no detJ array: |det J| w from the trilinear cell map
    if (ce <= cb)
      return FUS_OK;
    switch (c->N) {
    case 2: return launch_mass_tri_n<2>(c, x, coeff, y, cb, ce, st);
    case 3: return launch_mass_tri_n<3>(c, x, coeff, y, cb, ce, st);
    case 4: return launch_mass_tri_n<4>(c, x, coeff, y, cb, ce, st);
    case 5: return launch_mass_tri_n<5>(c, x, coeff, y, cb, ce, st);
    case 6: return launch_mass_tri_n<6>(c, x, coeff, y, cb, ce, st);
    case 7: return launch_mass_tri_n<7>(c, x, coeff, y, cb, ce, st);
    case 8: return launch_mass_tri_n<8>(c, x, coeff, y, cb, ce, st);
    }
    set_error("unsupported degree P=%d", c->P);
    return FUS_ERR_UNSUPPORTED;
  }
  if (!c->d_detJ) {
    set_error("context was created without detJ: mass operator unavailable");
    return FUS_ERR_STATE;
  }
  if (ce <= cb)
    return FUS_OK;
  // the kernel indexes points from 0: shift the per-point arrays, keep coeff indexed by cell
  const long long np = (ce - cb) * c->Nd;
  mass_kernel<<<grid_for(np, 256, c->num_sms * 8), 256, 0, st>>>(
      x, y, c->d_dofmap + cb * c->Nd, c->d_detJ + cb * c->Nd, coeff + cb, np, c->Nd);
  FUS_LAUNCHED();
  return FUS_OK;
}

template <int N>
int launch_geometry_n(fus_ctx* c, const double* d_xg, const int32_t* d_xd, bool want_G,
                      bool want_detJ) {
  Rule1D<N> R;
  FUS_TRY(gll(N - 1, R.pts, R.wts));
  geometry_kernel<N><<<grid_for(c->ncells * c->Nd, 128, c->num_sms * 16), 128, 0, c->stream>>>(
      d_xg, d_xd, c->ncells, want_G ? c->d_G2 : nullptr, want_detJ ? c->d_detJ : nullptr, R);
  FUS_LAUNCHED();
  return FUS_OK;
}

template <int N>
int launch_geometry_quad_n(fus_ctx* c, const double* d_xg, const int32_t* d_xd) {
  Rule1D<N> R;
  FUS_TRY(gll(N - 1, R.pts, R.wts));
  geometry_quad_kernel<N><<<grid_for(c->ncells * c->Nd, 128, c->num_sms * 16), 128, 0,
                            c->stream>>>(d_xg, d_xd, c->ncells, c->d_Gq, c->d_detJ, R);
  FUS_LAUNCHED();
  return FUS_OK;
}

template <int N>
int g_upload_n(fus_ctx* c, const double* G) {
  // stream the reference-layout G through a bounded staging buffer
  const long long chunk_cells = std::max<long long>(1, (64ll << 20) / (c->Nd * 48));
  DevPtr<double> stage;
  FUS_CUDA(cudaMalloc(&stage.p, (size_t)chunk_cells * c->Nd * 48));
  for (long long c0 = 0; c0 < c->ncells; c0 += chunk_cells) {
    const long long nc = std::min<long long>(chunk_cells, c->ncells - c0);
    FUS_CUDA(cudaMemcpyAsync(stage.p, G + c0 * c->Nd * 6, (size_t)nc * c->Nd * 48,
                             cudaMemcpyHostToDevice, c->stream));
    g_to_device_layout_kernel<N><<<grid_for(nc * c->Nd, 256, c->num_sms * 8), 256, 0,
                                   c->stream>>>(stage.p, nc, c->d_G2 + c0 * (3 * c->Nd));
    FUS_LAUNCHED();
    // the staging buffer is reused by the next chunk
    FUS_CUDA(cudaStreamSynchronize(c->stream));
  }
  return FUS_OK;
}

template <int N>
int g_download_n(fus_ctx* c, double* G) {
  const long long chunk_cells = std::max<long long>(1, (64ll << 20) / (c->Nd * 48));
  DevPtr<double> stage;
  FUS_CUDA(cudaMalloc(&stage.p, (size_t)chunk_cells * c->Nd * 48));
  for (long long c0 = 0; c0 < c->ncells; c0 += chunk_cells) {
    const long long nc = std::min<long long>(chunk_cells, c->ncells - c0);
    g_from_device_layout_kernel<N><<<grid_for(nc * c->Nd, 256, c->num_sms * 8), 256, 0,
                                     c->stream>>>(c->d_G2 + c0 * (3 * c->Nd), nc, stage.p);
    FUS_LAUNCHED();
    FUS_CUDA(cudaMemcpyAsync(G + c0 * c->Nd * 6, stage.p, (size_t)nc * c->Nd * 48,
                             cudaMemcpyDeviceToHost, c->stream));
    FUS_CUDA(cudaStreamSynchronize(c->stream));
  }
  return FUS_OK;
}

template <int N>
int affine_detect_n(fus_ctx* c, int* all_affine) {
  DMat<N> D;
  std::memcpy(D.d, c->dphi, sizeof(double) * N * N);
  std::memcpy(D.w, c->wts, sizeof(double) * N);
  std::memcpy(D.x, c->pts, sizeof(double) * N);
  DevPtr<int> flag;
  FUS_CUDA(cudaMalloc(&flag.p, sizeof(int)));
  int* d_flag = flag.p;
  const int one = 1;
  FUS_CUDA(cudaMemcpyAsync(d_flag, &one, sizeof(int), cudaMemcpyHostToDevice, c->stream));
  if (!c->d_Ghat)
    FUS_CUDA(cudaMalloc(&c->d_Ghat, sizeof(double2) * 3 * c->ncells));
  affine_detect_kernel<N><<<grid_for(c->ncells, 128, 1 << 30), 128, 0, c->stream>>>(
      c->d_G2, c->ncells, 1e-13, c->d_Ghat, d_flag, D);
  FUS_LAUNCHED();
  FUS_CUDA(cudaMemcpyAsync(all_affine, d_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  FUS_CUDA(cudaStreamSynchronize(c->stream));
  return FUS_OK;
}

#define FUS_DISPATCH_N(c, fn, ...)                                                                 \
  [&]() -> int {                                                                                   \
    switch ((c)->N) {                                                                              \
    case 2: return fn<2>(__VA_ARGS__);                                                             \
    case 3: return fn<3>(__VA_ARGS__);                                                             \
    case 4: return fn<4>(__VA_ARGS__);                                                             \
    case 5: return fn<5>(__VA_ARGS__);                                                             \
    case 6: return fn<6>(__VA_ARGS__);                                                             \
    case 7: return fn<7>(__VA_ARGS__);                                                             \
    case 8: return fn<8>(__VA_ARGS__);                                                             \
    }                                                                                              \
    set_error("unsupported degree P=%d", (c)->P);                                                  \
    return FUS_ERR_UNSUPPORTED;                                                                    \
  }()

int ctx_common(int P, int64_t ncells, int64_t ndofs, int64_t nowned, const int32_t* dm,
               int device, bool want_G, bool want_detJ, fus_ctx** out, int dim = 3) {
  if (!out || !dm || ncells < 1 || ndofs < 1 || nowned < 0 || nowned > ndofs) {
    set_error("fus_ctx_create: bad argument");
    return FUS_ERR_ARG;
  }
  if (P < 1 || P > 7) {
    set_error("unsupported degree P=%d (supported: 1..7)", P);
    return FUS_ERR_UNSUPPORTED;
  }
  if (ndofs > INT32_MAX) {
    set_error("local dof count exceeds int32 (the reference dofmap is int32 too)");
    return FUS_ERR_UNSUPPORTED;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1 || device < 0 || device >= ndev
      || device >= kMaxDevices) {
    set_error("no usable CUDA device (requested %d of %d); there is no CPU fallback", device, ndev);
    return FUS_ERR_CUDA;
  }
  cudaDeviceProp prop;
  FUS_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10 || prop.minor != 0) { // arch-specific SASS only: no other target can load it
    set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
              prop.minor);
    return FUS_ERR_CUDA;
  }
  fus_ctx* c = new fus_ctx();
  c->dim = dim;
  c->P = P;
  c->N = P + 1;
  c->Nd = (dim == 3) ? c->N * c->N * c->N : c->N * c->N;
  c->ncells = ncells;
  c->ndofs = ndofs;
  c->nowned = nowned;
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  gll(P, c->pts, c->wts);
  if (const char* e = std::getenv("FUS_STIFFNESS_VARIANT")) { // A/B runs of bench.py
    const int v = std::atoi(e);
    if (v >= -1 && v <= 6)
      c->variant = v;
  }
  if (const char* e = std::getenv("FUS_L2_PERSIST"))
    c->l2_persist = std::atoi(e) != 0;
  if (const char* e = std::getenv("FUS_REVERSE_OPERATOR"))
    c->reverse_op = std::atoi(e) != 0;
  if (const char* e = std::getenv("FUS_STAGE_HINTS"))
    c->stage_hints = std::atoi(e) != 0;
  if (const char* e = std::getenv("FUS_USE_GRAPH"))
    c->use_graph = std::atoi(e) != 0;
  *out = c;
  FUS_CUDA(cudaSetDevice(device));
  FUS_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  c->own_stream = true;
  // validate the dofmap on the host: every index inside [0, ndofs)
  const int64_t nent = ncells * c->Nd;
  for (int64_t i = 0; i < nent; ++i)
    if (dm[i] < 0 || dm[i] >= ndofs) {
      set_error("tensor_dofmap[%lld] = %d outside [0,%lld)", (long long)i, dm[i],
                (long long)ndofs);
      return FUS_ERR_ARG;
    }
  FUS_CUDA(cudaMalloc(&c->d_dofmap, sizeof(int32_t) * nent));
  FUS_CUDA(cudaMemcpyAsync(c->d_dofmap, dm, sizeof(int32_t) * nent, cudaMemcpyHostToDevice,
                           c->stream));
  if (want_G && dim == 3)
    FUS_CUDA(cudaMalloc(&c->d_G2, sizeof(double2) * 3 * nent));
  if (want_G && dim == 2)
    FUS_CUDA(cudaMalloc(&c->d_Gq, sizeof(double) * 3 * nent));
  if (want_detJ)
    FUS_CUDA(cudaMalloc(&c->d_detJ, sizeof(double) * nent));
  return FUS_OK;
}

} // namespace

extern "C" {

const char* fus_last_error(void) { return g_err.c_str(); }
int fus_version(void) { return 100; }
int64_t fus_launch_count(void) { return g_launches.load(); }

int fus_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int ok = 0;
  for (int d = 0; d < n; ++d) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, d) == cudaSuccess && p.major == 10 && p.minor == 0)
      ++ok;
  }
  return ok;
}

// ---- host setup ---------------------------------------------------------------------------------
int fus_gll(int P, double* pts, double* wts) { return gll(P, pts, wts); }
int fus_tabulate_dphi(int P, double* dphi) { return tabulate_dphi(P, dphi); }
int fus_box_mesh(const int n[3], const double lo[3], const double hi[3], double* xg,
                 int32_t* xdofmap) {
  return box_mesh(n, lo, hi, xg, xdofmap);
}
int fus_box_dofmap(int P, const int n[3], int numbering, int32_t* dm) {
  return box_dofmap(P, n, numbering, dm);
}
int64_t fus_box_num_dofs(int P, const int n[3]) { return box_num_dofs(P, n); }
int64_t fus_box_facets(const int n[3], int32_t* facets) { return box_facets(n, facets); }
int fus_boundary_vectors(int kind, int P, int64_t ncells, int64_t ndofs, const double* xg,
                         const int32_t* xdofmap, const int32_t* tensor_dofmap, int64_t nfacets,
                         const int32_t* facets, const double* c0, const double* rho0,
                         const double* delta0, double* src, double* dsrc, double* absb,
                         double* bmass) {
  return boundary_vectors(kind, P, ncells, ndofs, xg, xdofmap, tensor_dofmap, nfacets, facets, c0,
                          rho0, delta0, src, dsrc, absb, bmass);
}
int fus_trilinear_coeffs(int64_t ncells, const double* xg, const int32_t* xdofmap,
                         double* coeffs) {
  return trilinear_coeffs(ncells, xg, xdofmap, coeffs);
}
int fus_trilinear_geometry(int P, int64_t ncells, const double* coeffs, double* G, double* detJ) {
  return trilinear_geometry(P, ncells, coeffs, G, detJ);
}
int fus_rect_mesh(const int n[2], const double lo[2], const double hi[2], double* xg,
                  int32_t* xdofmap) {
  return rect_mesh(n, lo, hi, xg, xdofmap);
}
int fus_rect_dofmap(int P, const int n[2], int32_t* dm) { return rect_dofmap(P, n, dm); }
int64_t fus_rect_num_dofs(int P, const int n[2]) { return rect_num_dofs(P, n); }
int64_t fus_rect_facets(const int n[2], int32_t* facets) { return rect_facets(n, facets); }
int fus_boundary_vectors_2d(int kind, int P, int64_t ncells, int64_t ndofs, const double* xg,
                            const int32_t* xdofmap, const int32_t* tensor_dofmap, int64_t nfacets,
                            const int32_t* facets, const double* c0, const double* rho0,
                            const double* delta0, double* src, double* dsrc, double* absb,
                            double* bmass) {
  return boundary_vectors_2d(kind, P, ncells, ndofs, xg, xdofmap, tensor_dofmap, nfacets, facets,
                             c0, rho0, delta0, src, dsrc, absb, bmass);
}

// ---- context ----------------------------------------------------------------------------------
// a context that failed half-way through its construction is released, never handed back
static int ctx_fail(fus_ctx** out, int rc) {
  if (rc != FUS_OK && out && *out) {
    fus_ctx_destroy(*out);
    *out = nullptr;
  }
  return rc;
}

int fus_ctx_create(int P, int64_t ncells, int64_t ndofs, int64_t nowned,
                   const int32_t* tensor_dofmap, const double* G, const double* detJ,
                   const double* dphi, int device, fus_ctx** out) {
  if (out)
    *out = nullptr;
  if (!dphi || (!G && !detJ)) {
    set_error("fus_ctx_create: dphi and at least one of G, detJ are required");
    return FUS_ERR_ARG;
  }
  int r = ctx_common(P, ncells, ndofs, nowned, tensor_dofmap, device, G != nullptr,
                     detJ != nullptr, out);
  if (r != FUS_OK)
    return ctx_fail(out, r);
  fus_ctx* c = *out;
  std::memcpy(c->dphi, dphi, sizeof(double) * c->N * c->N);
  auto fill = [&]() -> int {
    if (detJ)
      FUS_CUDA(cudaMemcpyAsync(c->d_detJ, detJ, sizeof(double) * ncells * c->Nd,
                               cudaMemcpyHostToDevice, c->stream));
    if (G)
      FUS_TRY(FUS_DISPATCH_N(c, g_upload_n, c, G));
    FUS_CUDA(cudaStreamSynchronize(c->stream));
    return FUS_OK;
  };
  return ctx_fail(out, fill());
}

static int ctx_from_mesh(int P, int64_t ncells, int64_t ndofs, int64_t nowned,
                         const int32_t* tensor_dofmap, int64_t nverts, const double* xg,
                         const int32_t* xdofmap, int device, bool lean, fus_ctx** out) {
  if (out)
    *out = nullptr;
  if (!xg || !xdofmap || nverts < 8) {
    set_error("fus_ctx_create_from_mesh: mesh geometry required");
    return FUS_ERR_ARG;
  }
  for (int64_t i = 0; i < ncells * 8; ++i)
    if (xdofmap[i] < 0 || xdofmap[i] >= nverts) {
      set_error("xdofmap[%lld] = %d outside [0,%lld)", (long long)i, xdofmap[i],
                (long long)nverts);
      return FUS_ERR_ARG;
    }
  int mode_env = 0;
  if (const char* e = std::getenv("FUS_GEOMETRY_MODE")) { // "2": rebuild G on the fly, "lean": and
    if (!std::strcmp(e, "lean"))                          // do not even store G / detJ
      lean = true;
    else if (!std::strcmp(e, "2"))
      mode_env = 2;
  }
  int r = ctx_common(P, ncells, ndofs, nowned, tensor_dofmap, device, !lean, !lean, out);
  if (r != FUS_OK)
    return ctx_fail(out, r);
  fus_ctx* c = *out;
  c->lean = lean;
  auto fill = [&]() -> int {
    FUS_TRY(tabulate_dphi(P, c->dphi));
    DevPtr<double> d_xg;
    DevPtr<int32_t> d_xd;
    FUS_CUDA(cudaMalloc(&d_xg.p, sizeof(double) * 3 * nverts));
    FUS_CUDA(cudaMalloc(&d_xd.p, sizeof(int32_t) * 8 * ncells));
    FUS_CUDA(cudaMemcpyAsync(d_xg.p, xg, sizeof(double) * 3 * nverts, cudaMemcpyHostToDevice,
                             c->stream));
    FUS_CUDA(cudaMemcpyAsync(d_xd.p, xdofmap, sizeof(int32_t) * 8 * ncells,
                             cudaMemcpyHostToDevice, c->stream));
    if (!lean)
      FUS_TRY(FUS_DISPATCH_N(c, launch_geometry_n, c, d_xg.p, d_xd.p, true, true));
    // the trilinear map itself, 192 B per cell: what option geometry_mode = 2 reads instead of G
    FUS_CUDA(cudaMalloc(&c->d_tri, sizeof(double) * FUS_TRI_STRIDE * ncells));
    tri_coeff_kernel<<<grid_for(ncells, 128, 1 << 30), 128, 0, c->stream>>>(d_xg.p, d_xd.p, ncells,
                                                                          c->d_tri);
    FUS_LAUNCHED();
    FUS_CUDA(cudaStreamSynchronize(c->stream));
    if (lean || mode_env == 2)
      c->geom_active = 2;
    return FUS_OK;
  };
  return ctx_fail(out, fill());
}

int fus_ctx_create_from_mesh(int P, int64_t ncells, int64_t ndofs, int64_t nowned,
                             const int32_t* tensor_dofmap, int64_t nverts, const double* xg,
                             const int32_t* xdofmap, int device, fus_ctx** out) {
  return ctx_from_mesh(P, ncells, ndofs, nowned, tensor_dofmap, nverts, xg, xdofmap, device, false,
                       out);
}

int fus_ctx_create_from_mesh_lean(int P, int64_t ncells, int64_t ndofs, int64_t nowned,
                                  const int32_t* tensor_dofmap, int64_t nverts, const double* xg,
                                  const int32_t* xdofmap, int device, fus_ctx** out) {
  return ctx_from_mesh(P, ncells, ndofs, nowned, tensor_dofmap, nverts, xg, xdofmap, device, true,
                       out);
}

// ---- 2-D quadrilateral contexts (cpp/fenicsx-sf-naive/common/spectral_op.hpp:28-107,226-359) -----
int fus_ctx_create_2d(int P, int64_t ncells, int64_t ndofs, int64_t nowned,
                      const int32_t* tensor_dofmap, const double* G, const double* detJ,
                      const double* dphi, int device, fus_ctx** out) {
  if (out)
    *out = nullptr;
  if (!dphi || (!G && !detJ)) {
    set_error("fus_ctx_create_2d: dphi and at least one of G, detJ are required");
    return FUS_ERR_ARG;
  }
  int r = ctx_common(P, ncells, ndofs, nowned, tensor_dofmap, device, G != nullptr,
                     detJ != nullptr, out, 2);
  if (r != FUS_OK)
    return ctx_fail(out, r);
  fus_ctx* c = *out;
  std::memcpy(c->dphi, dphi, sizeof(double) * c->N * c->N);
  auto fill = [&]() -> int {
    const int64_t nent = ncells * c->Nd;
    if (detJ)
      FUS_CUDA(cudaMemcpyAsync(c->d_detJ, detJ, sizeof(double) * nent, cudaMemcpyHostToDevice,
                               c->stream));
    if (G) { // reference layout G[c][q][3] -> Gq[c][p][q]; 2-D data are small: transposed on the host
      std::vector<double> tmp((size_t)3 * nent);
      for (int64_t cell = 0; cell < ncells; ++cell)
        for (int q = 0; q < c->Nd; ++q)
          for (int p = 0; p < 3; ++p)
            tmp[(size_t)(cell * 3 + p) * c->Nd + q] = G[(size_t)(cell * c->Nd + q) * 3 + p];
      FUS_CUDA(cudaMemcpyAsync(c->d_Gq, tmp.data(), sizeof(double) * tmp.size(),
                               cudaMemcpyHostToDevice, c->stream));
      FUS_CUDA(cudaStreamSynchronize(c->stream));
    }
    FUS_CUDA(cudaStreamSynchronize(c->stream));
    return FUS_OK;
  };
  return ctx_fail(out, fill());
}

int fus_ctx_create_from_mesh_2d(int P, int64_t ncells, int64_t ndofs, int64_t nowned,
                                const int32_t* tensor_dofmap, int64_t nverts, const double* xg,
                                const int32_t* xdofmap, int device, fus_ctx** out) {
  if (out)
    *out = nullptr;
  if (!xg || !xdofmap || nverts < 4) {
    set_error("fus_ctx_create_from_mesh_2d: mesh geometry required");
    return FUS_ERR_ARG;
  }
  for (int64_t i = 0; i < ncells * 4; ++i)
    if (xdofmap[i] < 0 || xdofmap[i] >= nverts) {
      set_error("xdofmap[%lld] = %d outside [0,%lld)", (long long)i, xdofmap[i],
                (long long)nverts);
      return FUS_ERR_ARG;
    }
  int r = ctx_common(P, ncells, ndofs, nowned, tensor_dofmap, device, true, true, out, 2);
  if (r != FUS_OK)
    return ctx_fail(out, r);
  fus_ctx* c = *out;
  auto fill = [&]() -> int {
    FUS_TRY(tabulate_dphi(P, c->dphi));
    DevPtr<double> d_xg;
    DevPtr<int32_t> d_xd;
    FUS_CUDA(cudaMalloc(&d_xg.p, sizeof(double) * 3 * nverts));
    FUS_CUDA(cudaMalloc(&d_xd.p, sizeof(int32_t) * 4 * ncells));
    FUS_CUDA(cudaMemcpyAsync(d_xg.p, xg, sizeof(double) * 3 * nverts, cudaMemcpyHostToDevice,
                             c->stream));
    FUS_CUDA(cudaMemcpyAsync(d_xd.p, xdofmap, sizeof(int32_t) * 4 * ncells,
                             cudaMemcpyHostToDevice, c->stream));
    FUS_TRY(FUS_DISPATCH_N(c, launch_geometry_quad_n, c, d_xg.p, d_xd.p));
    FUS_CUDA(cudaStreamSynchronize(c->stream));
    return FUS_OK;
  };
  return ctx_fail(out, fill());
}

int fus_ctx_destroy(fus_ctx* c) {
  if (!c)
    return FUS_OK;
  if (c->live_models > 0) { // their device vectors live on this context's device and stream
    set_error("fus_ctx_destroy: %d model(s) still use this context; destroy them first",
              c->live_models);
    return FUS_ERR_STATE;
  }
  cudaSetDevice(c->device);
  if (c->stream)
    cudaStreamSynchronize(c->stream);
  if (c->halo)
    halo_destroy(c->halo);
  for (auto& fam : c->prof_events)
    for (auto& pr : fam) {
      cudaEventDestroy(pr.first);
      cudaEventDestroy(pr.second);
    }
  cudaFree(c->d_dofmap);
  cudaFree(c->d_Ghat);
  cudaFree(c->d_tri);
  cudaFree(c->d_G2f);
  cudaFree(c->d_detJf);
  cudaFree(c->d_Gt);
  cudaFree(c->d_dmt);
  cudaFree(c->d_G2);
  cudaFree(c->d_Gq);
  cudaFree(c->d_detJ);
  if (c->own_stream && c->stream)
    cudaStreamDestroy(c->stream);
  delete c;
  return FUS_OK;
}

int fus_ctx_set_stream(fus_ctx* c, void* s) {
  if (!c)
    return FUS_ERR_ARG;
  FUS_TRY(select_device(c));
  FUS_CUDA(cudaStreamSynchronize(c->stream));
  if (c->own_stream && c->stream)
    FUS_CUDA(cudaStreamDestroy(c->stream));
  c->stream = (cudaStream_t)s;
  c->own_stream = false;
  ++c->config_epoch;
  return FUS_OK;
}

int fus_ctx_set_option(fus_ctx* c, const char* name, int value) {
  if (!c || !name)
    return FUS_ERR_ARG;
  for (const char* k : {"stiffness_variant", "geometry_mode", "col_blocks_per_sm", "halo_overlap",
                        "halo_reserve_sms"})
    if (!std::strcmp(name, k))
      ++c->config_epoch; // a captured step graph would replay the previous choice
  if (!std::strcmp(name, "stiffness_variant")) {
    if (value < -1 || value > 7) {
      set_error("stiffness_variant must be -1 (auto) or 0..7");
      return FUS_ERR_ARG;
    }
    c->variant = value;
    return FUS_OK;
  }
  if (!std::strcmp(name, "profile_kernels")) {
    c->profile = value != 0;
    if (value)
      for (int f = 0; f < 3; ++f)
        c->prof_used[f] = 0;
    return FUS_OK;
  }
  if (!std::strcmp(name, "col_blocks_per_sm")) {
    c->col_blocks_per_sm = value;
    return FUS_OK;
  }
  if (!std::strcmp(name, "geometry_mode")) {
    // 0: stream G per point (default).  1: if EVERY cell is affine, keep one Ghat per cell and
    // rebuild G = w_q * Ghat in the kernel; otherwise stay on the streamed path.  2: rebuild G per
    // point from the trilinear cell map (any mesh; needs a context created from the mesh).
    if (c->dim == 2 && value != 0) {
      set_error("geometry_mode applies to hexahedral contexts only");
      return FUS_ERR_UNSUPPORTED;
    }
    if (c->lean && value != 2 && value != 3) {
      set_error("a lean context holds no G: geometry_mode is fixed at 2");
      return FUS_ERR_STATE;
    }
    if (value == 0) {
      c->geom_active = 0;
      return FUS_OK;
    }
    if (value == 2 || value == 3) { // 3: mode 2 under a 128-register cap (occupancy experiment)
      if (!c->d_tri) {
        set_error("geometry_mode 2 needs the cell vertices: create the context with "
                  "fus_ctx_create_from_mesh");
        return FUS_ERR_STATE;
      }
      c->geom_active = value;
      return FUS_OK;
    }
    if (value != 1) {
      set_error("geometry_mode must be 0, 1, 2 (or 3, the occupancy experiment of mode 2)");
      return FUS_ERR_ARG;
    }
    if (!c->d_G2) {
      set_error("geometry_mode 1 needs the stored G of the context");
      return FUS_ERR_STATE;
    }
    FUS_TRY(select_device(c));
    int all_affine = 0;
    FUS_TRY(FUS_DISPATCH_N(c, affine_detect_n, c, &all_affine));
    c->geom_active = all_affine ? 1 : 0;
    return FUS_OK;
  }
  if (!std::strcmp(name, "l2_persist")) {
    c->l2_persist = value != 0;
    return FUS_OK;
  }
  if (!std::strcmp(name, "reverse_operator")) {
    c->reverse_op = value != 0;
    ++c->config_epoch;
    return FUS_OK;
  }
  if (!std::strcmp(name, "stage_hints")) {
    c->stage_hints = value != 0;
    ++c->config_epoch;
    return FUS_OK;
  }
  if (!std::strcmp(name, "use_graph")) {
    c->use_graph = value != 0;
    return FUS_OK;
  }
  if (!std::strcmp(name, "halo_reserve_sms")) {
    if (value < 0 || value >= c->num_sms)
      return FUS_ERR_ARG;
    c->halo_reserve = value;
    return FUS_OK;
  }
  if (!std::strcmp(name, "halo_overlap") && c->halo) {
    halo_set_overlap(c->halo, value);
    return FUS_OK;
  }
  set_error("unknown option %s", name);
  return FUS_ERR_ARG;
}

static int check_peer_error(fus_ctx* c) {
  if (c->halo && halo_peer_error(c->halo)) {
    set_error("halo exchange timed out waiting for a neighbour (peer transport)");
    return FUS_ERR_COMM;
  }
  return FUS_OK;
}

int fus_ctx_get_option(fus_ctx* c, const char* name, int* value) {
  if (!c || !name || !value)
    return FUS_ERR_ARG;
  if (!std::strcmp(name, "geometry_compressed"))
    *value = c->geom_active;
  else if (!std::strcmp(name, "stiffness_variant"))
    *value = c->variant;
  else if (!std::strcmp(name, "halo_mode"))
    *value = c->halo ? halo_mode(c->halo) : -1;
  else {
    set_error("unknown option %s", name);
    return FUS_ERR_ARG;
  }
  return FUS_OK;
}

int fus_ctx_sync(fus_ctx* c) {
  if (!c)
    return FUS_ERR_ARG;
  FUS_TRY(select_device(c));
  FUS_CUDA(cudaStreamSynchronize(c->stream));
  return check_peer_error(c);
}

int fus_ctx_profile(fus_ctx* c, const char* kernel, int64_t* launches, double* total_ms) {
  if (!c || !kernel || !launches || !total_ms)
    return FUS_ERR_ARG;
  int fam = -1;
  if (!std::strcmp(kernel, "stiffness"))
    fam = 0;
  else if (!std::strcmp(kernel, "stage"))
    fam = 1;
  else if (!std::strcmp(kernel, "boundary"))
    fam = 2;
  if (fam < 0) {
    set_error("unknown kernel family %s", kernel);
    return FUS_ERR_ARG;
  }
  FUS_TRY(select_device(c));
  FUS_CUDA(cudaStreamSynchronize(c->stream));
  double tot = 0.0;
  for (size_t i = 0; i < c->prof_used[fam]; ++i) {
    float ms = 0.f;
    FUS_CUDA(cudaEventElapsedTime(&ms, c->prof_events[fam][i].first, c->prof_events[fam][i].second));
    tot += ms;
  }
  *launches = (int64_t)c->prof_used[fam];
  *total_ms = tot;
  return FUS_OK;
}

int fus_ctx_get_geometry(fus_ctx* c, double* G, double* detJ) {
  if (!c)
    return FUS_ERR_ARG;
  FUS_TRY(select_device(c));
  if (c->dim == 2) { // Gq[c][p][q] -> reference layout G[c][q][3]
    const int64_t nent = c->ncells * c->Nd;
    if (G) {
      if (!c->d_Gq) {
        set_error("fus_ctx_get_geometry: the context holds no G");
        return FUS_ERR_STATE;
      }
      std::vector<double> tmp((size_t)3 * nent);
      FUS_CUDA(cudaMemcpyAsync(tmp.data(), c->d_Gq, sizeof(double) * tmp.size(),
                               cudaMemcpyDeviceToHost, c->stream));
      FUS_CUDA(cudaStreamSynchronize(c->stream));
      for (int64_t cell = 0; cell < c->ncells; ++cell)
        for (int q = 0; q < c->Nd; ++q)
          for (int p = 0; p < 3; ++p)
            G[(size_t)(cell * c->Nd + q) * 3 + p] = tmp[(size_t)(cell * 3 + p) * c->Nd + q];
    }
    if (detJ) {
      if (!c->d_detJ) {
        set_error("fus_ctx_get_geometry: the context holds no detJ");
        return FUS_ERR_STATE;
      }
      FUS_CUDA(cudaMemcpyAsync(detJ, c->d_detJ, sizeof(double) * nent, cudaMemcpyDeviceToHost,
                               c->stream));
      FUS_CUDA(cudaStreamSynchronize(c->stream));
    }
    return FUS_OK;
  }
  if (c->lean) { // nothing stored: rebuild on the host from the cell map (tests, inspection)
    std::vector<double> co((size_t)c->ncells * FUS_TRI_STRIDE);
    FUS_CUDA(cudaMemcpyAsync(co.data(), c->d_tri, sizeof(double) * co.size(),
                             cudaMemcpyDeviceToHost, c->stream));
    FUS_CUDA(cudaStreamSynchronize(c->stream));
    return trilinear_geometry(c->P, c->ncells, co.data(), G, detJ);
  }
  if (G) {
    if (!c->d_G2) {
      set_error("fus_ctx_get_geometry: the context holds no G");
      return FUS_ERR_STATE;
    }
    FUS_TRY(FUS_DISPATCH_N(c, g_download_n, c, G));
  }
  if (detJ) {
    if (!c->d_detJ) {
      set_error("fus_ctx_get_geometry: the context holds no detJ");
      return FUS_ERR_STATE;
    }
    FUS_CUDA(cudaMemcpyAsync(detJ, c->d_detJ, sizeof(double) * c->ncells * c->Nd,
                             cudaMemcpyDeviceToHost, c->stream));
    FUS_CUDA(cudaStreamSynchronize(c->stream));
  }
  return FUS_OK;
}

// ---- operators --------------------------------------------------------------------------------
int fus_stiffness_apply_dev(fus_ctx* c, const double* x, const double* coeffs, double* y) {
  if (!c || !x || !coeffs || !y)
    return FUS_ERR_ARG;
  FUS_TRY(select_device(c));
  return launch_stiffness(c, x, nullptr, coeffs, nullptr, y, 0, c->ncells, c->stream);
}

int fus_mass_apply_dev(fus_ctx* c, const double* x, const double* coeffs, double* y) {
  if (!c || !x || !coeffs || !y)
    return FUS_ERR_ARG;
  FUS_TRY(select_device(c));
  return launch_mass(c, x, coeffs, y, 0, c->ncells, c->stream);
}

static int apply_host(fus_ctx* c, const double* x, const double* coeffs, double* y, bool stiff) {
  if (!c || !x || !coeffs || !y)
    return FUS_ERR_ARG;
  FUS_TRY(select_device(c));
  DevPtr<double> dx, dy, dc;
  const size_t vb = sizeof(double) * c->ndofs, cbytes = sizeof(double) * c->ncells;
  FUS_CUDA(cudaMalloc(&dx.p, vb));
  FUS_CUDA(cudaMalloc(&dy.p, vb));
  FUS_CUDA(cudaMalloc(&dc.p, cbytes));
  FUS_CUDA(cudaMemcpyAsync(dx.p, x, vb, cudaMemcpyHostToDevice, c->stream));
  FUS_CUDA(cudaMemcpyAsync(dy.p, y, vb, cudaMemcpyHostToDevice, c->stream));
  FUS_CUDA(cudaMemcpyAsync(dc.p, coeffs, cbytes, cudaMemcpyHostToDevice, c->stream));
  FUS_TRY(stiff ? launch_stiffness(c, dx.p, nullptr, dc.p, nullptr, dy.p, 0, c->ncells, c->stream)
                : launch_mass(c, dx.p, dc.p, dy.p, 0, c->ncells, c->stream));
  FUS_CUDA(cudaMemcpyAsync(y, dy.p, vb, cudaMemcpyDeviceToHost, c->stream));
  FUS_CUDA(cudaStreamSynchronize(c->stream));
  return FUS_OK;
}

int fus_stiffness_apply_host(fus_ctx* c, const double* x, const double* coeffs, double* y) {
  return apply_host(c, x, coeffs, y, true);
}
int fus_mass_apply_host(fus_ctx* c, const double* x, const double* coeffs, double* y) {
  return apply_host(c, x, coeffs, y, false);
}

// ---- FP32 operator entry points -------------------------------------------------------------------
int fus_stiffness_apply_f32_dev(fus_ctx* c, const float* x, const float* coeffs, float* y) {
  if (!c || !x || !coeffs || !y)
    return FUS_ERR_ARG;
  FUS_TRY(select_device(c));
  FUS_TRY(ensure_f32(c, true, false));
  return FUS_DISPATCH_N(c, launch_stiffness_f32_n, c, x, coeffs, y);
}

int fus_mass_apply_f32_dev(fus_ctx* c, const float* x, const float* coeffs, float* y) {
  if (!c || !x || !coeffs || !y)
    return FUS_ERR_ARG;
  FUS_TRY(select_device(c));
  FUS_TRY(ensure_f32(c, false, true));
  const long long np = c->ncells * c->Nd;
  mass_kernel_f32<<<grid_for(np, 256, c->num_sms * 8), 256, 0, c->stream>>>(
      x, y, c->d_dofmap, c->d_detJf, coeffs, np, c->Nd);
  FUS_LAUNCHED();
  return FUS_OK;
}

static int apply_host_f32(fus_ctx* c, const float* x, const float* coeffs, float* y, bool stiff) {
  if (!c || !x || !coeffs || !y)
    return FUS_ERR_ARG;
  FUS_TRY(select_device(c));
  DevPtr<float> dx, dy, dc;
  const size_t vb = sizeof(float) * c->ndofs, cbytes = sizeof(float) * c->ncells;
  FUS_CUDA(cudaMalloc(&dx.p, vb));
  FUS_CUDA(cudaMalloc(&dy.p, vb));
  FUS_CUDA(cudaMalloc(&dc.p, cbytes));
  FUS_CUDA(cudaMemcpyAsync(dx.p, x, vb, cudaMemcpyHostToDevice, c->stream));
  FUS_CUDA(cudaMemcpyAsync(dy.p, y, vb, cudaMemcpyHostToDevice, c->stream));
  FUS_CUDA(cudaMemcpyAsync(dc.p, coeffs, cbytes, cudaMemcpyHostToDevice, c->stream));
  FUS_TRY(stiff ? fus_stiffness_apply_f32_dev(c, dx.p, dc.p, dy.p)
                : fus_mass_apply_f32_dev(c, dx.p, dc.p, dy.p));
  FUS_CUDA(cudaMemcpyAsync(y, dy.p, vb, cudaMemcpyDeviceToHost, c->stream));
  FUS_CUDA(cudaStreamSynchronize(c->stream));
  return FUS_OK;
}

int fus_stiffness_apply_f32_host(fus_ctx* c, const float* x, const float* coeffs, float* y) {
  return apply_host_f32(c, x, coeffs, y, true);
}
int fus_mass_apply_f32_host(fus_ctx* c, const float* x, const float* coeffs, float* y) {
  return apply_host_f32(c, x, coeffs, y, false);
}

// ---- device memory helpers ----------------------------------------------------------------------
int fus_dev_alloc(fus_ctx* c, size_t bytes, void** p) {
  if (!c || !p)
    return FUS_ERR_ARG;
  FUS_TRY(select_device(c));
  FUS_CUDA(cudaMalloc(p, bytes ? bytes : 8));
  return FUS_OK;
}
int fus_dev_free(fus_ctx* c, void* p) {
  if (!c)
    return FUS_ERR_ARG;
  FUS_TRY(select_device(c));
  FUS_CUDA(cudaStreamSynchronize(c->stream));
  FUS_CUDA(cudaFree(p));
  return FUS_OK;
}
int fus_dev_upload(fus_ctx* c, void* dst, const void* src, size_t bytes) {
  if (!c)
    return FUS_ERR_ARG;
  FUS_TRY(select_device(c));
  FUS_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, c->stream));
  FUS_CUDA(cudaStreamSynchronize(c->stream));
  return FUS_OK;
}
int fus_dev_download(fus_ctx* c, void* dst, const void* src, size_t bytes) {
  if (!c)
    return FUS_ERR_ARG;
  FUS_TRY(select_device(c));
  FUS_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream));
  FUS_CUDA(cudaStreamSynchronize(c->stream));
  return FUS_OK;
}
int fus_dev_memset(fus_ctx* c, void* dst, int value, size_t bytes) {
  if (!c)
    return FUS_ERR_ARG;
  FUS_TRY(select_device(c));
  FUS_CUDA(cudaMemsetAsync(dst, value, bytes, c->stream));
  return FUS_OK;
}

// ---- models -------------------------------------------------------------------------------------
static int model_alloc_vec(fus_ctx* c, double** p) {
  FUS_CUDA(cudaMalloc(p, sizeof(double) * c->ndofs));
  FUS_CUDA(cudaMemsetAsync(*p, 0, sizeof(double) * c->ndofs, c->stream));
  return FUS_OK;
}

int fus_model_create(fus_ctx* c, int kind, const double* c0, const double* rho0,
                     const double* delta0, const double* beta0, const double* src,
                     const double* dsrc, const double* absb, const double* bmass, double freq,
                     double p0, double s0, fus_model** out) {
  if (out)
    *out = nullptr;
  if (!c || !out || !c0 || !rho0 || kind < 0 || kind > 2) {
    set_error("fus_model_create: bad argument");
    return FUS_ERR_ARG;
  }
  if ((kind >= FUS_LOSSY && !delta0) || (kind == FUS_WESTERVELT && !beta0)) {
    set_error("fus_model_create: delta0/beta0 required for this model kind");
    return FUS_ERR_ARG;
  }
  if ((!(c->d_G2 || c->d_Gq) || !c->d_detJ) && !c->lean) {
    set_error("fus_model_create: context needs both G and detJ");
    return FUS_ERR_STATE;
  }
  FUS_TRY(select_device(c));
  fus_model* m = new fus_model();
  m->ctx = c;
  ++c->live_models;
  m->kind = kind;
  m->freq = freq;
  m->p0 = p0;
  m->s0 = s0;
  m->w0 = 2 * M_PI * freq;
  m->period = 1.0 / freq;
  m->window_length = 4.0;
  const int64_t nc = c->ncells, nd = c->ndofs;
  // operator coefficients (Linear.hpp:148-155, Lossy.hpp:166-169, Westervelt.hpp:182-187)
  std::vector<double> lin(nc), att(nc, 0.0), mco(nc), nl2(nc, 0.0);
  for (int64_t i = 0; i < nc; ++i) {
    lin[i] = -1.0 / rho0[i];
    if (kind >= FUS_LOSSY)
      att[i] = -delta0[i] / rho0[i] / c0[i] / c0[i];
    if (kind == FUS_WESTERVELT)
      nl2[i] = 2.0 * beta0[i] / rho0[i] / rho0[i] / c0[i] / c0[i] / c0[i] / c0[i];
    mco[i] = 1.0 / rho0[i] / c0[i] / c0[i];
  }
  const size_t cb = sizeof(double) * nc;
  auto build = [&]() -> int {
    DevPtr<double> mco_dev, nl2_dev; // set-up scratch, released on every exit path
    FUS_CUDA(cudaMalloc(&m->d_lin, cb));
    FUS_CUDA(cudaMalloc(&m->d_att, cb));
    FUS_CUDA(cudaMalloc(&mco_dev.p, cb));
    double*& d_mco = mco_dev.p;
    double*& d_nl2 = nl2_dev.p;
    FUS_CUDA(cudaMemcpyAsync(m->d_lin, lin.data(), cb, cudaMemcpyHostToDevice, c->stream));
    FUS_CUDA(cudaMemcpyAsync(m->d_att, att.data(), cb, cudaMemcpyHostToDevice, c->stream));
    FUS_CUDA(cudaMemcpyAsync(d_mco, mco.data(), cb, cudaMemcpyHostToDevice, c->stream));
    for (double** v : {&m->d_m, &m->d_u0, &m->d_v0, &m->d_ua, &m->d_va, &m->d_un, &m->d_vn, &m->d_b})
      FUS_TRY(model_alloc_vec(c, v));
    // lumped mass: the form `a` assembled with u == 1 (Linear.hpp:127-134), + facet mass term
    fill_kernel<<<grid_for(nd, 256, 1 << 30), 256, 0, c->stream>>>(m->d_un, 1.0, nd);
    FUS_LAUNCHED();
    FUS_TRY(launch_mass(c, m->d_un, d_mco, m->d_m, 0, nc, c->stream));
    if (bmass) {
      FUS_CUDA(cudaMemcpyAsync(m->d_vn, bmass, sizeof(double) * nd, cudaMemcpyHostToDevice,
                               c->stream));
      add_kernel<<<grid_for(nd, 256, 1 << 30), 256, 0, c->stream>>>(m->d_m, m->d_vn, nd);
      FUS_LAUNCHED();
    }
    if (kind == FUS_WESTERVELT) {
      FUS_CUDA(cudaMalloc(&d_nl2, cb));
      FUS_CUDA(cudaMemcpyAsync(d_nl2, nl2.data(), cb, cudaMemcpyHostToDevice, c->stream));
      FUS_TRY(model_alloc_vec(c, &m->d_dnl));
      FUS_TRY(launch_mass(c, m->d_un, d_nl2, m->d_dnl, 0, nc, c->stream));
    }
    if (c->halo) { // sum the per-rank partial sums on the owners, once (Linear.hpp:134)
      FUS_TRY(halo_reverse(c->halo, m->d_m, nullptr, c->stream));
      if (m->d_dnl)
        FUS_TRY(halo_reverse(c->halo, m->d_dnl, nullptr, c->stream));
    }
    FUS_CUDA(cudaStreamSynchronize(c->stream));
    FUS_CUDA(cudaMemsetAsync(m->d_un, 0, sizeof(double) * nd, c->stream));
    FUS_CUDA(cudaMemsetAsync(m->d_vn, 0, sizeof(double) * nd, c->stream));
    // Boundary terms (facet-lumped vectors of the collocated `ds` forms).  When the mesh is
    // partitioned every rank has integrated its own facets, so a shared dof carries partial sums:
    // they are added up on the owner once, here (the term is linear in them and v[d] is the same on
    // every rank), and the per-stage boundary update then touches owned dofs only.
    std::vector<double> hs(src ? src : nullptr, src ? src + nd : nullptr),
        hd(dsrc ? dsrc : nullptr, dsrc ? dsrc + nd : nullptr),
        ha(absb ? absb : nullptr, absb ? absb + nd : nullptr);
    if (c->halo) {
      for (std::vector<double>* hv : {&hs, &hd, &ha}) {
        // collective: every rank reduces all three vectors, present or not, in the same order
        if (hv->empty())
          FUS_CUDA(cudaMemsetAsync(m->d_un, 0, sizeof(double) * nd, c->stream));
        else
          FUS_CUDA(cudaMemcpyAsync(m->d_un, hv->data(), sizeof(double) * nd,
                                   cudaMemcpyHostToDevice, c->stream));
        FUS_TRY(halo_reverse(c->halo, m->d_un, nullptr, c->stream));
        hv->resize((size_t)nd);
        FUS_CUDA(cudaMemcpyAsync(hv->data(), m->d_un, sizeof(double) * nd, cudaMemcpyDeviceToHost,
                                 c->stream));
        FUS_CUDA(cudaStreamSynchronize(c->stream));
      }
      FUS_CUDA(cudaMemsetAsync(m->d_un, 0, sizeof(double) * nd, c->stream));
    }
    // compact them over the owned dofs, in dof order, with the first entry of every epilogue chunk
    std::vector<int32_t> bidx;
    std::vector<double> bs, bd, ba;
    const int64_t nchunks = (nd + kStageChunk - 1) / kStageChunk;
    std::vector<long long> bchunk((size_t)nchunks + 1, 0);
    for (int64_t i = 0; i < c->nowned; ++i) {
      const double a = hs.empty() ? 0.0 : hs[i], b = hd.empty() ? 0.0 : hd[i],
                   e = ha.empty() ? 0.0 : ha[i];
      if (a != 0.0 || b != 0.0 || e != 0.0) {
        bidx.push_back((int32_t)i);
        bs.push_back(a);
        bd.push_back(b);
        ba.push_back(e);
        ++bchunk[(size_t)(i / kStageChunk) + 1];
      }
    }
    for (int64_t k = 0; k < nchunks; ++k)
      bchunk[(size_t)k + 1] += bchunk[(size_t)k];
    m->nb = (int64_t)bidx.size();
    FUS_CUDA(cudaMalloc(&m->d_done, sizeof(unsigned int)));
    FUS_CUDA(cudaMemset(m->d_done, 0, sizeof(unsigned int)));
    if (m->nb) {
      FUS_CUDA(cudaMalloc(&m->d_bidx, sizeof(int32_t) * m->nb));
      FUS_CUDA(cudaMalloc(&m->d_bsrc, sizeof(double) * m->nb));
      FUS_CUDA(cudaMalloc(&m->d_bdsrc, sizeof(double) * m->nb));
      FUS_CUDA(cudaMalloc(&m->d_babs, sizeof(double) * m->nb));
      FUS_CUDA(cudaMalloc(&m->d_bchunk, sizeof(long long) * bchunk.size()));
      FUS_CUDA(cudaMemcpy(m->d_bidx, bidx.data(), sizeof(int32_t) * m->nb, cudaMemcpyHostToDevice));
      FUS_CUDA(cudaMemcpy(m->d_bsrc, bs.data(), sizeof(double) * m->nb, cudaMemcpyHostToDevice));
      FUS_CUDA(cudaMemcpy(m->d_bdsrc, bd.data(), sizeof(double) * m->nb, cudaMemcpyHostToDevice));
      FUS_CUDA(cudaMemcpy(m->d_babs, ba.data(), sizeof(double) * m->nb, cudaMemcpyHostToDevice));
      FUS_CUDA(cudaMemcpy(m->d_bchunk, bchunk.data(), sizeof(long long) * bchunk.size(),
                          cudaMemcpyHostToDevice));
    }
    FUS_CUDA(cudaStreamSynchronize(c->stream));
    return FUS_OK;
  };
  const int rc = build();
  if (rc != FUS_OK) { // nothing half-built is handed back
    fus_model_destroy(m);
    return rc;
  }
  *out = m;
  return FUS_OK;
}

int fus_model_destroy(fus_model* m) {
  if (!m)
    return FUS_OK;
  cudaSetDevice(m->ctx->device);
  cudaStreamSynchronize(m->ctx->stream);
  if (m->step_graph)
    cudaGraphExecDestroy(m->step_graph);
  cudaFree(m->d_src);
  cudaFree(m->d_stepctr);
  for (void* p : {(void*)m->d_lin, (void*)m->d_att, (void*)m->d_m, (void*)m->d_dnl,
                  (void*)m->d_bidx, (void*)m->d_bsrc, (void*)m->d_bdsrc, (void*)m->d_babs,
                  (void*)m->d_bchunk, (void*)m->d_done,
                  (void*)m->d_u0, (void*)m->d_v0, (void*)m->d_ua, (void*)m->d_va, (void*)m->d_un,
                  (void*)m->d_vn, (void*)m->d_b})
    cudaFree(p);
  --m->ctx->live_models;
  delete m;
  return FUS_OK;
}

int fus_model_set_state(fus_model* m, const double* u, const double* v) {
  if (!m)
    return FUS_ERR_ARG;
  fus_ctx* c = m->ctx;
  FUS_TRY(select_device(c));
  const size_t vb = sizeof(double) * c->ndofs;
  if (u)
    FUS_CUDA(cudaMemcpyAsync(m->d_u0, u, vb, cudaMemcpyHostToDevice, c->stream));
  else
    FUS_CUDA(cudaMemsetAsync(m->d_u0, 0, vb, c->stream));
  if (v)
    FUS_CUDA(cudaMemcpyAsync(m->d_v0, v, vb, cudaMemcpyHostToDevice, c->stream));
  else
    FUS_CUDA(cudaMemsetAsync(m->d_v0, 0, vb, c->stream));
  FUS_CUDA(cudaStreamSynchronize(c->stream));
  return FUS_OK;
}

int fus_model_get_state(fus_model* m, double* u, double* v) {
  if (!m)
    return FUS_ERR_ARG;
  fus_ctx* c = m->ctx;
  FUS_TRY(select_device(c));
  const size_t vb = sizeof(double) * c->ndofs;
  if (u)
    FUS_CUDA(cudaMemcpyAsync(u, m->d_u0, vb, cudaMemcpyDeviceToHost, c->stream));
  if (v)
    FUS_CUDA(cudaMemcpyAsync(v, m->d_v0, vb, cudaMemcpyDeviceToHost, c->stream));
  FUS_CUDA(cudaStreamSynchronize(c->stream));
  return check_peer_error(c);
}

int fus_model_state_dev(fus_model* m, double** u, double** v) {
  if (!m)
    return FUS_ERR_ARG;
  if (u)
    *u = m->d_u0;
  if (v)
    *v = m->d_v0;
  return FUS_OK;
}

int fus_model_get_mass(fus_model* m, double* mass) {
  if (!m || !mass)
    return FUS_ERR_ARG;
  fus_ctx* c = m->ctx;
  FUS_TRY(select_device(c));
  FUS_CUDA(cudaMemcpyAsync(mass, m->d_m, sizeof(double) * c->ndofs, cudaMemcpyDeviceToHost,
                           c->stream));
  FUS_CUDA(cudaStreamSynchronize(c->stream));
  return FUS_OK;
}

// Source scalars at time t (Linear.hpp:185-192, Lossy.hpp:199-220, Westervelt.hpp:220-240)
static void source_scalars(const fus_model* m, double t, double* g, double* dg) {
  double window, dwindow;
  if (t < m->period * m->window_length) {
    window = 0.5 * (1.0 - cos(m->freq * M_PI * t / m->window_length));
    dwindow = 0.5 * M_PI * m->freq / m->window_length * sin(m->freq * M_PI * t / m->window_length);
  } else {
    window = 1.0;
    dwindow = 0.0;
  }
  if (m->kind == FUS_LINEAR) {
    *g = window * m->p0 * m->w0 / m->s0 * cos(m->w0 * t);
    *dg = 0.0;
  } else {
    *g = window * 2.0 * m->p0 * m->w0 / m->s0 * cos(m->w0 * t);
    *dg = dwindow * 2.0 * m->p0 * m->w0 / m->s0 * cos(m->w0 * t)
          - window * 2.0 * m->p0 * m->w0 * m->w0 / m->s0 * sin(m->w0 * t);
  }
}

// b += K(lin) u [+ K(att) v] + boundary terms, with the halo exchange around it when partitioned:
// the right-hand side assembly of f1 (Linear.hpp:203-206, Lossy.hpp:229-234, Westervelt.hpp:260-265).
// u, v must have fresh ghosts on entry.
// b[d] += g src[d] + dg dsrc[d] - absb[d] v[d] over the (owned) boundary dofs, as a kernel of its
// own: f1, and the first stage of an rk4 call (every later stage is seeded by the epilogue before it)
static int launch_boundary(fus_model* m, const double* v, double g, double dg, bool from_table) {
  fus_ctx* c = m->ctx;
  if (!m->nb)
    return FUS_OK;
  ProfScope prof(c, 2, c->stream);
  boundary_kernel<<<grid_for(m->nb, 256, 1 << 30), 256, 0, c->stream>>>(
      m->d_b, v, m->d_bidx, m->d_bsrc, m->d_bdsrc, m->d_babs, m->nb, g, dg,
      from_table ? m->d_src : nullptr, m->d_stepctr, 0);
  FUS_LAUNCHED();
  return FUS_OK;
}

static int assemble_rhs(fus_model* m, double t, const double* u, const double* v,
                        bool fwd_pending, bool with_boundary, bool fused = false) {
  fus_ctx* c = m->ctx;
  double g = 0.0, dg = 0.0;
  if (with_boundary)
    source_scalars(m, t, &g, &dg);
  const double* x2 = (m->kind >= FUS_LOSSY) ? v : nullptr;
  const double* c2 = (m->kind >= FUS_LOSSY) ? m->d_att : nullptr;
  auto boundary = [&]() -> int {
    return with_boundary ? launch_boundary(m, v, g, dg, false) : FUS_OK;
  };
  // inside the RK4 loop the operator walks the cells backwards (see stiffness_line_kernel)
  struct Rev {
    fus_ctx* c;
    Rev(fus_ctx* c_, bool on) : c(c_) { c->reverse_cells = on ? 1 : 0; }
    ~Rev() { c->reverse_cells = 0; }
  } rev(c, !with_boundary && c->reverse_op);
  if (!c->halo) {
    FUS_TRY(launch_stiffness(c, u, x2, m->d_lin, c2, m->d_b, 0, c->ncells, c->stream));
    return boundary();
  }
  if (fused) {
    // The stage kernels exchange themselves (fus_halo_kernels.cuh).  The cells that touch shared
    // dofs first, in a launch that ends by shipping the ghost partial sums to their owners; then
    // everything else through the plain kernel while they travel.
    const long long ni = halo_fused_interface_cells(c->halo);
    FUS_TRY(launch_stiffness(c, u, x2, m->d_lin, c2, m->d_b, 0, ni, c->stream, halo_fused(c->halo)));
    if (ni == 0) // no cell to run: the exchange number still has to advance
      FUS_TRY(halo_fused_operator_skipped(c->halo, c->stream));
    return launch_stiffness(c, u, x2, m->d_lin, c2, m->d_b, ni, c->ncells, c->stream);
  }
  // NCCL transport.  In order (default): all cells, boundary terms, then the ghost -> owner sum.
  // Overlapped (option halo_overlap): cells are ordered [interface | interior] and the interior is
  // split in two so that BOTH exchanges hide behind cells that touch no shared dof:
  //   interior A  ||  owner->ghost update of (u,v) started by the caller (halo_forward_begin)
  //   interface cells + boundary terms (need the fresh ghosts)
  //   interior B  ||  ghost->owner sum of b
  // A few SMs are left free so that NCCL's kernels can start while a cell kernel runs.
  const bool ov = halo_overlap(c->halo) != 0;
  const long long ni = halo_interface_cells(c->halo);
  const long long mid = ov ? ni + (c->ncells - ni) / 2 : c->ncells;
  c->reserve_sms = ov ? c->halo_reserve : 0;
  int rc = launch_stiffness(c, u, x2, m->d_lin, c2, m->d_b, ni, mid, c->stream);
  if (rc == FUS_OK && fwd_pending)
    rc = halo_forward_end(c->halo, const_cast<double*>(u), const_cast<double*>(v), c->stream);
  if (rc == FUS_OK)
    rc = launch_stiffness(c, u, x2, m->d_lin, c2, m->d_b, 0, ni, c->stream);
  if (rc == FUS_OK)
    rc = boundary();
  if (rc == FUS_OK)
    rc = halo_reverse_begin(c->halo, m->d_b, c->stream);
  if (rc == FUS_OK)
    rc = launch_stiffness(c, u, x2, m->d_lin, c2, m->d_b, mid, c->ncells, c->stream);
  c->reserve_sms = 0;
  if (rc == FUS_OK)
    rc = halo_reverse_end(c->halo, m->d_b, c->stream);
  return rc;
}

int fus_model_f1(fus_model* m, double t, const double* u, const double* v, double* result) {
  if (!m || !u || !v || !result)
    return FUS_ERR_ARG;
  fus_ctx* c = m->ctx;
  FUS_TRY(select_device(c));
  if (c->nowned < c->ndofs && !c->halo) {
    set_error("fus_model_f1: the context has ghost dofs but no halo (fus_halo_setup)");
    return FUS_ERR_STATE;
  }
  const int64_t nd = c->ndofs;
  const size_t vb = sizeof(double) * nd;
  FUS_CUDA(cudaMemcpyAsync(m->d_un, u, vb, cudaMemcpyHostToDevice, c->stream));
  FUS_CUDA(cudaMemcpyAsync(m->d_vn, v, vb, cudaMemcpyHostToDevice, c->stream));
  FUS_CUDA(cudaMemsetAsync(m->d_b, 0, vb, c->stream));
  if (c->halo)
    FUS_TRY(halo_forward(c->halo, m->d_un, m->d_vn, c->stream));
  FUS_TRY(assemble_rhs(m, t, m->d_un, m->d_vn, false, true));
  const int grid = grid_for(nd, 256, 1 << 30);
  if (m->kind == FUS_WESTERVELT)
    f1_finish_kernel<true><<<grid, 256, 0, c->stream>>>(m->d_b, m->d_m, m->d_dnl, m->d_un,
                                                        m->d_vn, m->d_ua, nd);
  else
    f1_finish_kernel<false><<<grid, 256, 0, c->stream>>>(m->d_b, m->d_m, nullptr, m->d_un,
                                                         m->d_vn, m->d_ua, nd);
  FUS_LAUNCHED();
  FUS_CUDA(cudaMemcpyAsync(result, m->d_ua, vb, cudaMemcpyDeviceToHost, c->stream));
  FUS_CUDA(cudaMemsetAsync(m->d_b, 0, vb, c->stream));
  FUS_CUDA(cudaStreamSynchronize(c->stream));
  return FUS_OK;
}

} // extern "C"

template <int STAGE>
static int launch_stage(fus_model* m, const StageArgs& A) {
  fus_ctx* c = m->ctx;
  ProfScope prof(c, 1, c->stream);
  const int grid = grid_for(A.ntotal, kStageChunk, c->num_sms * 8);
  const bool west = m->kind == FUS_WESTERVELT;
  if (A.halo) {
    if (west)
      rk4_stage_kernel<STAGE, true, false, true><<<grid, kStageThreads, 0, c->stream>>>(A);
    else
      rk4_stage_kernel<STAGE, false, false, true><<<grid, kStageThreads, 0, c->stream>>>(A);
  } else if (c->stage_hints) {
    if (west)
      rk4_stage_kernel<STAGE, true, true><<<grid, kStageThreads, 0, c->stream>>>(A);
    else
      rk4_stage_kernel<STAGE, false, true><<<grid, kStageThreads, 0, c->stream>>>(A);
  } else {
    if (west)
      rk4_stage_kernel<STAGE, true, false><<<grid, kStageThreads, 0, c->stream>>>(A);
    else
      rk4_stage_kernel<STAGE, false, false><<<grid, kStageThreads, 0, c->stream>>>(A);
  }
  FUS_LAUNCHED();
  return FUS_OK;
}

// One RK4 step = 4 x (operator + boundary terms + fused epilogue), plus the halo traffic when
// partitioned.  Issued eagerly or captured into a CUDA graph by fus_model_rk4.
static int issue_step(fus_model* m, StageArgs& A, double dt) {
  fus_ctx* c = m->ctx;
  const bool fused = A.halo != nullptr;
  if (c->halo && !fused) // scatter_fwd of the step's first stage input (Linear.hpp:196-199)
    FUS_TRY(halo_forward_begin(c->halo, m->d_u0, m->d_v0, c->stream));
  for (int i = 0; i < 4; ++i) {
    const double* u_in = (i == 0) ? m->d_u0 : m->d_un;
    const double* v_in = (i == 0) ? m->d_v0 : m->d_vn;
    FUS_TRY(assemble_rhs(m, 0.0, u_in, v_in, true, false, fused));
    stage_coefficients(A, i, dt);
    switch (i) {
    case 0: FUS_TRY(launch_stage<0>(m, A)); break;
    case 1: FUS_TRY(launch_stage<1>(m, A)); break;
    case 2: FUS_TRY(launch_stage<2>(m, A)); break;
    case 3: FUS_TRY(launch_stage<3>(m, A)); break;
    }
    if (c->halo && !fused && i < 3) // next stage input; joined inside the next assemble_rhs
      FUS_TRY(halo_forward_begin(c->halo, m->d_un, m->d_vn, c->stream));
  }
  return FUS_OK;
}

extern "C" {

int fus_model_rk4(fus_model* m, double startTime, double finalTime, double timeStep,
                  int* nsteps) {
  if (!m || !(timeStep > 0.0)) {
    set_error("fus_model_rk4: bad argument");
    return FUS_ERR_ARG;
  }
  fus_ctx* c = m->ctx;
  FUS_TRY(select_device(c));
  if (c->nowned < c->ndofs && !c->halo) {
    set_error("fus_model_rk4: the context has ghost dofs but no halo (fus_halo_setup)");
    return FUS_ERR_STATE;
  }
  FUS_TRY(ensure_cell_layout(c)); // allocates: ahead of any stream capture
  // Same host-side time arithmetic as the reference loop (Linear.hpp:231-298), run ahead of the
  // device: the step sizes and the source scalars of every (step, stage) are tabulated first.
  const double c_runge[4] = {0.0, 0.5, 0.5, 1.0};
  std::vector<double> dts, table;
  {
    double t = startTime, tf = finalTime, dt = timeStep;
    while (t < tf) {
      dt = std::min(dt, tf - t);
      for (int i = 0; i < 4; ++i) {
        double g, dg;
        source_scalars(m, t + c_runge[i] * dt, &g, &dg);
        table.push_back(g);
        table.push_back(dg);
      }
      dts.push_back(dt);
      t += dt;
      if (t >= tf) { // row read by the last stage-3 epilogue when it seeds a step that is not taken
        double g, dg;
        source_scalars(m, t, &g, &dg);
        table.push_back(g);
        table.push_back(dg);
      }
      if (dts.size() > (size_t)100000000) {
        set_error("fus_model_rk4: more than 1e8 steps requested");
        return FUS_ERR_ARG;
      }
    }
  }
  const int step_total = (int)dts.size();
  if (nsteps)
    *nsteps = step_total;
  if (step_total == 0)
    return FUS_OK;
  if (table.size() > m->src_cap) {
    FUS_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(m->d_src);
    m->d_src = nullptr;
    m->src_cap = std::max<size_t>(table.size(), 4096);
    FUS_CUDA(cudaMalloc(&m->d_src, sizeof(double) * m->src_cap));
    if (m->step_graph) { // the graph holds the old table pointer
      cudaGraphExecDestroy(m->step_graph);
      m->step_graph = nullptr;
    }
  }
  if (!m->d_stepctr)
    FUS_CUDA(cudaMalloc(&m->d_stepctr, sizeof(int)));
  // the previous call may still be reading the table: the copy is stream-ordered after it
  FUS_CUDA(cudaMemcpyAsync(m->d_src, table.data(), sizeof(double) * table.size(),
                           cudaMemcpyHostToDevice, c->stream));
  FUS_CUDA(cudaStreamSynchronize(c->stream)); // `table` is pageable and goes out of scope
  FUS_CUDA(cudaMemsetAsync(m->d_stepctr, 0, sizeof(int), c->stream));

  StageArgs A;
  A.b = m->d_b;
  A.m = m->d_m;
  A.dnl = m->d_dnl;
  A.u0 = m->d_u0;
  A.v0 = m->d_v0;
  A.ua = m->d_ua;
  A.va = m->d_va;
  A.un = m->d_un;
  A.vn = m->d_vn;
  A.nowned = c->nowned;
  A.ntotal = c->ndofs;
  A.step_ctr = m->d_stepctr;
  A.done_ctr = m->d_done;
  A.nb = m->nb;
  A.bidx = m->d_bidx;
  A.bsrc = m->d_bsrc;
  A.bdsrc = m->d_bdsrc;
  A.babs = m->d_babs;
  A.bchunk = m->d_bchunk;
  A.src_table = m->d_src;
  // Fused peer transport: streamed G only (the HALO kernels are built for it); otherwise this call
  // runs on NCCL in stream order.
  const bool fused = c->halo && halo_mode(c->halo) == 2 && c->geom_active == 0 && c->dim == 3;
  A.halo = fused ? halo_fused(c->halo) : nullptr;
  if (fused) // handshake with the neighbours + owner -> ghost update of (u_n, v_n)
    FUS_TRY(halo_fused_entry(c->halo, m->d_u0, m->d_v0, c->stream));
  // boundary terms of the first stage; every later stage is seeded by the epilogue before it
  FUS_CUDA(cudaMemsetAsync(m->d_b, 0, sizeof(double) * c->ndofs, c->stream));
  FUS_TRY(launch_boundary(m, m->d_v0, 0.0, 0.0, true));
  // Optional: pin b in the persisting part of the 126 MB L2 (measured slower overall, off).
  bool l2_window = false;
  if (c->l2_persist) {
    cudaDeviceProp prop;
    FUS_CUDA(cudaGetDeviceProperties(&prop, c->device));
    const size_t want = sizeof(double) * (size_t)c->ndofs;
    const size_t set_aside = std::min<size_t>((size_t)prop.persistingL2CacheMaxSize, want);
    const size_t window = std::min<size_t>((size_t)prop.accessPolicyMaxWindowSize, want);
    if (set_aside > 0 && window > 0) {
      FUS_CUDA(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, set_aside));
      cudaStreamAttrValue av;
      std::memset(&av, 0, sizeof(av));
      av.accessPolicyWindow.base_ptr = m->d_b;
      av.accessPolicyWindow.num_bytes = window;
      av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)set_aside / (double)window);
      av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      FUS_CUDA(cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &av));
      l2_window = true;
    }
  }

  // CUDA graph: steps are identical launches once the scalars come from the table, so one step
  // is captured (with the fused peer transport it is still a single stream) and replayed.  Not used while
  // per-kernel profiling is on (event pairs), with the NCCL transport, or for the odd last step.
  const int hmode = c->halo ? (fused ? 2 : halo_mode(c->halo) % 2) : -1;
  const bool graph_ok = m->use_graph && c->use_graph && !c->profile && !l2_window
                        && (hmode == -1 || hmode == 2);
  if (m->step_graph
      && (m->graph_dt != dts[0] || m->graph_stream != c->stream || m->graph_halo_mode != hmode
          || m->graph_epoch != c->config_epoch)) {
    cudaGraphExecDestroy(m->step_graph);
    m->step_graph = nullptr;
  }
  int rc = FUS_OK;
  for (int s = 0; s < step_total && rc == FUS_OK; ++s) {
    const bool full = dts[s] == dts[0];
    if (graph_ok && full && m->step_graph) {
      FUS_CUDA(cudaGraphLaunch(m->step_graph, c->stream));
      g_launches.fetch_add(m->graph_launches, std::memory_order_relaxed);
      continue;
    }
    // capture once everything lazy (function attributes, occupancy) has run eagerly: step >= 1
    const bool capture = graph_ok && m->use_graph && full && s >= 1 && step_total - s >= 2;
    if (capture) {
      cudaGraph_t graph = nullptr;
      const long long before = g_launches.load();
      if (cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
        // e.g. the legacy default stream cannot be captured: stay on eager issue
        cudaGetLastError();
        m->use_graph = false;
        rc = issue_step(m, A, dts[s]);
        continue;
      }
      rc = issue_step(m, A, dts[s]);
      cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
      m->graph_launches = g_launches.load() - before; // kernels per replay (captured, not run)
      g_launches.store(before);
      if (rc == FUS_OK && ce == cudaSuccess && graph) {
        ce = cudaGraphInstantiate(&m->step_graph, graph, 0);
        if (ce == cudaSuccess) {
          m->graph_dt = dts[0];
          m->graph_stream = c->stream;
          m->graph_halo_mode = hmode;
          m->graph_epoch = c->config_epoch;
        } else {
          m->step_graph = nullptr;
        }
      }
      if (graph)
        cudaGraphDestroy(graph);
      if (rc != FUS_OK)
        break;
      if (!m->step_graph) { // capture unavailable here: fall back to eager issue for good
        cudaGetLastError();
        m->use_graph = false;
        rc = issue_step(m, A, dts[s]);
      } else {
        FUS_CUDA(cudaGraphLaunch(m->step_graph, c->stream));
        g_launches.fetch_add(m->graph_launches, std::memory_order_relaxed);
      }
      continue;
    }
    rc = issue_step(m, A, dts[s]);
  }
  if (rc != FUS_OK)
    return rc;
  if (fused) { // u_n, v_n leave with fresh ghosts (Linear.hpp:312-313): sent by the last epilogue
    FUS_TRY(halo_fused_exit(c->halo, m->d_u0, m->d_v0, c->stream));
  } else if (c->halo) {
    FUS_TRY(halo_forward_begin(c->halo, m->d_u0, m->d_v0, c->stream));
    FUS_TRY(halo_forward_end(c->halo, m->d_u0, m->d_v0, c->stream));
  }
  if (l2_window) {
    cudaStreamAttrValue av;
    std::memset(&av, 0, sizeof(av));
    FUS_CUDA(cudaStreamSetAttribute(c->stream, cudaStreamAttributeAccessPolicyWindow, &av));
    FUS_CUDA(cudaCtxResetPersistingL2Cache());
  }
  return FUS_OK;
}

// ---- halo ---------------------------------------------------------------------------------------
int fus_comm_unique_id(void* id128) { return halo_unique_id(id128); }

int fus_halo_setup(fus_ctx* c, int rank, int nranks, const void* uid, int nneigh, const int* neigh,
                   const int64_t* send_off, const int32_t* send_idx, const int64_t* recv_off,
                   const int32_t* recv_idx, int64_t ninterface_cells) {
  if (!c || nranks < 1 || rank < 0 || rank >= nranks || nneigh < 0 || ninterface_cells < 0
      || ninterface_cells > c->ncells) {
    set_error("fus_halo_setup: bad argument");
    return FUS_ERR_ARG;
  }
  FUS_TRY(select_device(c));
  ++c->config_epoch;
  if (c->halo) {
    halo_destroy(c->halo);
    c->halo = nullptr;
  }
  if (c->live_models > 0) { // their lumped mass and boundary vectors were reduced over the old halo
    set_error("fus_halo_setup: %d model(s) built on this context are alive; set the halo up first",
              c->live_models);
    return FUS_ERR_STATE;
  }
  int rc = halo_create(&c->halo, c->device, rank, nranks, uid, nneigh, neigh, send_off, send_idx,
                       recv_off, recv_idx, c->nowned, c->ndofs, ninterface_cells);
  if (rc != FUS_OK && c->halo) { // nothing half-built stays attached to the context
    halo_destroy(c->halo);
    c->halo = nullptr;
  }
  if (rc == FUS_OK) { // experiment knobs
    if (const char* e = std::getenv("FUS_HALO_OVERLAP"))
      halo_set_overlap(c->halo, std::atoi(e));
    if (const char* e = std::getenv("FUS_HALO_RESERVE"))
      c->halo_reserve = std::max(0, std::min(c->num_sms - 1, std::atoi(e)));
  }
  return rc;
}

int fus_halo_mailbox_layout(int64_t nsend, int64_t nrecv, int nneigh, int64_t* layout6) {
  if (nsend < 0 || nrecv < 0 || nneigh < 0 || !layout6)
    return FUS_ERR_ARG;
  halo_mailbox_layout(nsend, nrecv, nneigh, layout6);
  return FUS_OK;
}

int fus_halo_peer_offsets(const int64_t* q_layout6, const int64_t* q_send_off,
                          const int64_t* q_recv_off, int j, int64_t* out6) {
  if (!q_layout6 || !q_send_off || !q_recv_off || j < 0 || !out6)
    return FUS_ERR_ARG;
  out6[0] = 8 * q_recv_off[j];
  out6[1] = q_layout6[0] + 8 * q_recv_off[j];
  out6[2] = q_layout6[1] + 8 * q_send_off[j];
  out6[3] = q_layout6[2] + 8 * (int64_t)j;
  out6[4] = q_layout6[3] + 8 * (int64_t)j;
  out6[5] = q_layout6[4] + 8 * (int64_t)j;
  return FUS_OK;
}

int fus_halo_peer_export(fus_ctx* c, void* ipc_handle64, int64_t* layout6, void** base) {
  if (!c || !c->halo) {
    set_error("fus_halo_peer_export: call fus_halo_setup first");
    return FUS_ERR_STATE;
  }
  FUS_TRY(select_device(c));
  return halo_peer_export(c->halo, ipc_handle64, layout6, base);
}

int fus_halo_peer_connect(fus_ctx* c, const void* handles, const int64_t* byte_off) {
  if (!c || !c->halo)
    return FUS_ERR_STATE;
  FUS_TRY(select_device(c));
  ++c->config_epoch;
  return halo_peer_connect(c->halo, handles, byte_off);
}

int fus_halo_peer_connect_local(fus_ctx* c, void* const* bases, const int* devices,
                                const int64_t* byte_off) {
  if (!c || !c->halo)
    return FUS_ERR_STATE;
  FUS_TRY(select_device(c));
  ++c->config_epoch;
  return halo_peer_connect_local(c->halo, bases, devices, byte_off);
}

int fus_scatter_fwd_dev(fus_ctx* c, double* x) {
  if (!c || !x)
    return FUS_ERR_ARG;
  if (!c->halo)
    return FUS_OK;
  FUS_TRY(select_device(c));
  return halo_forward(c->halo, x, nullptr, c->stream);
}

int fus_scatter_rev_dev(fus_ctx* c, double* x) {
  if (!c || !x)
    return FUS_ERR_ARG;
  if (!c->halo)
    return FUS_OK;
  FUS_TRY(select_device(c));
  return halo_reverse(c->halo, x, nullptr, c->stream);
}

} // extern "C"
