// fus_halo.cu -- halo exchange over NCCL point-to-point (NVLink 5 / NVSwitch on an 8xB200 box).
//
// The exchange is a neighbour halo, not a reduction over all ranks: per neighbour one grouped
// ncclSend + ncclRecv of the packed interface values.  NCCL is resolved at run time from the
// libnccl.so.2 already loaded in the process (torch's) or found by the loader, so that the
// single-GPU path has no NCCL dependency.
#include "fus_halo.hpp"
#include "fus_halo_kernels.cuh"
#include "fus_internal.hpp"

#include <dlfcn.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace fus {

namespace {
// Minimal NCCL surface (ABI-stable since NCCL 2.7)
typedef struct ncclComm* ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;
constexpr int kNcclFloat64 = 8;

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;

int load_nccl() {
  if (g_nccl.lib)
    return FUS_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!h)
    h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h)
    h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    set_error("cannot load libnccl.so.2: %s", dlerror());
    return FUS_ERR_COMM;
  }
  auto sym = [&](const char* n) { return dlsym(h, n); };
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))sym("ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))sym("ncclCommInitRank");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))sym("ncclCommDestroy");
  g_nccl.GroupStart = (decltype(g_nccl.GroupStart))sym("ncclGroupStart");
  g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))sym("ncclGroupEnd");
  g_nccl.Send = (decltype(g_nccl.Send))sym("ncclSend");
  g_nccl.Recv = (decltype(g_nccl.Recv))sym("ncclRecv");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))sym("ncclGetErrorString");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.GroupStart || !g_nccl.GroupEnd
      || !g_nccl.Send || !g_nccl.Recv) {
    set_error("libnccl.so.2 lacks a required symbol");
    return FUS_ERR_COMM;
  }
  g_nccl.lib = h;
  return FUS_OK;
}

#define FUS_NCCL(call)                                                                             \
  do {                                                                                             \
    ncclResult_t r__ = (call);                                                                     \
    if (r__ != 0) {                                                                                \
      set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,                                      \
                g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "nccl error");                \
      return FUS_ERR_COMM;                                                                         \
    }                                                                                              \
  } while (0)

#define FUS_CUDA_H(call)                                                                           \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess) {                                                                      \
      set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__));            \
      return FUS_ERR_CUDA;                                                                         \
    }                                                                                              \
  } while (0)

inline int blocks_for(long long n) { return (int)std::max<long long>(1, (n + 255) / 256); }

// FUS_HALO_PROF=1: device time of every exchange phase, printed per rank when the halo is destroyed
struct PhaseProf {
  const char* name;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev;
};
bool g_prof = std::getenv("FUS_HALO_PROF") != nullptr;
PhaseProf g_phase[6] = {{"unused0", {}}, {"unused1", {}}, {"unused2", {}},
                        {"unused3", {}}, {"nccl_fwd", {}}, {"nccl_rev", {}}};
struct PhaseScope {
  cudaEvent_t stop = nullptr;
  cudaStream_t st;
  PhaseScope(int phase, cudaStream_t s) : st(s) {
    if (!g_prof || g_phase[phase].ev.size() > 4000)
      return;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a, st);
    g_phase[phase].ev.push_back({a, b});
    stop = b;
  }
  ~PhaseScope() {
    if (stop)
      cudaEventRecord(stop, st);
  }
};
void phase_report(int rank) {
  if (!g_prof)
    return;
  cudaDeviceSynchronize();
  for (auto& ph : g_phase) {
    if (ph.ev.empty())
      continue;
    std::vector<float> v;
    for (auto& e : ph.ev) {
      float ms = 0;
      cudaEventElapsedTime(&ms, e.first, e.second);
      v.push_back(ms);
    }
    std::sort(v.begin(), v.end());
    std::fprintf(stderr, "[fus halo rank %d] %-9s n=%zu median=%.1f us p90=%.1f us max=%.1f us\n",
                 rank, ph.name, v.size(), 1e3 * v[v.size() / 2], 1e3 * v[(v.size() * 9) / 10],
                 1e3 * v.back());
    ph.ev.clear();
  }
}
} // namespace

struct Halo {
  int device = 0, rank = 0, nranks = 1;
  ncclComm_t comm = nullptr;
  std::vector<int> neigh;
  std::vector<int64_t> send_off, recv_off; // per neighbour, in entries
  int64_t nsend = 0, nrecv = 0;
  int32_t *d_send_idx = nullptr, *d_recv_idx = nullptr;
  int64_t *d_soff = nullptr, *d_roff = nullptr; // device copies of send_off / recv_off
  double *d_sbuf = nullptr, *d_rbuf = nullptr; // 2 vectors deep
  int64_t nowned = 0, ndofs = 0, ninterface = 0;
  int overlap = 0; // NCCL on a side stream: measured slower than in-order beyond 2 ranks (profiles/)
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_ready = nullptr, ev_done = nullptr, ev_fwd_ready = nullptr, ev_fwd_done = nullptr;
  // ---- fused peer transport (fus_halo_kernels.cuh): the stage kernels exchange over NVLink ----
  // mailbox = [fwd_u: nghost][fwd_v: nghost][rev: nsend][fwd flags][rev flags][ready flags]
  bool peer = false;
  bool ipc = false;                          // peer_base[] were opened with cudaIpcOpenMemHandle
  char* d_mbox = nullptr;
  int64_t lay[6] = {0, 0, 0, 0, 0, 0};       // byte offsets {fwd_v, rev, fwd flags, rev flags, ready, total}
  std::vector<void*> peer_base;              // the neighbours' mailboxes
  FusedHalo* d_fh = nullptr;                 // device copy of the transport state
  FusedHalo h_fh;                            // host copy (options rewrite single fields)
  int64_t nshared = 0;
  int32_t* d_spos_off = nullptr;
  int32_t* d_spos = nullptr;
  signed char* d_spos_nb = nullptr;
  unsigned int* d_ctr = nullptr;
  int* d_error = nullptr;
  unsigned long long* d_seq = nullptr;       // exchange numbers, advanced on the device so that a
                                             // captured CUDA graph can replay the step
};

int halo_unique_id(void* id128) {
  if (!id128)
    return FUS_ERR_ARG;
  int r = load_nccl();
  if (r != FUS_OK)
    return r;
  ncclUniqueId id;
  FUS_NCCL(g_nccl.GetUniqueId(&id));
  std::memcpy(id128, &id, sizeof(id));
  return FUS_OK;
}

int halo_create(Halo** out, int device, int rank, int nranks, const void* uid, int nneigh,
                const int* neigh, const int64_t* send_off, const int32_t* send_idx,
                const int64_t* recv_off, const int32_t* recv_idx, int64_t nowned, int64_t ndofs,
                int64_t ninterface_cells) {
  int r = load_nccl();
  if (r != FUS_OK)
    return r;
  if (!uid || (nneigh > 0 && (!neigh || !send_off || !recv_off))) {
    set_error("halo_create: bad argument");
    return FUS_ERR_ARG;
  }
  // validate everything that can be validated on this rank before anything is allocated or any
  // collective is entered (a rank that returned early would leave the others in ncclCommInitRank)
  if (nneigh > 0) {
    const int64_t ns = send_off[nneigh], nr = recv_off[nneigh];
    if (ns < 0 || nr < 0 || (ns > 0 && !send_idx) || (nr > 0 && !recv_idx)) {
      set_error("halo_create: bad lists");
      return FUS_ERR_ARG;
    }
    for (int64_t i = 0; i < ns; ++i)
      if (send_idx[i] < 0 || send_idx[i] >= nowned) {
        set_error("halo_create: send index %d is not an owned dof", send_idx[i]);
        return FUS_ERR_ARG;
      }
    for (int64_t i = 0; i < nr; ++i)
      if (recv_idx[i] < nowned || recv_idx[i] >= ndofs) {
        set_error("halo_create: recv index %d is not a ghost dof", recv_idx[i]);
        return FUS_ERR_ARG;
      }
  }
  Halo* h = new Halo();
  *out = h;
  h->device = device;
  h->rank = rank;
  h->nranks = nranks;
  h->nowned = nowned;
  h->ndofs = ndofs;
  h->ninterface = ninterface_cells;
  h->neigh.assign(neigh, neigh + nneigh);
  if (nneigh > 0) {
    h->send_off.assign(send_off, send_off + nneigh + 1);
    h->recv_off.assign(recv_off, recv_off + nneigh + 1);
  } else {
    h->send_off.assign(1, 0);
    h->recv_off.assign(1, 0);
  }
  h->nsend = nneigh ? send_off[nneigh] : 0;
  h->nrecv = nneigh ? recv_off[nneigh] : 0;
  FUS_CUDA_H(cudaSetDevice(device));
  FUS_CUDA_H(cudaMalloc(&h->d_send_idx, sizeof(int32_t) * std::max<int64_t>(1, h->nsend)));
  FUS_CUDA_H(cudaMalloc(&h->d_recv_idx, sizeof(int32_t) * std::max<int64_t>(1, h->nrecv)));
  const int64_t nb = std::max<int64_t>(1, std::max(h->nsend, h->nrecv));
  FUS_CUDA_H(cudaMalloc(&h->d_sbuf, sizeof(double) * 2 * nb));
  FUS_CUDA_H(cudaMalloc(&h->d_rbuf, sizeof(double) * 2 * nb));
  if (h->nsend)
    FUS_CUDA_H(cudaMemcpy(h->d_send_idx, send_idx, sizeof(int32_t) * h->nsend,
                          cudaMemcpyHostToDevice));
  if (h->nrecv)
    FUS_CUDA_H(cudaMemcpy(h->d_recv_idx, recv_idx, sizeof(int32_t) * h->nrecv,
                          cudaMemcpyHostToDevice));
  int prio_lo = 0, prio_hi = 0;
  FUS_CUDA_H(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
  {
    const size_t ob = sizeof(int64_t) * (size_t)(nneigh + 1);
    FUS_CUDA_H(cudaMalloc(&h->d_soff, ob));
    FUS_CUDA_H(cudaMalloc(&h->d_roff, ob));
    FUS_CUDA_H(cudaMemcpy(h->d_soff, h->send_off.data(), ob, cudaMemcpyHostToDevice));
    FUS_CUDA_H(cudaMemcpy(h->d_roff, h->recv_off.data(), ob, cudaMemcpyHostToDevice));
  }
  FUS_CUDA_H(cudaStreamCreateWithPriority(&h->comm_stream, cudaStreamNonBlocking, prio_hi));
  FUS_CUDA_H(cudaEventCreateWithFlags(&h->ev_ready, cudaEventDisableTiming));
  FUS_CUDA_H(cudaEventCreateWithFlags(&h->ev_done, cudaEventDisableTiming));
  FUS_CUDA_H(cudaEventCreateWithFlags(&h->ev_fwd_ready, cudaEventDisableTiming));
  FUS_CUDA_H(cudaEventCreateWithFlags(&h->ev_fwd_done, cudaEventDisableTiming));
  ncclUniqueId id;
  std::memcpy(&id, uid, sizeof(id));
  FUS_NCCL(g_nccl.CommInitRank(&h->comm, nranks, id, rank));
  return FUS_OK;
}

void halo_destroy(Halo* h) {
  if (!h)
    return;
  cudaSetDevice(h->device);
  phase_report(h->rank);
  if (h->comm_stream)
    cudaStreamSynchronize(h->comm_stream);
  if (h->comm && g_nccl.CommDestroy)
    g_nccl.CommDestroy(h->comm);
  if (h->ipc)
    for (void* pb : h->peer_base)
      if (pb)
        cudaIpcCloseMemHandle(pb);
  cudaFree(h->d_mbox);
  cudaFree(h->d_fh);
  cudaFree(h->d_spos_off);
  cudaFree(h->d_spos);
  cudaFree(h->d_spos_nb);
  cudaFree(h->d_ctr);
  cudaFree(h->d_seq);
  cudaFree(h->d_error);
  cudaFree(h->d_soff);
  cudaFree(h->d_roff);
  cudaFree(h->d_send_idx);
  cudaFree(h->d_recv_idx);
  cudaFree(h->d_sbuf);
  cudaFree(h->d_rbuf);
  if (h->ev_ready)
    cudaEventDestroy(h->ev_ready);
  if (h->ev_done)
    cudaEventDestroy(h->ev_done);
  if (h->ev_fwd_ready)
    cudaEventDestroy(h->ev_fwd_ready);
  if (h->ev_fwd_done)
    cudaEventDestroy(h->ev_fwd_done);
  if (h->comm_stream)
    cudaStreamDestroy(h->comm_stream);
  delete h;
}

void halo_set_overlap(Halo* h, int on) { h->overlap = on; }
int halo_overlap(const Halo* h) { return h->overlap; }
int halo_mode(const Halo* h) { return h->peer ? 2 : (h->overlap ? 1 : 0); }
long long halo_interface_cells(const Halo* h) { return h->overlap ? h->ninterface : 0; }

// One grouped exchange.  `fwd`: owners send send_idx entries, ghosts receive; otherwise reversed.
// nv vectors are concatenated per neighbour: [neighbour k][vector][entry].
static int exchange(Halo* h, bool fwd, int nv, cudaStream_t st) {
  PhaseScope ps(fwd ? 4 : 5, st);
  const std::vector<int64_t>& soff = fwd ? h->send_off : h->recv_off;
  const std::vector<int64_t>& roff = fwd ? h->recv_off : h->send_off;
  FUS_NCCL(g_nccl.GroupStart());
  for (size_t k = 0; k < h->neigh.size(); ++k) {
    const int64_t ns = soff[k + 1] - soff[k], nr = roff[k + 1] - roff[k];
    if (ns)
      FUS_NCCL(g_nccl.Send(h->d_sbuf + nv * soff[k], (size_t)(nv * ns), kNcclFloat64, h->neigh[k],
                           h->comm, st));
    if (nr)
      FUS_NCCL(g_nccl.Recv(h->d_rbuf + nv * roff[k], (size_t)(nv * nr), kNcclFloat64, h->neigh[k],
                           h->comm, st));
  }
  FUS_NCCL(g_nccl.GroupEnd());
  return FUS_OK;
}

namespace {
// device copies of the per-neighbour offset tables (owned by the Halo, made in halo_create)
struct OffTables {
  int64_t *d_soff, *d_roff;
};
inline OffTables tables(Halo* h) { return OffTables{h->d_soff, h->d_roff}; }
} // namespace

int halo_forward(Halo* h, double* a, double* b, cudaStream_t st) {
  if (h->neigh.empty())
    return FUS_OK;
  const int nv = b ? 2 : 1, nn = (int)h->neigh.size();
  const OffTables T = tables(h);
  if (h->nsend)
    halo_pack_kernel<<<blocks_for(h->nsend), 256, 0, st>>>(a, b, h->d_send_idx, T.d_soff, nn,
                                                          h->d_sbuf, h->nsend, nv);
  int r = exchange(h, true, nv, st);
  if (r != FUS_OK)
    return r;
  if (h->nrecv)
    halo_unpack_kernel<false><<<blocks_for(h->nrecv), 256, 0, st>>>(
        a, b, h->d_recv_idx, T.d_roff, nn, h->d_rbuf, h->nrecv, nv);
  FUS_CUDA_H(cudaGetLastError());
  return FUS_OK;
}

// ---------------------------------------------------------------------------------------------
// Fused peer transport: set-up.  Every rank owns a mailbox in device memory that its neighbours map
// (CUDA IPC across processes, plain peer access between the devices of one process); the exchange
// itself happens inside the stage kernels (fus_halo_kernels.cuh, fus_kernels.cuh).
// ---------------------------------------------------------------------------------------------
// mailbox = [fwd_u: nrecv doubles][fwd_v: nrecv][rev: nsend][fwd flags][rev flags][ready flags];
// layout6 = byte offsets {fwd_v, rev, forward flags, reverse flags, ready flags, total size}
void halo_mailbox_layout(int64_t nsend, int64_t nrecv, int nneigh, int64_t* layout6) {
  auto up = [](int64_t v) { return (v + 127) / 128 * 128; };
  const int64_t nn = std::max(1, nneigh);
  const int64_t vec = up((int64_t)sizeof(double) * std::max<int64_t>(1, nrecv));
  layout6[0] = vec;
  layout6[1] = 2 * vec;
  layout6[2] = layout6[1] + up((int64_t)sizeof(double) * std::max<int64_t>(1, nsend));
  layout6[3] = layout6[2] + up((int64_t)sizeof(unsigned long long) * nn);
  layout6[4] = layout6[3] + up((int64_t)sizeof(unsigned long long) * nn);
  layout6[5] = layout6[4] + up((int64_t)sizeof(unsigned long long) * nn);
}

int halo_peer_export(Halo* h, void* ipc_handle64, int64_t* layout6, void** base) {
  if (!h || !layout6)
    return FUS_ERR_ARG;
  if ((int)h->neigh.size() > kMaxNeigh) {
    set_error("peer transport supports at most %d neighbours", kMaxNeigh);
    return FUS_ERR_UNSUPPORTED;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  FUS_CUDA_H(cudaSetDevice(h->device));
  if (!h->d_mbox) {
    halo_mailbox_layout(h->nsend, h->nrecv, (int)h->neigh.size(), h->lay);
    FUS_CUDA_H(cudaMalloc(&h->d_mbox, (size_t)h->lay[5]));
    FUS_CUDA_H(cudaMemset(h->d_mbox, 0, (size_t)h->lay[5]));
    FUS_CUDA_H(cudaMalloc(&h->d_ctr, CTR_COUNT * sizeof(unsigned int)));
    FUS_CUDA_H(cudaMemset(h->d_ctr, 0, CTR_COUNT * sizeof(unsigned int)));
    FUS_CUDA_H(cudaMalloc(&h->d_seq, SEQ_COUNT * sizeof(unsigned long long)));
    FUS_CUDA_H(cudaMemset(h->d_seq, 0, SEQ_COUNT * sizeof(unsigned long long)));
    FUS_CUDA_H(cudaMalloc(&h->d_error, sizeof(int)));
    FUS_CUDA_H(cudaMemset(h->d_error, 0, sizeof(int)));
    FUS_CUDA_H(cudaDeviceSynchronize());
  }
  if (ipc_handle64) {
    cudaIpcMemHandle_t hd;
    FUS_CUDA_H(cudaIpcGetMemHandle(&hd, h->d_mbox));
    std::memcpy(ipc_handle64, &hd, sizeof(hd));
  }
  if (base)
    *base = h->d_mbox;
  std::memcpy(layout6, h->lay, sizeof(h->lay));
  return FUS_OK;
}

// bases[k]: neighbour k's mailbox as mapped into this process (IPC) or its device pointer (same
// process); byte_off[k][6]: byte offsets inside it of {my forward-u run, my forward-v run, my reverse
// run, my forward flag, my reverse flag, my ready flag}.  From neighbour q's layout6 = L, its offset
// tables and j = this rank's position in q's neighbour list:
//   fwd_u = 8 roff_q[j]   fwd_v = L[0] + 8 roff_q[j]   rev = L[1] + 8 soff_q[j]
//   flags = L[2] + 8 j,  L[3] + 8 j,  L[4] + 8 j
static int peer_connect_bases(Halo* h, const int64_t* byte_off) {
  const size_t nn = h->neigh.size();
  std::vector<int32_t> sidx((size_t)h->nsend), ridx((size_t)h->nrecv);
  if (h->nsend)
    FUS_CUDA_H(cudaMemcpy(sidx.data(), h->d_send_idx, sizeof(int32_t) * h->nsend,
                          cudaMemcpyDeviceToHost));
  if (h->nrecv)
    FUS_CUDA_H(cudaMemcpy(ridx.data(), h->d_recv_idx, sizeof(int32_t) * h->nrecv,
                          cudaMemcpyDeviceToHost));
  std::vector<int32_t> spos_off, spos;
  std::vector<signed char> spos_nb;
  const std::string why = fused_halo_lists(h->nowned, (int)nn, h->send_off.data(), sidx.data(),
                                           ridx.data(), h->nrecv, &h->nshared, spos_off, spos, spos_nb);
  if (!why.empty()) {
    set_error("fused halo: %s", why.c_str());
    return FUS_ERR_UNSUPPORTED;
  }
  FUS_CUDA_H(cudaMalloc(&h->d_spos_off, sizeof(int32_t) * spos_off.size()));
  FUS_CUDA_H(cudaMalloc(&h->d_spos, sizeof(int32_t) * std::max<size_t>(1, spos.size())));
  FUS_CUDA_H(cudaMalloc(&h->d_spos_nb, std::max<size_t>(1, spos_nb.size())));
  FUS_CUDA_H(cudaMemcpy(h->d_spos_off, spos_off.data(), sizeof(int32_t) * spos_off.size(),
                        cudaMemcpyHostToDevice));
  if (!spos.empty()) {
    FUS_CUDA_H(cudaMemcpy(h->d_spos, spos.data(), sizeof(int32_t) * spos.size(),
                          cudaMemcpyHostToDevice));
    FUS_CUDA_H(cudaMemcpy(h->d_spos_nb, spos_nb.data(), spos_nb.size(), cudaMemcpyHostToDevice));
  }
  FusedHalo& F = h->h_fh;
  std::memset(&F, 0, sizeof(F));
  F.nneigh = (int)nn;
  F.nowned = h->nowned;
  F.nghost = h->nrecv;
  F.nshared = h->nshared;
  F.nsend = h->nsend;
  F.fwd_u = (const double*)h->d_mbox;
  F.fwd_v = (const double*)(h->d_mbox + h->lay[0]);
  F.rev = (const double*)(h->d_mbox + h->lay[1]);
  F.fwd_flag = (const unsigned long long*)(h->d_mbox + h->lay[2]);
  F.rev_flag = (const unsigned long long*)(h->d_mbox + h->lay[3]);
  F.ready_flag = (const unsigned long long*)(h->d_mbox + h->lay[4]);
  for (size_t k = 0; k < nn; ++k) {
    char* cb = (char*)h->peer_base[k];
    F.r_fwd_u[k] = (double*)(cb + byte_off[6 * k + 0]);
    F.r_fwd_v[k] = (double*)(cb + byte_off[6 * k + 1]);
    F.r_rev[k] = (double*)(cb + byte_off[6 * k + 2]);
    F.r_fwd_flag[k] = (unsigned long long*)(cb + byte_off[6 * k + 3]);
    F.r_rev_flag[k] = (unsigned long long*)(cb + byte_off[6 * k + 4]);
    F.r_ready_flag[k] = (unsigned long long*)(cb + byte_off[6 * k + 5]);
  }
  F.soff = h->d_soff;
  F.roff = h->d_roff;
  F.sidx = h->d_send_idx;
  F.spos_off = h->d_spos_off;
  F.spos = h->d_spos;
  F.spos_nb = h->d_spos_nb;
  F.seq = h->d_seq;
  F.ctr = h->d_ctr;
  F.error = h->d_error;
  double timeout_s = 30.0; // a neighbour that is this late is taken for dead (FUS_HALO_TIMEOUT_S)
  if (const char* e = std::getenv("FUS_HALO_TIMEOUT_S"))
    timeout_s = std::max(0.001, std::atof(e));
  F.timeout_ns = (unsigned long long)(timeout_s * 1e9);
  if (!h->d_fh)
    FUS_CUDA_H(cudaMalloc(&h->d_fh, sizeof(FusedHalo)));
  FUS_CUDA_H(cudaMemcpy(h->d_fh, &F, sizeof(F), cudaMemcpyHostToDevice));
  h->peer = true;
  return FUS_OK;
}

int halo_peer_connect(Halo* h, const void* handles, const int64_t* byte_off) {
  if (!h || !h->d_mbox || (!h->neigh.empty() && (!handles || !byte_off))) {
    set_error("halo_peer_connect: export first, then pass the neighbours' handles and offsets");
    return FUS_ERR_ARG;
  }
  FUS_CUDA_H(cudaSetDevice(h->device));
  const size_t nn = h->neigh.size();
  h->peer_base.assign(nn, nullptr);
  h->ipc = true;
  for (size_t k = 0; k < nn; ++k) {
    cudaIpcMemHandle_t hd;
    std::memcpy(&hd, (const char*)handles + 64 * k, sizeof(hd));
    void* base = nullptr;
    FUS_CUDA_H(cudaIpcOpenMemHandle(&base, hd, cudaIpcMemLazyEnablePeerAccess));
    h->peer_base[k] = base;
  }
  return peer_connect_bases(h, byte_off);
}

// Same transport between the devices of ONE process (one host thread per GPU): the neighbours'
// mailboxes are plain device pointers; peer access is enabled here.
int halo_peer_connect_local(Halo* h, void* const* bases, const int* devices,
                            const int64_t* byte_off) {
  if (!h || !h->d_mbox || (!h->neigh.empty() && (!bases || !devices || !byte_off))) {
    set_error("halo_peer_connect_local: export first, then pass the neighbours' mailboxes");
    return FUS_ERR_ARG;
  }
  FUS_CUDA_H(cudaSetDevice(h->device));
  const size_t nn = h->neigh.size();
  h->peer_base.assign(nn, nullptr);
  h->ipc = false;
  for (size_t k = 0; k < nn; ++k) {
    if (devices[k] != h->device) {
      int can = 0;
      FUS_CUDA_H(cudaDeviceCanAccessPeer(&can, h->device, devices[k]));
      if (!can) {
        set_error("device %d cannot access device %d", h->device, devices[k]);
        return FUS_ERR_UNSUPPORTED;
      }
      const cudaError_t e = cudaDeviceEnablePeerAccess(devices[k], 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
        set_error("cudaDeviceEnablePeerAccess(%d): %s", devices[k], cudaGetErrorString(e));
        return FUS_ERR_CUDA;
      }
      cudaGetLastError();
    }
    h->peer_base[k] = bases[k];
  }
  return peer_connect_bases(h, byte_off);
}

int halo_peer_error(Halo* h) {
  if (!h || !h->d_error)
    return 0;
  int e = 0;
  if (cudaMemcpy(&e, h->d_error, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess)
    return 1;
  return e;
}

const FusedHalo* halo_fused(const Halo* h) { return (h && h->peer) ? h->d_fh : nullptr; }

// The by-value halo parameter of the stiffness launch over the interface cells [0, ninterface)
HaloLaunch halo_fused_launch(const Halo* h) {
  HaloLaunch L;
  std::memset(&L, 0, sizeof(L));
  if (!h || !h->peer)
    return L;
  L.H = h->d_fh;
  L.nown = h->nowned;
  L.mbu = h->h_fh.fwd_u - h->nowned;
  L.mbv = h->h_fh.fwd_v - h->nowned;
  return L;
}
long long halo_fused_interface_cells(const Halo* h) { return h ? h->ninterface : 0; }
int halo_fused_operator_skipped(Halo* h, cudaStream_t st) {
  halo_operator_skipped_kernel<<<1, 32, 0, st>>>(h->d_fh);
  FUS_CUDA_H(cudaGetLastError());
  return FUS_OK;
}

// Entry of an rk4 call in fused mode: handshake with the neighbours, then the owner -> ghost update
// of the state with the protocol the epilogues continue (see fus_halo_kernels.cuh).
int halo_fused_entry(Halo* h, const double* u, const double* v, cudaStream_t st) {
  halo_ready_kernel<<<1, 32, 0, st>>>(h->d_fh, 3);
  halo_entry_put_kernel<<<blocks_for(h->nsend), 256, 0, st>>>(h->d_fh, u, v, 0);
  FUS_CUDA_H(cudaGetLastError());
  return FUS_OK;
}

// Exit: the new state's ghost values, sent by the last epilogue, into the ghost entries of (u, v).
int halo_fused_exit(Halo* h, double* u, double* v, cudaStream_t st) {
  halo_exit_unpack_kernel<<<blocks_for(h->nrecv), 256, 0, st>>>(h->d_fh, u, v);
  FUS_CUDA_H(cudaGetLastError());
  return FUS_OK;
}

int halo_forward_begin(Halo* h, double* a, double* b, cudaStream_t st) {
  if (h->neigh.empty())
    return FUS_OK;
  if (!h->overlap)
    return halo_forward(h, a, b, st);
  FUS_CUDA_H(cudaEventRecord(h->ev_fwd_ready, st));
  FUS_CUDA_H(cudaStreamWaitEvent(h->comm_stream, h->ev_fwd_ready, 0));
  int r = halo_forward(h, a, b, h->comm_stream);
  if (r != FUS_OK)
    return r;
  FUS_CUDA_H(cudaEventRecord(h->ev_fwd_done, h->comm_stream));
  return FUS_OK;
}

int halo_forward_end(Halo* h, double* a, double* b, cudaStream_t st) {
  if (h->neigh.empty())
    return FUS_OK;
  if (!h->overlap)
    return FUS_OK;
  FUS_CUDA_H(cudaStreamWaitEvent(st, h->ev_fwd_done, 0));
  return FUS_OK;
}

static int reverse_on(Halo* h, double* a, double* b, cudaStream_t st) {
  const int nv = b ? 2 : 1, nn = (int)h->neigh.size();
  const OffTables T = tables(h);
  if (h->nrecv)
    halo_pack_kernel<<<blocks_for(h->nrecv), 256, 0, st>>>(a, b, h->d_recv_idx, T.d_roff, nn,
                                                          h->d_sbuf, h->nrecv, nv);
  int r = exchange(h, false, nv, st);
  if (r != FUS_OK)
    return r;
  if (h->nsend)
    halo_unpack_kernel<true><<<blocks_for(h->nsend), 256, 0, st>>>(
        a, b, h->d_send_idx, T.d_soff, nn, h->d_rbuf, h->nsend, nv);
  FUS_CUDA_H(cudaGetLastError());
  return FUS_OK;
}

int halo_reverse(Halo* h, double* a, double* b, cudaStream_t st) {
  if (h->neigh.empty())
    return FUS_OK;
  return reverse_on(h, a, b, st);
}

int halo_reverse_begin(Halo* h, double* a, cudaStream_t st) {
  if (h->neigh.empty())
    return FUS_OK;
  if (!h->overlap)
    return FUS_OK; // whole exchange happens in _end, after all cells
  FUS_CUDA_H(cudaEventRecord(h->ev_ready, st));
  FUS_CUDA_H(cudaStreamWaitEvent(h->comm_stream, h->ev_ready, 0));
  int r = reverse_on(h, a, nullptr, h->comm_stream);
  if (r != FUS_OK)
    return r;
  FUS_CUDA_H(cudaEventRecord(h->ev_done, h->comm_stream));
  return FUS_OK;
}

int halo_reverse_end(Halo* h, double* a, cudaStream_t st) {
  if (h->neigh.empty())
    return FUS_OK;
  if (!h->overlap)
    return reverse_on(h, a, nullptr, st);
  FUS_CUDA_H(cudaStreamWaitEvent(st, h->ev_done, 0));
  return FUS_OK;
}

} // namespace fus
