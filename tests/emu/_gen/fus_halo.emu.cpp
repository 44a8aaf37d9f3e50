// GENERATED from fenicsx-fus_b200/csrc/fus_halo.cu by build_emulated_library.py
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
// FUS_HALO_LIGHTFENCE=1: one system fence per block (thread 0, after the block barrier)
bool g_lightfence = std::getenv("FUS_HALO_LIGHTFENCE") != nullptr;
PhaseProf g_phase[6] = {{"put_fwd", {}}, {"wait_fwd", {}}, {"put_rev", {}},
                        {"wait_rev", {}}, {"nccl_fwd", {}}, {"nccl_rev", {}}};
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
  // ---- peer-direct transport: one-sided puts into the neighbours' mailboxes over NVLink ----
  // mailbox = [fwd data: 2*nrecv doubles][rev data: nsend doubles][fwd flags][rev flags]
  bool peer = false;
  char* d_mbox = nullptr;
  size_t off_rev = 0, off_fflag = 0, off_rflag = 0, mbox_bytes = 0;
  std::vector<void*> peer_base;              // opened mailboxes (one per neighbour)
  struct PeerTable* d_tab = nullptr;         // device copy of the per-neighbour destination table
  unsigned int* d_counter = nullptr;         // block-completion counters (fwd, rev)
  int* d_error = nullptr;
  unsigned long long* d_epoch = nullptr;     // [fwd, rev] exchange counters, advanced on the device
                                             // so that a captured CUDA graph can replay the step
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
  for (int64_t i = 0; i < h->nsend; ++i)
    if (send_idx[i] < 0 || send_idx[i] >= nowned) {
      set_error("halo_create: send index %d is not an owned dof", send_idx[i]);
      return FUS_ERR_ARG;
    }
  for (int64_t i = 0; i < h->nrecv; ++i)
    if (recv_idx[i] < nowned || recv_idx[i] >= ndofs) {
      set_error("halo_create: recv index %d is not a ghost dof", recv_idx[i]);
      return FUS_ERR_ARG;
    }
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
  for (void* pb : h->peer_base)
    if (pb)
      cudaIpcCloseMemHandle(pb);
  cudaFree(h->d_mbox);
  cudaFree(h->d_tab);
  cudaFree(h->d_counter);
  cudaFree(h->d_epoch);
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
int halo_overlap(const Halo* h) { return h->overlap || h->peer; }
int halo_mode(const Halo* h) { return h->peer ? 2 : (h->overlap ? 1 : 0); }
long long halo_interface_cells(const Halo* h) {
  return (h->overlap || h->peer) ? h->ninterface : 0;
}

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
    FUS_EMU_LAUNCH((halo_pack_kernel), blocks_for(h->nsend), 256, 0, st, a, b, h->d_send_idx, T.d_soff, nn,
                                                          h->d_sbuf, h->nsend, nv);
  int r = exchange(h, true, nv, st);
  if (r != FUS_OK)
    return r;
  if (h->nrecv)
    FUS_EMU_LAUNCH((halo_unpack_kernel<false>), blocks_for(h->nrecv), 256, 0, st, a, b, h->d_recv_idx, T.d_roff, nn, h->d_rbuf, h->nrecv, nv);
  FUS_CUDA_H(cudaGetLastError());
  return FUS_OK;
}

// ---------------------------------------------------------------------------------------------
// Peer-direct transport.  A put kernel gathers the interface values and stores them straight into
// the neighbours' mailboxes (IPC-mapped peer memory, NVLink); the last block to finish raises one
// epoch flag per neighbour with system-scope release ordering.  The receiving side's wait kernel
// spins on its own flags (acquire, bounded), then unpacks with L1-bypassing loads.
// Puts are one-sided, so they are issued as early as possible and the waits as late as possible:
// the cells that touch no shared dof run in between on the same stream.
// A mailbox segment is never overwritten before it is consumed because every exchanging pair
// alternates forward (owner -> ghost) and reverse (ghost -> owner) messages: the owner cannot put
// stage s+1 before it has received the reverse message of stage s, which the ghost side only
// sends after it has unpacked the forward message of stage s (and symmetrically).
// ---------------------------------------------------------------------------------------------
static int peer_put(Halo* h, bool fwd, double* a, double* b, cudaStream_t st) {
  if (fwd && !b) {
    set_error("peer transport: the forward update always carries two vectors");
    return FUS_ERR_ARG;
  }
  const int nv = fwd ? 2 : 1, nn = (int)h->neigh.size();
  const OffTables T = tables(h);
  const long long n = fwd ? h->nsend : h->nrecv;
  // launched even with nothing to send: the kernel also advances the exchange counter
  PhaseScope ps(fwd ? 0 : 2, st);
  FUS_EMU_LAUNCH((peer_put_kernel), blocks_for(n), 256, 0, st, a, b, fwd ? h->d_send_idx : h->d_recv_idx,
                                                 fwd ? T.d_soff : T.d_roff, nn, n, nv, h->d_tab,
                                                 fwd ? 1 : 0, h->d_counter + (fwd ? 0 : 1),
                                                 h->d_epoch + (fwd ? 0 : 1), g_lightfence ? 1 : 0);
  FUS_CUDA_H(cudaGetLastError());
  return FUS_OK;
}

static int peer_wait(Halo* h, bool fwd, double* a, double* b, cudaStream_t st) {
  const int nv = fwd ? 2 : 1, nn = (int)h->neigh.size();
  const OffTables T = tables(h);
  const long long n = fwd ? h->nrecv : h->nsend;
  const unsigned long long* epoch = h->d_epoch + (fwd ? 0 : 1);
  if (n == 0)
    return FUS_OK;
  const double* data = (const double*)(h->d_mbox + (fwd ? 0 : h->off_rev));
  const unsigned long long* flags
      = (const unsigned long long*)(h->d_mbox + (fwd ? h->off_fflag : h->off_rflag));
  PhaseScope ps(fwd ? 1 : 3, st);
  if (fwd)
    FUS_EMU_LAUNCH((peer_wait_kernel<false>), blocks_for(n), 256, 0, st, a, b, h->d_recv_idx, T.d_roff, nn, n,
                                                           nv, data, flags, epoch, h->d_error);
  else
    FUS_EMU_LAUNCH((peer_wait_kernel<true>), blocks_for(n), 256, 0, st, a, nullptr, h->d_send_idx, T.d_soff, nn,
                                                          n, 1, data, flags, epoch, h->d_error);
  FUS_CUDA_H(cudaGetLastError());
  return FUS_OK;
}

// mailbox = [fwd data: 2*nrecv doubles][rev data: nsend doubles][fwd flags][rev flags];
// layout4 = byte offsets {reverse data, forward flags, reverse flags, total size}
void halo_mailbox_layout(int64_t nsend, int64_t nrecv, int nneigh, int64_t* layout4) {
  const int64_t nn = std::max(1, nneigh);
  layout4[0] = (int64_t)sizeof(double) * 2 * std::max<int64_t>(1, nrecv);
  layout4[1] = layout4[0] + (int64_t)sizeof(double) * std::max<int64_t>(1, nsend);
  layout4[2] = layout4[1] + (int64_t)sizeof(unsigned long long) * nn;
  layout4[3] = layout4[2] + (int64_t)sizeof(unsigned long long) * nn;
}

int halo_peer_export(Halo* h, void* ipc_handle64, int64_t* layout3) {
  if (!h || !ipc_handle64 || !layout3)
    return FUS_ERR_ARG;
  if ((int)h->neigh.size() > kMaxNeigh) {
    set_error("peer transport supports at most %d neighbours", kMaxNeigh);
    return FUS_ERR_UNSUPPORTED;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  FUS_CUDA_H(cudaSetDevice(h->device));
  if (!h->d_mbox) {
    int64_t lay[4];
    halo_mailbox_layout(h->nsend, h->nrecv, (int)h->neigh.size(), lay);
    h->off_rev = (size_t)lay[0];
    h->off_fflag = (size_t)lay[1];
    h->off_rflag = (size_t)lay[2];
    h->mbox_bytes = (size_t)lay[3];
    FUS_CUDA_H(cudaMalloc(&h->d_mbox, h->mbox_bytes));
    FUS_CUDA_H(cudaMemset(h->d_mbox, 0, h->mbox_bytes));
    FUS_CUDA_H(cudaMalloc(&h->d_counter, 2 * sizeof(unsigned int)));
    FUS_CUDA_H(cudaMemset(h->d_counter, 0, 2 * sizeof(unsigned int)));
    FUS_CUDA_H(cudaMalloc(&h->d_epoch, 2 * sizeof(unsigned long long)));
    FUS_CUDA_H(cudaMemset(h->d_epoch, 0, 2 * sizeof(unsigned long long)));
    FUS_CUDA_H(cudaMalloc(&h->d_error, sizeof(int)));
    FUS_CUDA_H(cudaMemset(h->d_error, 0, sizeof(int)));
    FUS_CUDA_H(cudaDeviceSynchronize());
  }
  cudaIpcMemHandle_t hd;
  FUS_CUDA_H(cudaIpcGetMemHandle(&hd, h->d_mbox));
  std::memcpy(ipc_handle64, &hd, sizeof(hd));
  layout3[0] = (int64_t)h->off_rev;
  layout3[1] = (int64_t)h->off_fflag;
  layout3[2] = (int64_t)h->off_rflag;
  return FUS_OK;
}

// handles: one 64-byte IPC handle per neighbour (same order as the neighbour list);
// byte_off[k][4]: byte offsets inside neighbour k's mailbox of {my forward data segment, my forward
// flag, my reverse data segment, my reverse flag}.  The caller derives them from the layout triple
// that halo_peer_export returned on that neighbour and from its offset tables:
//   fwd data  = 8 * 2 * recv_off_q[j]            fwd flag = off_fflag_q + 8 * j
//   rev data  = off_rev_q + 8 * send_off_q[j]    rev flag = off_rflag_q + 8 * j
// with j = this rank's position in neighbour q's neighbour list.
int halo_peer_connect(Halo* h, const void* handles, const int64_t* byte_off) {
  if (!h || !h->d_mbox || (!h->neigh.empty() && (!handles || !byte_off))) {
    set_error("halo_peer_connect: export first, then pass the neighbours' handles and offsets");
    return FUS_ERR_ARG;
  }
  FUS_CUDA_H(cudaSetDevice(h->device));
  const size_t nn = h->neigh.size();
  PeerTable tab;
  std::memset(&tab, 0, sizeof(tab));
  h->peer_base.assign(nn, nullptr);
  for (size_t k = 0; k < nn; ++k) {
    cudaIpcMemHandle_t hd;
    std::memcpy(&hd, (const char*)handles + 64 * k, sizeof(hd));
    void* base = nullptr;
    FUS_CUDA_H(cudaIpcOpenMemHandle(&base, hd, cudaIpcMemLazyEnablePeerAccess));
    h->peer_base[k] = base;
    char* cb = (char*)base;
    tab.fwd_dst[k] = (double*)(cb + byte_off[4 * k + 0]);
    tab.fwd_flag[k] = (unsigned long long*)(cb + byte_off[4 * k + 1]);
    tab.rev_dst[k] = (double*)(cb + byte_off[4 * k + 2]);
    tab.rev_flag[k] = (unsigned long long*)(cb + byte_off[4 * k + 3]);
  }
  if (!h->d_tab)
    FUS_CUDA_H(cudaMalloc(&h->d_tab, sizeof(PeerTable)));
  FUS_CUDA_H(cudaMemcpy(h->d_tab, &tab, sizeof(tab), cudaMemcpyHostToDevice));
  h->peer = true;
  return FUS_OK;
}

int halo_peer_error(Halo* h) {
  if (!h || !h->d_error)
    return 0;
  int e = 0;
  if (cudaMemcpy(&e, h->d_error, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess)
    return 1;
  return e;
}

int halo_forward_begin(Halo* h, double* a, double* b, cudaStream_t st) {
  if (h->neigh.empty())
    return FUS_OK;
  if (h->peer) { // one-sided put on the side stream, concurrent with the next cells on `st`
    FUS_CUDA_H(cudaEventRecord(h->ev_fwd_ready, st));
    FUS_CUDA_H(cudaStreamWaitEvent(h->comm_stream, h->ev_fwd_ready, 0));
    int r = peer_put(h, true, a, b, h->comm_stream);
    if (r != FUS_OK)
      return r;
    FUS_CUDA_H(cudaEventRecord(h->ev_fwd_done, h->comm_stream));
    return FUS_OK;
  }
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
  if (h->peer) {
    // our own put has read a/b before anything later on `st` may overwrite them
    FUS_CUDA_H(cudaStreamWaitEvent(st, h->ev_fwd_done, 0));
    return peer_wait(h, true, a, b, st);
  }
  if (!h->overlap)
    return FUS_OK;
  FUS_CUDA_H(cudaStreamWaitEvent(st, h->ev_fwd_done, 0));
  return FUS_OK;
}

static int reverse_on(Halo* h, double* a, double* b, cudaStream_t st) {
  const int nv = b ? 2 : 1, nn = (int)h->neigh.size();
  const OffTables T = tables(h);
  if (h->nrecv)
    FUS_EMU_LAUNCH((halo_pack_kernel), blocks_for(h->nrecv), 256, 0, st, a, b, h->d_recv_idx, T.d_roff, nn,
                                                          h->d_sbuf, h->nrecv, nv);
  int r = exchange(h, false, nv, st);
  if (r != FUS_OK)
    return r;
  if (h->nsend)
    FUS_EMU_LAUNCH((halo_unpack_kernel<true>), blocks_for(h->nsend), 256, 0, st, a, b, h->d_send_idx, T.d_soff, nn, h->d_rbuf, h->nsend, nv);
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
  if (h->peer) {
    FUS_CUDA_H(cudaEventRecord(h->ev_ready, st));
    FUS_CUDA_H(cudaStreamWaitEvent(h->comm_stream, h->ev_ready, 0));
    int r = peer_put(h, false, a, nullptr, h->comm_stream);
    if (r != FUS_OK)
      return r;
    FUS_CUDA_H(cudaEventRecord(h->ev_done, h->comm_stream));
    return FUS_OK;
  }
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
  if (h->peer) {
    FUS_CUDA_H(cudaStreamWaitEvent(st, h->ev_done, 0)); // ghost partial sums have been read
    return peer_wait(h, false, a, nullptr, st);
  }
  if (!h->overlap)
    return reverse_on(h, a, nullptr, st);
  FUS_CUDA_H(cudaStreamWaitEvent(st, h->ev_done, 0));
  return FUS_OK;
}

} // namespace fus
