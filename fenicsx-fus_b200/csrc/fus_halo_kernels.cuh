// fus_halo_kernels.cuh -- device side of the halo exchange (fus_halo.cu).
//
//  * pack / unpack for the NCCL transport (ncclSend/ncclRecv per neighbour);
//  * the FUSED peer transport used inside fus_model_rk4: there is no exchange kernel at all.  The
//    two kernels of a stage do the exchange themselves over NVLink peer memory --
//      - the RK4 epilogue (rk4_stage_kernel<..., HALO>) handles the dofs this rank shares with its
//        neighbours FIRST: it adds the neighbours' partial sums of b (ghost -> owner, scatter_rev)
//        from its mailbox, and stores the next stage input it has just computed straight into the
//        neighbours' mailboxes (owner -> ghost, scatter_fwd), raising one flag per neighbour as soon
//        as those few blocks are done, while the rest of the epilogue is still streaming;
//        before it ends it makes sure the neighbours' forward data have landed (they were sent first
//        thing too), so the next operator starts without a wait;
//      - the stiffness kernel runs the cells that touch a shared dof first, in a launch of their
//        own (stiffness_line_kernel<..., HALO>): ghost values are gathered straight from the
//        mailbox, and after the last cell the warp groups copy the ghost part of b into the owners'
//        mailboxes in 2 KB chunks and raise the reverse flags; the rest of the mesh follows in the
//        plain kernel while the partial sums travel.
//    Sequence numbers live on the device, so one captured CUDA graph replays any step.
// Kept in a header so that tests/emu can run the very same code on host threads
// (FUS_HOST_EMULATION, see fus_kernels.cuh); device builds are unaffected.
#pragma once
#ifndef FUS_HOST_EMULATION
#include <cuda_runtime.h>
#endif

#include <cstdint>
#include <string>
#include <vector>

namespace fus {

constexpr int kMaxNeigh = 26;

// pack/unpack with the per-neighbour [vector][entry] interleave
static __global__ void __launch_bounds__(256)
    halo_pack_kernel(const double* __restrict__ a, const double* __restrict__ b,
                     const int32_t* __restrict__ idx, const int64_t* __restrict__ off, int nneigh,
                     double* __restrict__ buf, long long n, int nv) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  int k = 0;
  while (k + 1 < nneigh && i >= off[k + 1])
    ++k;
  const long long base = nv * off[k], len = off[k + 1] - off[k], j = i - off[k];
  const int d = idx[i];
  buf[base + j] = a[d];
  if (nv == 2)
    buf[base + len + j] = b[d];
}

template <bool ADD>
__global__ void __launch_bounds__(256)
    halo_unpack_kernel(double* __restrict__ a, double* __restrict__ b,
                       const int32_t* __restrict__ idx, const int64_t* __restrict__ off,
                       int nneigh, const double* __restrict__ buf, long long n, int nv) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  int k = 0;
  while (k + 1 < nneigh && i >= off[k + 1])
    ++k;
  const long long base = nv * off[k], len = off[k + 1] - off[k], j = i - off[k];
  const int d = idx[i];
  if (ADD) {
    atomicAdd(a + d, buf[base + j]); // an owned dof may be a ghost on several neighbours
    if (nv == 2)
      atomicAdd(b + d, buf[base + len + j]);
  } else {
    a[d] = buf[base + j];
    if (nv == 2)
      b[d] = buf[base + len + j];
  }
}

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
#ifdef FUS_HOST_EMULATION
  std::atomic_ref<unsigned long long>(*p).store(v, std::memory_order_release);
#else
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
#endif
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
#ifdef FUS_HOST_EMULATION
  return std::atomic_ref<unsigned long long>(*const_cast<unsigned long long*>(p))
      .load(std::memory_order_acquire);
#else
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
#endif
}
__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
#ifdef FUS_HOST_EMULATION
  return std::atomic_ref<unsigned int>(*const_cast<unsigned int*>(p)).load(std::memory_order_acquire);
#else
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
#endif
}
__device__ __forceinline__ unsigned long long halo_time_ns() {
#ifdef FUS_HOST_EMULATION
  return (unsigned long long)clock64();
#else
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
#endif
}

// ------------------------------------------------------------------------------------------------
// State of the fused peer transport of one rank (lives in device memory; built by halo_peer_connect).
// Mailbox of a rank = [fwd_u: nghost][fwd_v: nghost][rev: nsend][fwd flags][rev flags][ready flags];
// neighbour k writes its own segments of it: the ghosts it owns are the contiguous run
// nowned + roff[k] .. nowned + roff[k+1], the reverse data of send-list segment k sits at
// soff[k] .. soff[k+1].  Requirements on the local numbering (checked at connect time):
//   (P1) the owned dofs that appear in the send lists are exactly [0, nshared);
//   (P2) recv_idx == nowned + (0, 1, 2, ...): every neighbour's ghosts are one contiguous run.
// Sequence numbers: exchange number e of a direction is raised as flag value e by the sender and
// awaited as `flag >= e` by the receiver.  Every kernel reads the number it waits for from `seq`
// and only the kernel BEFORE it in stream order has written it (no kernel reads what it advances).
// ------------------------------------------------------------------------------------------------
enum HaloSeq { SEQ_FWD_SENT = 0, SEQ_FWD_EXPECT = 1, SEQ_REV_SENT = 2, SEQ_REV_EXPECT = 3,
               SEQ_CALLS = 4, SEQ_COUNT = 8 };
enum HaloCtr { CTR_GROUPS_PAST = 0, CTR_NEXT_CHUNK = 1, CTR_CHUNKS_DONE = 2, CTR_SHARED_DONE = 3,
               CTR_EPI_NEXT = 4, CTR_COUNT = 8 };
constexpr int kRevChunk = 256; // ghost entries of b copied per helper step (2 KB)

struct FusedHalo {
  int nneigh;
  long long nowned, nghost, nshared, nsend;
  // this rank's mailbox (local memory, written by the neighbours)
  const double* fwd_u;            // ghost values of the stage input, slot = dof - nowned
  const double* fwd_v;
  const double* rev;              // neighbours' partial sums of b, in send-list order
  const unsigned long long* fwd_flag;   // [nneigh]
  const unsigned long long* rev_flag;   // [nneigh]
  const unsigned long long* ready_flag; // [nneigh]
  // the neighbours' mailboxes (peer memory over NVLink): where MY segments go
  double* r_fwd_u[kMaxNeigh];
  double* r_fwd_v[kMaxNeigh];
  double* r_rev[kMaxNeigh];
  unsigned long long* r_fwd_flag[kMaxNeigh];
  unsigned long long* r_rev_flag[kMaxNeigh];
  unsigned long long* r_ready_flag[kMaxNeigh];
  // lists (device copies)
  const int64_t* soff;            // [nneigh + 1] send-list segments
  const int64_t* roff;            // [nneigh + 1] ghost segments
  const int32_t* sidx;            // [nsend] the send list: owned dof of every position
  const int32_t* spos_off;        // [nshared + 1] CSR over shared dofs: positions in the send list
  const int32_t* spos;            // [nsend] position i  <->  neighbour spos_nb, entry i - soff[nb]
  const signed char* spos_nb;     // [nsend]
  unsigned long long* seq;        // [SEQ_COUNT]
  unsigned int* ctr;              // [CTR_COUNT]
  int* error;                     // set when a wait timed out: every later kernel returns at once
  unsigned long long timeout_ns;
};

// What a HALO launch of the stiffness kernel needs per cell and per gather, passed BY VALUE as a
// kernel parameter: it then lives in the constant bank and costs the cell loop no registers.
struct HaloLaunch {
  const FusedHalo* H;   // nullptr outside HALO launches
  const double* mbu;    // fwd_u - nowned: ghost dof d of the first gathered vector is mbu[d]
  const double* mbv;    // same for the second gathered vector
  long long nown;       // owned dofs
};

// Host side: checks (P1) and (P2) on the lists given to fus_halo_setup and builds the CSR over the
// shared dofs (which positions of the send list carry dof s, and to which neighbour they go).
// Returns an empty string, or what is wrong with the numbering.
inline std::string fused_halo_lists(int64_t nowned, int nneigh, const int64_t* send_off,
                                    const int32_t* sidx, const int32_t* ridx, int64_t nrecv,
                                    int64_t* nshared, std::vector<int32_t>& spos_off,
                                    std::vector<int32_t>& spos, std::vector<signed char>& spos_nb) {
  const int64_t nsend = nneigh ? send_off[nneigh] : 0;
  for (int64_t i = 0; i < nrecv; ++i)
    if (ridx[i] != (int32_t)(nowned + i))
      return "ghosts must be numbered neighbour by neighbour in receive-list order (recv_idx["
             + std::to_string(i) + "] = " + std::to_string(ridx[i]) + ", expected "
             + std::to_string(nowned + i) + ")";
  int64_t mx = -1;
  for (int64_t i = 0; i < nsend; ++i)
    mx = sidx[i] > mx ? sidx[i] : mx;
  std::vector<char> seen((size_t)(mx + 1), 0);
  int64_t distinct = 0;
  for (int64_t i = 0; i < nsend; ++i)
    if (!seen[(size_t)sidx[i]]) {
      seen[(size_t)sidx[i]] = 1;
      ++distinct;
    }
  if (distinct != mx + 1)
    return "the dofs shared with neighbours must be numbered first (the send lists hold "
           + std::to_string(distinct) + " distinct dofs, the largest is " + std::to_string(mx) + ")";
  *nshared = distinct;
  spos_off.assign((size_t)distinct + 1, 0);
  spos.assign((size_t)nsend, 0);
  spos_nb.assign((size_t)nsend, 0);
  for (int64_t i = 0; i < nsend; ++i)
    ++spos_off[(size_t)sidx[i] + 1];
  for (int64_t d = 0; d < distinct; ++d)
    spos_off[(size_t)d + 1] += spos_off[(size_t)d];
  std::vector<int32_t> fill(spos_off.begin(), spos_off.end() - 1);
  for (int k = 0; k < nneigh; ++k)
    for (int64_t i = send_off[k]; i < send_off[k + 1]; ++i) {
      const int32_t at = fill[(size_t)sidx[i]]++;
      spos[(size_t)at] = (int32_t)i;
      spos_nb[(size_t)at] = (signed char)k;
    }
  return std::string();
}

// bounded wait for *flag >= want; false (and *error = 1) on time-out
__device__ __forceinline__ bool halo_wait_flag(const unsigned long long* flag, unsigned long long want,
                                               unsigned long long timeout_ns, int* error) {
  if (ld_acquire_sys(flag) >= want)
    return true;
  const unsigned long long t0 = halo_time_ns();
  for (;;) {
#pragma unroll 1
    for (int spin = 0; spin < 64; ++spin) {
      if (ld_acquire_sys(flag) >= want)
        return true;
      __nanosleep(40);
    }
    if (halo_time_ns() - t0 > timeout_ns || *(volatile int*)error) {
      atomicExch(error, 1);
      return false;
    }
  }
}

// one thread waits for every neighbour that sends to this rank in the given direction
__device__ __forceinline__ bool halo_wait_all(const FusedHalo& H, bool forward) {
  const unsigned long long want = H.seq[forward ? SEQ_FWD_EXPECT : SEQ_REV_EXPECT];
  const int64_t* off = forward ? H.roff : H.soff; // forward data fills my ghost runs
  const unsigned long long* flags = forward ? H.fwd_flag : H.rev_flag;
  bool ok = true;
  for (int k = 0; k < H.nneigh && ok; ++k)
    if (off[k + 1] > off[k])
      ok = halo_wait_flag(flags + k, want, H.timeout_ns, H.error);
  return ok;
}

// One thread: the forward exchange the next operator relies on is number seq[FWD_EXPECT] + 1; wait
// until every owner of this rank's ghosts has raised it, then publish the number.
__device__ __forceinline__ bool halo_forward_landed(const FusedHalo& H) {
  const unsigned long long want = H.seq[SEQ_FWD_EXPECT] + 1ull;
  bool ok = true;
  for (int k = 0; k < H.nneigh && ok; ++k)
    if (H.roff[k + 1] > H.roff[k])
      ok = halo_wait_flag(H.fwd_flag + k, want, H.timeout_ns, H.error);
  H.seq[SEQ_FWD_EXPECT] = want;
  return ok;
}

// The same two steps with one thread per neighbour (called by ALL threads of a block, or of a warp
// group with its own barrier): waiting for, or raising, seven flags one after the other costs seven
// NVLink round trips in a row, and both sit on the critical path of a stage.
template <typename Sync>
__device__ __forceinline__ bool halo_wait_parallel(const FusedHalo& H, bool forward,
                                                   unsigned long long want, int lane, int* word,
                                                   Sync sync) {
  if (lane == 0)
    *word = 1;
  sync();
  const int64_t* off = forward ? H.roff : H.soff; // forward data fill my ghost runs
  const unsigned long long* flags = forward ? H.fwd_flag : H.rev_flag;
  if (lane < H.nneigh && off[lane + 1] > off[lane])
    if (!halo_wait_flag(flags + lane, want, H.timeout_ns, H.error))
      *word = 0;
  sync();
  const bool ok = *word != 0;
  sync();
  return ok;
}

// all data stores have been fenced system-wide by their writers and ordered before this call
template <typename Sync>
__device__ __forceinline__ void halo_raise_parallel(const FusedHalo& H, bool forward, int lane,
                                                    unsigned long long* word, Sync sync) {
  if (lane == 0) {
    unsigned long long* sent = H.seq + (forward ? SEQ_FWD_SENT : SEQ_REV_SENT);
    *word = *sent + 1ull;
    *sent = *word;
  }
  sync();
  const int64_t* off = forward ? H.soff : H.roff;
  if (lane < H.nneigh && off[lane + 1] > off[lane]) {
    __threadfence_system();
    st_release_sys(forward ? H.r_fwd_flag[lane] : H.r_rev_flag[lane], *word);
  }
  sync();
}

// Raise this rank's flag of one direction on every neighbour it has sent to (one thread; all the
// data stores have been fenced system-wide and ordered before this call by the caller).
__device__ __forceinline__ void halo_raise(const FusedHalo& H, bool forward) {
  unsigned long long* sent = H.seq + (forward ? SEQ_FWD_SENT : SEQ_REV_SENT);
  const unsigned long long e = *sent + 1ull;
  *sent = e;
  const int64_t* off = forward ? H.soff : H.roff;
  __threadfence_system();
  for (int k = 0; k < H.nneigh; ++k)
    if (off[k + 1] > off[k])
      st_release_sys(forward ? H.r_fwd_flag[k] : H.r_rev_flag[k], e);
}

// ------------------------------------------------------------------------------------------------
// Entry of an rk4 call: (1) handshake -- every rank tells its neighbours that it has entered call
// number c and waits until they all have (the counterpart of the implicit synchronisation of the
// reference's first scatter_fwd: a neighbour that is late simply makes this rank wait, it cannot make
// it compute on stale ghosts; bounded by the configurable time-out); (2) the owner -> ghost update of
// the state (u_n, v_n) with the same protocol the epilogues use afterwards.
// phase: 1 = signal, 2 = wait, 3 = both (the emulation runs the phases of all ranks one by one).
// ------------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(32) halo_ready_kernel(const FusedHalo* Hp, int phase) {
  const FusedHalo& H = *Hp;
  if (threadIdx.x != 0 || blockIdx.x != 0)
    return;
  if (phase & 1) {
    const unsigned long long c = H.seq[SEQ_CALLS] + 1ull;
    H.seq[SEQ_CALLS] = c;
    for (int q = 0; q < CTR_COUNT; ++q)
      H.ctr[q] = 0u;
    __threadfence_system();
    for (int k = 0; k < H.nneigh; ++k)
      st_release_sys(H.r_ready_flag[k], c);
  }
  if (phase & 2) {
    const unsigned long long c = H.seq[SEQ_CALLS];
    for (int k = 0; k < H.nneigh; ++k)
      if (!halo_wait_flag(H.ready_flag + k, c, H.timeout_ns, H.error))
        return;
  }
}

static __global__ void __launch_bounds__(256)
    halo_entry_put_kernel(const FusedHalo* Hp, const double* __restrict__ u,
                          const double* __restrict__ v, int defer_wait) {
  const FusedHalo& H = *Hp;
  if (*(volatile int*)H.error)
    return;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < H.nsend) {
    int k = 0;
    while (k + 1 < H.nneigh && i >= H.soff[k + 1])
      ++k;
    const long long j = i - H.soff[k];
    const int d = H.sidx[i];
    H.r_fwd_u[k][j] = u[d];
    H.r_fwd_v[k][j] = v[d];
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (atomicAdd(H.ctr + CTR_SHARED_DONE, 1u) == gridDim.x - 1) {
      H.ctr[CTR_SHARED_DONE] = 0u;
      __threadfence();
      halo_raise(H, true);
      if (!defer_wait)
        halo_forward_landed(H); // the first operator of the call gathers without waiting
    }
  }
}

// A rank without a single interface cell launches no HALO stiffness kernel: this does its
// bookkeeping (the exchange numbers advance once per operator application on every rank).
static __global__ void __launch_bounds__(32) halo_operator_skipped_kernel(const FusedHalo* Hp) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    Hp->seq[SEQ_REV_EXPECT] += 1ull;
    Hp->ctr[CTR_SHARED_DONE] = 0u;
    Hp->ctr[CTR_EPI_NEXT] = 0u;
    halo_raise(*Hp, false);
  }
}

// tests/emu only (ranks are emulated one after another there, so a rank cannot wait inside its
// epilogue for a neighbour whose epilogue has not run yet): the closing wait of an epilogue / entry put
static __global__ void __launch_bounds__(32) halo_forward_landed_kernel(const FusedHalo* Hp) {
  if (threadIdx.x == 0 && blockIdx.x == 0)
    halo_forward_landed(*Hp);
}

// Exit of an rk4 call: the last epilogue has sent the new state to the ghosts' mailboxes; bring it
// into the ghost entries of (u_n, v_n) so that they leave with fresh ghosts (Linear.hpp:312-313).
static __global__ void __launch_bounds__(256)
    halo_exit_unpack_kernel(const FusedHalo* Hp, double* __restrict__ u, double* __restrict__ v) {
  const FusedHalo& H = *Hp;
  if (*(volatile int*)H.error)
    return;
  // the last epilogue has waited for the neighbours' forward data: nothing to wait for here
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < H.nghost) {
    u[H.nowned + i] = __ldcg(H.fwd_u + i);
    v[H.nowned + i] = __ldcg(H.fwd_v + i);
  }
}

} // namespace fus
