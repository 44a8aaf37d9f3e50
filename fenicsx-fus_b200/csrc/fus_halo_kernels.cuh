// fus_halo_kernels.cuh -- device side of the halo exchange (fus_halo.cu): pack / unpack for the
// NCCL transport, one-sided put / wait for the peer-direct transport.  Kept in a header so that
// tests/emu can run the very same kernels on host threads (FUS_HOST_EMULATION, see
// fus_kernels.cuh); device builds are unaffected.
#pragma once
#ifndef FUS_HOST_EMULATION
#include <cuda_runtime.h>
#endif

#include <cstdint>

namespace fus {

constexpr int kMaxNeigh = 26;
struct PeerTable {
  double* fwd_dst[kMaxNeigh];               // where my packed owner values go on neighbour k
  double* rev_dst[kMaxNeigh];               // where my ghost partial sums go on neighbour k
  unsigned long long* fwd_flag[kMaxNeigh];  // flag on neighbour k that I raise after a forward put
  unsigned long long* rev_flag[kMaxNeigh];
};

// pack/unpack with the per-neighbour [vector][entry] interleave
__global__ void __launch_bounds__(256)
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

__global__ void __launch_bounds__(256)
    peer_put_kernel(const double* __restrict__ a, const double* __restrict__ b,
                    const int32_t* __restrict__ idx, const int64_t* __restrict__ off, int nneigh,
                    long long n, int nv, const PeerTable* __restrict__ tab, int forward,
                    unsigned int* counter, unsigned long long* epoch_ctr, int lightfence) {
  // every block reads the counter before it can be advanced: the last block only advances it
  // after all blocks have passed their atomicAdd below
  const unsigned long long epoch = *(volatile unsigned long long*)epoch_ctr + 1ull;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    int k = 0;
    while (k + 1 < nneigh && i >= off[k + 1])
      ++k;
    const long long len = off[k + 1] - off[k], j = i - off[k];
    double* dst = forward ? tab->fwd_dst[k] : tab->rev_dst[k];
    const int d = idx[i];
    dst[j] = a[d];
    if (nv == 2)
      dst[len + j] = b[d];
  }
  if (!lightfence)
    __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    if (lightfence)
      __threadfence_system(); // cumulative over the block's stores ordered by the barrier
    const unsigned int prev = atomicAdd(counter, 1u);
    if (prev == gridDim.x - 1) { // every block's stores are fenced before its increment
      *counter = 0;
      *epoch_ctr = epoch;
      __threadfence_system();
      for (int k = 0; k < nneigh; ++k)
        if (off[k + 1] > off[k])
          st_release_sys(forward ? tab->fwd_flag[k] : tab->rev_flag[k], epoch);
    }
  }
}

template <bool ADD>
__global__ void __launch_bounds__(256)
    peer_wait_kernel(double* __restrict__ a, double* __restrict__ b,
                     const int32_t* __restrict__ idx, const int64_t* __restrict__ off, int nneigh,
                     long long n, int nv, const double* mbox_data,
                     const unsigned long long* flags, const unsigned long long* epoch_ctr,
                     int* error) {
  // the local put of this exchange is ordered before this kernel and has advanced the counter
  const unsigned long long epoch = *(volatile const unsigned long long*)epoch_ctr;
  if (threadIdx.x < nneigh && off[threadIdx.x + 1] > off[threadIdx.x]) {
    const long long t0 = clock64();
    while (ld_acquire_sys(flags + threadIdx.x) < epoch) {
      if (clock64() - t0 > 4000000000ll) { // ~2 s: a peer died; report instead of hanging the GPU
        atomicExch(error, 1);
        break;
      }
      __nanosleep(100);
    }
  }
  __syncthreads();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  int k = 0;
  while (k + 1 < nneigh && i >= off[k + 1])
    ++k;
  const long long base = nv * off[k], len = off[k + 1] - off[k], j = i - off[k];
  const int d = idx[i];
  const double va = __ldcg(mbox_data + base + j); // written by a peer: never trust L1
  if (ADD) {
    atomicAdd(a + d, va);
  } else {
    a[d] = va;
    if (nv == 2)
      b[d] = __ldcg(mbox_data + base + len + j);
  }
}

} // namespace fus
