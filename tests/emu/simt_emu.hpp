// simt_emu.hpp -- TEST INFRASTRUCTURE: a small SIMT emulator for the kernels in
// fenicsx-fus_b200/csrc/fus_kernels.cuh.
//
// The build container has no GPU.  To keep the kernels' *logic* (thread-to-data mapping, shared
// memory staging, software pipelining, barriers, tail handling) under test on the CPU, this header
// lets the very same kernel source compile as host C++: every CUDA thread of a block runs as an OS
// thread, __syncthreads / __syncwarp / named barriers are real barriers, atomics are atomics, blocks
// run one after another.  Threads of a warp are NOT in lockstep here, so code that silently relies
// on warp-synchronous execution without a __syncwarp shows up as a race.  What it cannot show:
// anything about the memory model of the device, performance, or PTX-level behaviour.
// Only tests/ may use it; the product never loads it (there is no CPU fallback).
#pragma once
#define FUS_HOST_EMULATION 1

#include <algorithm>
#include <array>
#include <atomic>
#include <barrier>
#include <chrono>
#include <cstdlib>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

struct double2 {
  double x, y;
};
inline double2 make_double2(double a, double b) { return double2{a, b}; }
struct float2 {
  float x, y;
};

struct emu_dim3 {
  unsigned x = 1, y = 1, z = 1;
};
inline thread_local emu_dim3 threadIdx, blockIdx;
inline emu_dim3 blockDim, gridDim; // one launch at a time

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __grid_constant__
#define __launch_bounds__(...)
#define __shared__ static // blocks run one after another, so one copy per kernel instantiation

namespace fus_emu {
inline void jitter();
struct BlockState {
  std::unique_ptr<std::barrier<>> block;
  std::vector<std::unique_ptr<std::barrier<>>> warps;
  std::map<int, std::unique_ptr<std::barrier<>>> named;
  std::mutex mu;
  std::vector<double> dyn;
  std::vector<std::array<unsigned long long, 32>> shfl; // one exchange row per warp
};
inline BlockState* g_block = nullptr;

inline double* dynamic_shared() { return g_block->dyn.data(); }
inline void named_barrier(int id, int count) {
  std::barrier<>* b;
  {
    std::lock_guard<std::mutex> lk(g_block->mu);
    auto& slot = g_block->named[id];
    if (!slot)
      slot = std::make_unique<std::barrier<>>(count);
    b = slot.get();
  }
  b->arrive_and_wait();
  jitter();
}

// kernel<<<grid, block, smem_bytes>>>(args...)  ->  launch(grid, block, smem_bytes, [&]{ kernel(args...); })
inline void launch(unsigned grid, unsigned block, size_t smem_bytes, const std::function<void()>& body) {
  gridDim.x = grid;
  blockDim.x = block;
  for (unsigned b = 0; b < grid; ++b) {
    BlockState st;
    st.block = std::make_unique<std::barrier<>>(block);
    for (unsigned w = 0; w < (block + 31) / 32; ++w)
      st.warps.push_back(std::make_unique<std::barrier<>>(std::min(32u, block - 32 * w)));
    st.dyn.assign(smem_bytes / sizeof(double) + 1, 0.0);
    st.shfl.resize((block + 31) / 32);
    g_block = &st;
    std::vector<std::thread> th;
    th.reserve(block);
    for (unsigned t = 0; t < block; ++t)
      th.emplace_back([&, t, b] {
        threadIdx.x = t;
        blockIdx.x = b;
        body();
      });
    for (auto& x : th)
      x.join();
    g_block = nullptr;
  }
}
} // namespace fus_emu

// FUS_EMU_JITTER=<microseconds>: every thread sleeps a pseudo-random time below that bound after
// each barrier, so that threads run far apart and a missing barrier is very likely to be observed.
namespace fus_emu {
inline void jitter() {
  static const int bound = [] {
    const char* e = std::getenv("FUS_EMU_JITTER");
    return e ? std::atoi(e) : 0;
  }();
  if (bound > 0) {
    thread_local unsigned state = 2654435761u * (threadIdx.x + 1) + blockIdx.x;
    state = state * 1664525u + 1013904223u;
    std::this_thread::sleep_for(std::chrono::microseconds((state >> 8) % (unsigned)bound));
  }
}
} // namespace fus_emu

inline void __syncthreads() {
  fus_emu::g_block->block->arrive_and_wait();
  fus_emu::jitter();
}
inline void __syncwarp() {
  fus_emu::g_block->warps[threadIdx.x / 32]->arrive_and_wait();
  fus_emu::jitter();
}
// warp shuffles: every lane of the warp must call (mask is not interpreted); values up to 8 bytes
namespace fus_emu {
template <typename T>
inline T shfl_from(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle of at most 64 bits");
  auto& row = g_block->shfl[threadIdx.x / 32];
  const int lane = threadIdx.x % 32;
  unsigned long long bits = 0;
  std::memcpy(&bits, &v, sizeof(T));
  row[lane] = bits;
  g_block->warps[threadIdx.x / 32]->arrive_and_wait();
  const unsigned long long got = row[(src_lane >= 0 && src_lane < 32) ? src_lane : lane];
  g_block->warps[threadIdx.x / 32]->arrive_and_wait();
  T out;
  std::memcpy(&out, &got, sizeof(T));
  return out;
}
} // namespace fus_emu
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int m) {
  return fus_emu::shfl_from(v, (int)(threadIdx.x % 32) ^ m);
}
template <typename T>
inline T __shfl_up_sync(unsigned, T v, int d) {
  const int lane = threadIdx.x % 32;
  return fus_emu::shfl_from(v, lane >= d ? lane - d : lane);
}
template <typename T>
inline T __shfl_down_sync(unsigned, T v, int d) {
  const int lane = threadIdx.x % 32;
  return fus_emu::shfl_from(v, lane + d < 32 ? lane + d : lane);
}
template <typename T>
inline T __ldg(const T* p) {
  return *p;
}
inline double atomicAdd(double* p, double v) { return std::atomic_ref<double>(*p).fetch_add(v); }
inline float atomicAdd(float* p, float v) { return std::atomic_ref<float>(*p).fetch_add(v); }
inline int atomicExch(int* p, int v) { return std::atomic_ref<int>(*p).exchange(v); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return std::atomic_ref<unsigned>(*p).fetch_add(v); }
// device-only intrinsics of the halo kernels
inline long long clock64() {
  return std::chrono::duration_cast<std::chrono::nanoseconds>(
             std::chrono::steady_clock::now().time_since_epoch())
      .count();
}
inline void __nanosleep(unsigned) { std::this_thread::yield(); }
inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
template <typename T>
inline T __ldcg(const T* p) {
  return *p;
}
using std::fabs;
using std::fma;
using std::fmax;
