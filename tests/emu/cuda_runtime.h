// cuda_runtime.h -- TEST INFRASTRUCTURE: a host stand-in for the slice of the CUDA runtime API that
// fus_capi.cu / fus_halo.cu use, so that the whole library (host plumbing + kernels through
// simt_emu.hpp) can be built as a CPU-only shared object and driven by the same Python tests that
// run on a B200 (tests/emu/build_emulated_library.py).  Everything is synchronous: streams and
// events are no-ops, "device" memory is host memory, IPC handles carry raw pointers (one process),
// graph capture reports failure so that the library stays on eager issue.  Found through the include
// path of the emulation build only; never part of the product.
#pragma once
#include "simt_emu.hpp"

#include <chrono>
#include <cstdlib>
#include <cstring>

typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;
constexpr cudaError_t cudaErrorEmulated = 999;
typedef struct emu_stream* cudaStream_t;
typedef struct emu_graph* cudaGraph_t;
typedef struct emu_graph_exec* cudaGraphExec_t;
struct emu_event {
  std::chrono::steady_clock::time_point t;
};
typedef emu_event* cudaEvent_t;

enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaStreamCaptureMode { cudaStreamCaptureModeThreadLocal = 1 };
enum cudaLimit { cudaLimitPersistingL2CacheSize = 6 };
enum cudaStreamAttrID { cudaStreamAttributeAccessPolicyWindow = 1 };
enum cudaAccessProperty { cudaAccessPropertyStreaming = 1, cudaAccessPropertyPersisting = 2 };
struct cudaAccessPolicyWindow {
  void* base_ptr;
  size_t num_bytes;
  float hitRatio;
  cudaAccessProperty hitProp, missProp;
};
union cudaStreamAttrValue {
  cudaAccessPolicyWindow accessPolicyWindow;
};
struct cudaIpcMemHandle_t {
  char reserved[64];
};
struct cudaDeviceProp {
  int major = 10, minor = 0;
  int multiProcessorCount = 2; // small grids: few emulated blocks per launch
  int persistingL2CacheMaxSize = 0, accessPolicyMaxWindowSize = 0;
};

inline const char* cudaGetErrorString(cudaError_t e) { return e ? "emulated CUDA error" : "no error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) {
  *n = 1;
  return cudaSuccess;
}
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
  *p = cudaDeviceProp();
  return cudaSuccess;
}
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
inline cudaError_t cudaDeviceSetLimit(cudaLimit, size_t) { return cudaSuccess; }
inline cudaError_t cudaCtxResetPersistingL2Cache() { return cudaSuccess; }
inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) {
  *lo = 0;
  *hi = -1;
  return cudaSuccess;
}

template <typename T>
inline cudaError_t cudaMalloc(T** p, size_t bytes) {
  *p = static_cast<T*>(std::calloc(bytes ? bytes : 1, 1));
  return *p ? cudaSuccess : cudaErrorEmulated;
}
inline cudaError_t cudaFree(void* p) {
  std::free(p);
  return cudaSuccess;
}
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) {
  std::memmove(d, s, n);
  return cudaSuccess;
}
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind k, cudaStream_t = nullptr) {
  return cudaMemcpy(d, s, n, k);
}
inline cudaError_t cudaMemset(void* d, int v, size_t n) {
  std::memset(d, v, n);
  return cudaSuccess;
}
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) {
  return cudaMemset(d, v, n);
}

inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) {
  *s = reinterpret_cast<cudaStream_t>(new char);
  return cudaSuccess;
}
inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned f, int) {
  return cudaStreamCreateWithFlags(s, f);
}
inline cudaError_t cudaStreamDestroy(cudaStream_t s) {
  delete reinterpret_cast<char*>(s);
  return cudaSuccess;
}
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return cudaSuccess; }
inline cudaError_t cudaStreamSetAttribute(cudaStream_t, cudaStreamAttrID, const cudaStreamAttrValue*) {
  return cudaSuccess;
}
// graph capture is not emulated: the library falls back to eager issue (fus_model_rk4)
inline cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode) { return cudaErrorEmulated; }
inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* g) {
  *g = nullptr;
  return cudaErrorEmulated;
}
inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t, unsigned long long) {
  *e = nullptr;
  return cudaErrorEmulated;
}
inline cudaError_t cudaGraphLaunch(cudaGraphExec_t, cudaStream_t) { return cudaErrorEmulated; }
inline cudaError_t cudaGraphDestroy(cudaGraph_t) { return cudaSuccess; }
inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t) { return cudaSuccess; }

inline cudaError_t cudaEventCreate(cudaEvent_t* e) {
  *e = new emu_event();
  return cudaSuccess;
}
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
inline cudaError_t cudaEventDestroy(cudaEvent_t e) {
  delete e;
  return cudaSuccess;
}
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t = nullptr) {
  e->t = std::chrono::steady_clock::now();
  return cudaSuccess;
}
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
  *ms = std::chrono::duration<float, std::milli>(b->t - a->t).count();
  return cudaSuccess;
}

template <typename F>
inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) {
  return cudaSuccess;
}
template <typename F>
inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, F, int, size_t) {
  *n = 1;
  return cudaSuccess;
}

// one process: a handle is the pointer itself
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) {
  std::memset(h, 0, sizeof(*h));
  std::memcpy(h->reserved, &p, sizeof(p));
  return cudaSuccess;
}
inline cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) {
  std::memcpy(p, h.reserved, sizeof(*p));
  return cudaSuccess;
}
inline cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }

// kernel<<<grid, block, smem, stream>>>(args...) is rewritten by the emulation build into
// FUS_EMU_LAUNCH(kernel, grid, block, smem, stream, args...)
#define FUS_EMU_LAUNCH(kernel, grid, block, smem, stream, ...)                                    \
  fus_emu::launch((unsigned)(grid), (unsigned)(block), (size_t)(smem), [&] { kernel(__VA_ARGS__); })
