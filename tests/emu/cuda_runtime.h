// cuda_runtime.h -- TEST INFRASTRUCTURE: a host stand-in for the slice of the CUDA runtime API that
// fus_capi.cu / fus_halo.cu use, so that the whole library (host plumbing + kernels through
// simt_emu.hpp) can be built as a CPU-only shared object and driven by the same Python tests that
// run on a B200 (tests/emu/build_emulated_library.py).  Everything is synchronous: streams and
// events are no-ops, "device" memory is host memory, IPC handles carry raw pointers (one process).
// Stream capture records the launches and asynchronous copies issued while it is active as closures
// (arguments by value, nothing runs) and cudaGraphLaunch replays them in order, so that the library's
// graph path -- capture of one RK4 step, replay, invalidation -- is exercised too; FUS_EMU_NO_GRAPH=1
// (or the null stream, as on a real device's legacy stream) makes capture fail and the library stay
// on eager issue.  Found through the include path of the emulation build only; never part of the
// product.
#pragma once
#include "simt_emu.hpp"

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <tuple>
#include <vector>

typedef int cudaError_t;
constexpr cudaError_t cudaSuccess = 0;
constexpr cudaError_t cudaErrorEmulated = 999;
constexpr cudaError_t cudaErrorPeerAccessAlreadyEnabled = 704;
typedef struct emu_stream* cudaStream_t;
struct emu_graph {
  std::vector<std::function<void()>> ops;
};
struct emu_graph_exec {
  std::vector<std::function<void()>> ops;
};
typedef emu_graph* cudaGraph_t;
typedef emu_graph_exec* cudaGraphExec_t;
namespace fus_emu {
// the graph being captured, if any (one capture at a time: the library captures on one stream and
// joins its side streams into it through events)
inline emu_graph*& capturing() {
  static emu_graph* g = nullptr;
  return g;
}
// run now, or record while a capture is active
template <typename F>
inline void submit(F&& op) {
  if (capturing())
    capturing()->ops.emplace_back(std::forward<F>(op));
  else
    op();
}
} // namespace fus_emu
struct emu_event {
  std::chrono::steady_clock::time_point t;
};
typedef emu_event* cudaEvent_t;

enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1, cudaEventDisableTiming = 2, cudaIpcMemLazyEnablePeerAccess = 1 };
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum cudaStreamCaptureMode { cudaStreamCaptureModeThreadLocal = 1 };
enum cudaStreamCaptureStatus { cudaStreamCaptureStatusNone = 0, cudaStreamCaptureStatusActive = 1 };
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
inline cudaError_t cudaDeviceCanAccessPeer(int* can, int, int) {
  *can = 1;
  return cudaSuccess;
}
inline cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
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
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) {
  fus_emu::submit([=] { std::memmove(d, s, n); });
  return cudaSuccess;
}
inline cudaError_t cudaMemset(void* d, int v, size_t n) {
  std::memset(d, v, n);
  return cudaSuccess;
}
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) {
  fus_emu::submit([=] { std::memset(d, v, n); });
  return cudaSuccess;
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
inline cudaError_t cudaStreamBeginCapture(cudaStream_t s, cudaStreamCaptureMode) {
  if (!s || fus_emu::capturing() || std::getenv("FUS_EMU_NO_GRAPH"))
    return cudaErrorEmulated; // the library then stays on eager issue (fus_model_rk4)
  fus_emu::capturing() = new emu_graph();
  return cudaSuccess;
}
inline cudaError_t cudaStreamIsCapturing(cudaStream_t, cudaStreamCaptureStatus* st) {
  *st = fus_emu::capturing() ? cudaStreamCaptureStatusActive : cudaStreamCaptureStatusNone;
  return cudaSuccess;
}
inline cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t* g) {
  *g = fus_emu::capturing();
  fus_emu::capturing() = nullptr;
  return *g ? cudaSuccess : cudaErrorEmulated;
}
inline cudaError_t cudaGraphInstantiate(cudaGraphExec_t* e, cudaGraph_t g, unsigned long long) {
  *e = new emu_graph_exec{g->ops};
  return cudaSuccess;
}
namespace fus_emu {
inline long long& graph_launches() { // replays so far (tests assert that the graph path was taken)
  static long long n = 0;
  return n;
}
} // namespace fus_emu
inline cudaError_t cudaGraphLaunch(cudaGraphExec_t e, cudaStream_t) {
  if (!e || fus_emu::capturing())
    return cudaErrorEmulated;
  for (auto& op : e->ops)
    op();
  ++fus_emu::graph_launches();
  return cudaSuccess;
}
inline cudaError_t cudaGraphDestroy(cudaGraph_t g) {
  delete g;
  return cudaSuccess;
}
inline cudaError_t cudaGraphExecDestroy(cudaGraphExec_t e) {
  delete e;
  return cudaSuccess;
}

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
// (arguments are evaluated and copied when the launch is issued: a captured launch runs later, from
// cudaGraphLaunch)
namespace fus_emu {
template <typename K, typename... A>
inline void submit_launch(unsigned grid, unsigned block, size_t smem, K call, A... args) {
  auto pack = std::make_tuple(args...);
  submit([=] { launch(grid, block, smem, [&] { std::apply(call, pack); }); });
}
} // namespace fus_emu
#define FUS_EMU_LAUNCH(kernel, grid, block, smem, stream, ...)                                    \
  fus_emu::submit_launch((unsigned)(grid), (unsigned)(block), (size_t)(smem),                     \
                         [=](auto... emu_a) { kernel(emu_a...); }, __VA_ARGS__)
