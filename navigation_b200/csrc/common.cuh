// common.cuh -- shared host/device helpers of libnavgpu (sm_100a only).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <utility>

#include "../../include/navgpu.h"

namespace navgpu {

// ---- error plumbing: the C ABI never throws; the last message is kept per thread -------------------------------
inline std::string& last_error_ref() {
  static thread_local std::string s;
  return s;
}
inline int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  last_error_ref() = buf;
  return code;
}

#define NAVGPU_CUDA(call)                                                                                      \
  do {                                                                                                         \
    cudaError_t e_ = (call);                                                                                   \
    if (e_ != cudaSuccess)                                                                                     \
      return ::navgpu::fail(NAVGPU_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                            __LINE__);                                                                         \
  } while (0)

#define NAVGPU_TRY(expr)           \
  do {                             \
    int rc_ = (expr);              \
    if (rc_ != NAVGPU_OK) return rc_; \
  } while (0)

// Programmatic dependent launch: the kernel may be scheduled while its predecessor in the stream is still draining;
// it must call cudaGridDependencySynchronize() before touching anything the predecessor wrote.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

inline std::atomic<uint64_t>& launch_counter() {
  static std::atomic<uint64_t> c{0};
  return c;
}
#define NAVGPU_LAUNCHED(n) ::navgpu::launch_counter().fetch_add((n), std::memory_order_relaxed)

// Every device grid uses a row pitch that is a multiple of 128 bytes so rows start on a 128-B line and 8/16-byte
// vector accesses are always aligned.
inline uint32_t grid_pitch(uint32_t size_x) { return (size_x + 127u) & ~127u; }

enum : uint8_t { kFree = 0, kInscribed = 253, kLethal = 254, kNoInfo = 255 };  // cost_values.h:42-45

// ---- monotone encoding of doubles for atomicMin/atomicMax on bounds ---------------------------------------------
__host__ __device__ inline unsigned long long enc_double(double d) {
  unsigned long long u;
#ifdef __CUDA_ARCH__
  u = (unsigned long long)__double_as_longlong(d);
#else
  memcpy(&u, &d, 8);
#endif
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__host__ __device__ inline double dec_double(unsigned long long u) {
  u = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)u);
#else
  double d;
  memcpy(&d, &u, 8);
  return d;
#endif
}

// per-layer bounding box accumulated on the device by the ray-trace / marking kernels (CostmapLayer::touch)
struct DevBox {
  unsigned long long minx, miny, maxx, maxy;  // enc_double
};

// the update window of the current cycle, produced on the device by k_finalize_bounds
struct DevWindow {
  int x0, xn, y0, yn;  // LayeredCostmap::bx0_, bxn_, by0_, byn_
  int valid;           // 0 when xn < x0 || yn < y0 (updateMap returns before resetMap, layered_costmap.cpp:128-135)
  int pad_[3];
};

}  // namespace navgpu
