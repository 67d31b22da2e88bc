// Shared helpers for the dhd_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/dhd_b200.h"

namespace dhd {

extern thread_local char g_err[512];

inline int fail(int code, const char* fmt, const char* a = "", long b = 0, long c = 0) {
  snprintf(g_err, sizeof(g_err), fmt, a, b, c);
  return code;
}

#define DHD_REQUIRE(cond, msg)                                                    \
  do {                                                                            \
    if (!(cond)) return ::dhd::fail(DHD_EINVAL, "%s (" #cond ")", msg);            \
  } while (0)

extern long g_launches;  // kernels enqueued by this library since load (bench bookkeeping)

#define DHD_CUDA_LAUNCH_CHECK(name)                                               \
  do {                                                                            \
    ++::dhd::g_launches;                                                          \
    cudaError_t e__ = cudaGetLastError();                                         \
    if (e__ != cudaSuccess)                                                       \
      return ::dhd::fail((int)e__, "%s launch failed: %ld", name, (long)e__);      \
  } while (0)

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}

// streaming (evict-first) 64-bit / 128-bit stores for write-once outputs
__device__ __forceinline__ void st_cs(float2* p, float2 v) {
  asm volatile("st.global.cs.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void st_cs(float4* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_cs(float* p, float v) {
  asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

// a / b and a % b for a flat element index: 32-bit unsigned division whenever the index fits (a 64-bit division is
// ~100 instructions, more than the rest of a streaming kernel's thread)
__device__ __forceinline__ long fast_div(long a, int b, int* rem) {
  if (a < (1L << 32)) {
    const unsigned u = (unsigned)a, q = u / (unsigned)b;
    *rem = (int)(u - q * (unsigned)b);
    return (long)q;
  }
  *rem = (int)(a % b);
  return a / b;
}
__device__ __forceinline__ long fast_div(long a, int b) {
  int r;
  return fast_div(a, b, &r);
}

inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

}  // namespace dhd
