// Shared helpers for the dhd_b200 sm_100a kernels.
#pragma once
#include <atomic>
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

extern std::atomic<long> g_launches;  // kernels enqueued by this library since load (bench bookkeeping; any host thread)

#define DHD_CUDA_LAUNCH_CHECK(name)                                               \
  do {                                                                            \
    ::dhd::g_launches.fetch_add(1, std::memory_order_relaxed);                    \
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

// ---- class index of predictor.get_occ (occ_head.py:141-153): `occ_pred.softmax(-1).argmax(-1)` ----------------
// softmax is monotone, so this is the argmax of the logits -- except where fp32 rounding of exp / the division maps
// two DIFFERENT logits onto the SAME probability: torch then returns the lower class index.  That needs the top two
// logits within ~2e-7 of each other; callers run the plain argmax (first maximum wins) while tracking the runner-up
// and fall back to this restatement of torch's kernel when `best - second <= kSoftmaxTieGap`.
// torch (aten/src/ATen/native/cuda/PersistentSoftmax.cuh, softmax_warp_forward, n <= 32: one element per lane):
// m = max x; e_k = expf(x_k - m); s = butterfly sum over 32 lanes (masked lanes hold exp(-inf) = 0);
// p_k = e_k / s; then argmax with the lower index winning ties.  The butterfly adds lane l and lane l ^ offset for
// offset = W/2 .. 1, i.e. the balanced tree below (fp32 addition is commutative, so every lane holds the same sum).
constexpr float kSoftmaxTieGap = 1e-6f;

template <int N>
__device__ __forceinline__ int softmax_argmax_torch(const float (&x)[N]) {
  static_assert(N >= 1 && N <= 32, "one element per lane");
  constexpr int W = N <= 1 ? 1 : N <= 2 ? 2 : N <= 4 ? 4 : N <= 8 ? 8 : N <= 16 ? 16 : 32;
  float m = x[0];
#pragma unroll
  for (int k = 1; k < N; ++k) m = fmaxf(m, x[k]);
  float e[W], t[W];
#pragma unroll
  for (int k = 0; k < W; ++k) e[k] = t[k] = k < N ? expf(x[k] - m) : 0.f;
#pragma unroll
  for (int off = W / 2; off >= 1; off >>= 1)
#pragma unroll
    for (int l = 0; l < off; ++l) t[l] = __fadd_rn(t[l], t[l + off]);
  const float s = t[0];
  float best = __fdiv_rn(e[0], s);
  int arg = 0;
#pragma unroll
  for (int k = 1; k < N; ++k) {
    const float p = __fdiv_rn(e[k], s);
    if (p > best) {
      best = p;
      arg = k;
    }
  }
  return arg;
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

// 16-byte streaming load through the read-only path without polluting L1
__device__ __forceinline__ uint4 ld_nc_u4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

}  // namespace dhd
