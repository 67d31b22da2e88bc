// Write-bandwidth probes: what a write-only kernel can reach on this GPU for a given store
// instruction and address pattern.  Used by scripts/bench_pool.py and bench.py to put the fused
// pool (which writes 99 % of its bytes) next to a write-only ceiling, beside the copy peak in
// MEASURED_PEAKS.json.  Not on the product path.
#include "common.cuh"

namespace dhd {

__device__ __forceinline__ void probe_bulk(void* gdst, uint32_t ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc),
               "r"(bytes)
               : "memory");
}

// mode 0: grid-stride st.global.cs.v4   mode 1: grid-stride st.global.v4
// mode 2: every warp owns one contiguous run, st.cs.v4 (the pool's static split)
__global__ void __launch_bounds__(256) probe_write_kernel(float4* dst, size_t n16, int mode) {
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t nthr = (size_t)gridDim.x * blockDim.x;
  if (mode == 0) {
    for (size_t i = tid; i < n16; i += nthr) st_cs(dst + i, z);
  } else if (mode == 1) {
    for (size_t i = tid; i < n16; i += nthr) dst[i] = z;
  } else {
    const size_t warps = nthr / 32, w = tid / 32, lane = tid % 32;
    const size_t per = (n16 + warps - 1) / warps;
    const size_t lo = w * per, hi = lo + per < n16 ? lo + per : n16;
    for (size_t i = lo + lane; i < hi; i += 32) st_cs(dst + i, z);
  }
}

// mode 3: cp.async.bulk from a zero tile; every CTA owns one contiguous run and thread 0 issues
// `chunk`-byte copies.  mode 4: every warp owns a contiguous run, lane 0 issues.
__global__ void __launch_bounds__(256) probe_bulk_kernel(char* dst, size_t bytes, int chunk, int per_warp) {
  extern __shared__ __align__(128) uint8_t ztile[];
  for (int i = threadIdx.x; i < chunk / 16; i += blockDim.x)
    reinterpret_cast<float4*>(ztile)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const uint32_t zs = (uint32_t)__cvta_generic_to_shared(ztile);
  const bool issuer = per_warp ? (threadIdx.x & 31) == 0 : threadIdx.x == 0;
  if (issuer) {
    const size_t parts = per_warp ? (size_t)gridDim.x * (blockDim.x / 32) : gridDim.x;
    const size_t me = per_warp ? (size_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32 : blockIdx.x;
    size_t per = (bytes + parts - 1) / parts;
    per = (per + 255) / 256 * 256;
    size_t lo = me * per, hi = lo + per < bytes ? lo + per : bytes;
    while (lo < hi) {
      const uint32_t b = hi - lo < (size_t)chunk ? (uint32_t)(hi - lo) : (uint32_t)chunk;
      probe_bulk(dst + lo, zs, b);
      lo += b;
    }
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

}  // namespace dhd

using namespace dhd;

extern "C" int dhd_probe_write_bw(void* dst, size_t bytes, int mode, int chunk_bytes, int blocks_per_sm,
                                  void* stream) {
  DHD_REQUIRE(dst != nullptr && bytes % 256 == 0 && ((uintptr_t)dst & 15) == 0, "dst must be 16-byte aligned, bytes % 256 == 0");
  DHD_REQUIRE(blocks_per_sm >= 1 && blocks_per_sm <= 8, "blocks_per_sm out of range");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = sm_count() * blocks_per_sm;
  if (mode <= 2) {
    probe_write_kernel<<<grid, 256, 0, st>>>((float4*)dst, bytes / 16, mode);
  } else {
    DHD_REQUIRE(chunk_bytes >= 256 && chunk_bytes <= 64 * 1024 && chunk_bytes % 256 == 0, "bad chunk size");
    static bool attr = false;
    if (!attr) {
      cudaFuncSetAttribute(probe_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
      attr = true;
    }
    probe_bulk_kernel<<<grid, 256, chunk_bytes, st>>>((char*)dst, bytes, chunk_bytes, mode == 4);
  }
  DHD_CUDA_LAUNCH_CHECK("probe_write");
  return DHD_OK;
}
