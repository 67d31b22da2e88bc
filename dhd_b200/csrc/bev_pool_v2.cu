// Drop-in bev_pool_v2 operator (same tensor contracts as the reference extension,
// reference: projects/mmdet3d_plugin/ops/bev_pool_v2/src/bev_pool_cuda.cu:21-50, 69-123).
//
// Layout of work differs from the reference: one WARP owns one interval, lanes own
// channels (c, c+32, ...), so the three index arrays are read once per point per warp
// (broadcast) instead of once per (point, channel), feat rows are read as full 128-byte
// lines, and the backward gets channel parallelism (the reference walks 64 channels
// serially in one thread).  Grid = multiple of the SM count, warps stride over intervals.
#include "common.cuh"

namespace dhd {

thread_local char g_err[512] = "";
std::atomic<long> g_launches{0};

template <int CPL>  // channels per lane: C <= 32*CPL
__global__ void __launch_bounds__(256)
bev_pool_v2_fwd_kernel(int c, int n_intervals, const float* __restrict__ depth,
                       const float* __restrict__ feat, const int* __restrict__ ranks_depth,
                       const int* __restrict__ ranks_feat, const int* __restrict__ ranks_bev,
                       const int* __restrict__ starts, const int* __restrict__ lengths,
                       float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int iv = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; iv < n_intervals; iv += warps) {
    const int s = starts[iv], n = lengths[iv];
    float acc[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) acc[k] = 0.f;
    for (int base = 0; base < n; base += 32) {
      // lanes fetch 32 points' indices and depth in parallel, then broadcast one by one
      const int m = min(32, n - base);
      int rf = 0;
      float dv = 0.f;
      if (lane < m) {
        rf = ranks_feat[s + base + lane];
        dv = depth[ranks_depth[s + base + lane]];
      }
      for (int j = 0; j < m; ++j) {
        const int f = __shfl_sync(kFull, rf, j);
        const float d = __shfl_sync(kFull, dv, j);
        const float* row = feat + (size_t)f * c;
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
          const int ch = lane + 32 * k;
          if (ch < c) acc[k] = fmaf(row[ch], d, acc[k]);
        }
      }
    }
    float* o = out + (size_t)ranks_bev[s] * c;
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      const int ch = lane + 32 * k;
      if (ch < c) o[ch] = acc[k];
    }
  }
}

template <int CPL>
__global__ void __launch_bounds__(256)
bev_pool_v2_bwd_kernel(int c, int n_intervals, const float* __restrict__ out_grad,
                       const float* __restrict__ depth, const float* __restrict__ feat,
                       const int* __restrict__ ranks_depth, const int* __restrict__ ranks_feat,
                       const int* __restrict__ ranks_bev, const int* __restrict__ starts,
                       const int* __restrict__ lengths, float* __restrict__ depth_grad,
                       float* __restrict__ feat_grad) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  for (int iv = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; iv < n_intervals; iv += warps) {
    const int s = starts[iv], n = lengths[iv];
    float facc[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) facc[k] = 0.f;
    for (int base = 0; base < n; base += 32) {
      const int m = min(32, n - base);
      int rf = 0, rb = 0, rd = 0;
      float dv = 0.f;
      if (lane < m) {
        rf = ranks_feat[s + base + lane];
        rb = ranks_bev[s + base + lane];
        rd = ranks_depth[s + base + lane];
        dv = depth[rd];
      }
      for (int j = 0; j < m; ++j) {
        const int f = __shfl_sync(kFull, rf, j);
        const int b = __shfl_sync(kFull, rb, j);
        const int dd = __shfl_sync(kFull, rd, j);
        const float d = __shfl_sync(kFull, dv, j);
        const float* g = out_grad + (size_t)b * c;
        const float* fr = feat + (size_t)f * c;
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
          const int ch = lane + 32 * k;
          if (ch < c) {
            const float gv = g[ch];
            dot = fmaf(gv, fr[ch], dot);
            facc[k] = fmaf(gv, d, facc[k]);
          }
        }
        dot = warp_sum(dot);
        if (lane == 0) depth_grad[dd] = dot;
      }
    }
    float* fg = feat_grad + (size_t)ranks_feat[s] * c;
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      const int ch = lane + 32 * k;
      if (ch < c) fg[ch] = facc[k];
    }
  }
}

static int grid_for_warps(long warps_needed) {
  const long per_block = 8;  // 256 threads
  long blocks = (warps_needed + per_block - 1) / per_block;
  const long cap = (long)sm_count() * 8;  // 8 resident 256-thread blocks per SM
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace dhd

using namespace dhd;

extern "C" const char* dhd_last_error(void) { return dhd::g_err; }
extern "C" int dhd_abi_version(void) { return DHD_ABI_VERSION; }
// sizeof of the ABI structs, so a binding can verify its own layout (tests/test_abi.py does for the ctypes mirror)
extern "C" size_t dhd_abi_sizeof(int which) {
  switch (which) {
    case 0: return sizeof(dhd_mghs_cfg);
    case 1: return sizeof(dhd_conv_seg);
    case 2: return sizeof(dhd_conv_desc);
    case 3: return sizeof(dhd_wgrad_desc);
    case 4: return sizeof(dhd_stereo_desc);
    case 5: return sizeof(dhd_predictor_tail_desc);
    case 6: return sizeof(dhd_pack_desc);
    default: return 0;
  }
}

extern "C" int dhd_bev_pool_v2_fwd(int c, int n_intervals, const float* depth, const float* feat,
                                   const int32_t* ranks_depth, const int32_t* ranks_feat,
                                   const int32_t* ranks_bev, const int32_t* interval_starts,
                                   const int32_t* interval_lengths, float* out, void* stream) {
  DHD_REQUIRE(c > 0 && c <= 256, "bev_pool_v2: channel count must be in 1..256");
  DHD_REQUIRE(n_intervals >= 0, "bev_pool_v2: negative interval count");
  if (n_intervals == 0) return DHD_OK;
  DHD_REQUIRE(depth && feat && ranks_depth && ranks_feat && ranks_bev && interval_starts &&
                  interval_lengths && out, "bev_pool_v2: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for_warps(n_intervals);
#define LAUNCH(CPL)                                                                              \
  bev_pool_v2_fwd_kernel<CPL><<<grid, 256, 0, st>>>(c, n_intervals, depth, feat, ranks_depth,    \
                                                    ranks_feat, ranks_bev, interval_starts,      \
                                                    interval_lengths, out)
  if (c <= 32) LAUNCH(1);
  else if (c <= 64) LAUNCH(2);
  else if (c <= 128) LAUNCH(4);
  else LAUNCH(8);
#undef LAUNCH
  DHD_CUDA_LAUNCH_CHECK("bev_pool_v2_fwd");
  return DHD_OK;
}

extern "C" int dhd_bev_pool_v2_bwd(int c, int n_intervals, const float* out_grad,
                                   const float* depth, const float* feat,
                                   const int32_t* ranks_depth, const int32_t* ranks_feat,
                                   const int32_t* ranks_bev, const int32_t* interval_starts,
                                   const int32_t* interval_lengths, float* depth_grad,
                                   float* feat_grad, void* stream) {
  DHD_REQUIRE(c > 0 && c <= 256, "bev_pool_v2: channel count must be in 1..256");
  DHD_REQUIRE(n_intervals >= 0, "bev_pool_v2: negative interval count");
  if (n_intervals == 0) return DHD_OK;
  DHD_REQUIRE(out_grad && depth && feat && ranks_depth && ranks_feat && ranks_bev &&
                  interval_starts && interval_lengths && depth_grad && feat_grad,
              "bev_pool_v2: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for_warps(n_intervals);
#define LAUNCH(CPL)                                                                              \
  bev_pool_v2_bwd_kernel<CPL><<<grid, 256, 0, st>>>(c, n_intervals, out_grad, depth, feat,       \
                                                    ranks_depth, ranks_feat, ranks_bev,          \
                                                    interval_starts, interval_lengths,           \
                                                    depth_grad, feat_grad)
  if (c <= 32) LAUNCH(1);
  else if (c <= 64) LAUNCH(2);
  else if (c <= 128) LAUNCH(4);
  else LAUNCH(8);
#undef LAUNCH
  DHD_CUDA_LAUNCH_CHECK("bev_pool_v2_bwd");
  return DHD_OK;
}
