// Plane-sweep stereo cost volume of the camera-aware DepthNet (DHD-M / DHD-L):
//   reference  models/model_utils/depthnet.py:245-308 (gen_grid) + 310-361 (calculate_cost_volumn)
// The reference materialises the (B*N, D*H, W, 2) sampling grid, then runs C/4 grid_sample calls that each write a
// (B*N, 4, D, H, W) warped tensor, an abs/sum pass per group and a softmax: ~100 GB of HBM traffic at DHD-L size for a
// 190 MB result.  Here it is ONE kernel: a warp owns one pixel of the 1/4-resolution map; features are NHWC, so every
// bilinear tap is one contiguous row of C values; the sampling coordinates of 32 depth hypotheses are computed at a
// time (one per lane, separately rounded fp32 ops in the reference's order) and handed round by shuffle, the L1
// distance is reduced inside 8-lane groups, and the softmax over depth happens in registers.  Nothing but the
// probabilities is written.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"

namespace dhd {

namespace {

constexpr int kTileX = 4, kTileY = 2;          // pixels of one CTA (8 warps): neighbours share taps through L1
constexpr int kMaxBatches = 4;                 // 32 depth hypotheses per batch -> D <= 128

struct StereoParams {
  dhd_stereo_desc d;
  int tiles_x, tiles_y;
  int pf_rows;          // L2 prefetch distance in map rows (0 = off; DHD_STEREO_PF overrides the default)
};

// one 16-byte load per lane: 4 fp32 or 8 bf16 channels
template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int N = 4;
  static __device__ __forceinline__ void load(const float* p, float* v) {
    const float4 r = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = r.x; v[1] = r.y; v[2] = r.z; v[3] = r.w;
  }
};
template <> struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float* v) {
    const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {                       // bf16 -> fp32 is a 16-bit shift
      v[2 * i] = __uint_as_float(w[i] << 16);
      v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
};

// 3x3 (row-major) times vector, products and adds rounded separately, left to right: the order torch's batched
// matmul gives the reference (same helper as the voxel pool's geometry, pinned there bit for bit)
__device__ __forceinline__ void matvec_rn(const float* __restrict__ m, float x, float y, float z, float* ox,
                                          float* oy, float* oz) {
  *ox = __fadd_rn(__fadd_rn(__fmul_rn(m[0], x), __fmul_rn(m[1], y)), __fmul_rn(m[2], z));
  *oy = __fadd_rn(__fadd_rn(__fmul_rn(m[3], x), __fmul_rn(m[4], y)), __fmul_rn(m[5], z));
  *oz = __fadd_rn(__fadd_rn(__fmul_rn(m[6], x), __fmul_rn(m[7], y)), __fmul_rn(m[8], z));
}

// gen_grid for one frustum point: normalised sampling coordinate in the previous frame's feature map
__device__ __forceinline__ void sampling_coordinate(const float* __restrict__ cam, float u, float v, float dd,
                                                    float wm1, float hm1, float* gx, float* gy) {
  float x = __fsub_rn(u, cam[9]), y = __fsub_rn(v, cam[10]), z = __fsub_rn(dd, cam[11]);
  float rx, ry, rz;
  matvec_rn(cam, x, y, z, &rx, &ry, &rz);                    // inverse(post_rots)
  x = __fmul_rn(rx, rz);
  y = __fmul_rn(ry, rz);
  z = rz;
  matvec_rn(cam + 12, x, y, z, &rx, &ry, &rz);               // k2s rotation @ inverse(intrins)
  rx = __fadd_rn(rx, cam[21]);
  ry = __fadd_rn(ry, cam[22]);
  rz = __fadd_rn(rz, cam[23]);
  const bool behind = rz < 1e-3f;
  matvec_rn(cam + 24, rx, ry, rz, &x, &y, &z);               // intrins
  const float px = __fdiv_rn(x, z), py = __fdiv_rn(y, z);
  const float ax = __fadd_rn(__fadd_rn(__fmul_rn(cam[33], px), __fmul_rn(cam[34], py)), cam[37]);
  const float ay = __fadd_rn(__fadd_rn(__fmul_rn(cam[35], px), __fmul_rn(cam[36], py)), cam[38]);
  *gx = behind ? -2.f : __fsub_rn(__fmul_rn(__fdiv_rn(ax, wm1), 2.f), 1.f);
  *gy = behind ? -2.f : __fsub_rn(__fmul_rn(__fdiv_rn(ay, hm1), 2.f), 1.f);
}

// Work split inside a warp: the 32 lanes first compute the sampling parameters of 32 depth hypotheses (one each),
// then walk them FOUR at a time -- lane group g = lane / 8 takes hypothesis 4j + g, and its 8 lanes own the channels
// (16-byte slot q = l + 8k of the pixel's row, so one load instruction of a group covers 128 contiguous bytes of a tap).
// The L1 distance then needs a 3-step butterfly per four hypotheses instead of a 5-step one per hypothesis, and the
// parameters travel by one indexed shuffle each.
template <typename T, int NK, bool FULL>
__global__ void __launch_bounds__(256, 3) stereo_cost_volume_kernel(const StereoParams P) {
  constexpr int V = Vec<T>::N;
  const dhd_stereo_desc& c = P.d;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int grp = lane >> 3, gl = lane & 7;
  int t = blockIdx.x;
  const int tx = t % P.tiles_x;
  t /= P.tiles_x;
  const int ty = t % P.tiles_y;
  const int bn = t / P.tiles_y;
  const int x = tx * kTileX + (warp % kTileX), y = ty * kTileY + (warp / kTileX);
  if (x >= c.W || y >= c.H) return;                          // whole warp leaves together
  const int H = c.H, W = c.W, C = c.C, D = c.D;
  // Every feature row is first touched from DRAM exactly once, and a warp that waits ~1 us for it has nothing else to
  // do: tiles therefore ask the L2 for the rows the grid will reach `pf_rows` rows from now (one bulk prefetch per
  // tile row and tensor; the tile at the top of the launch also covers the first pf_rows rows).
  if (P.pf_rows > 0) {
    const int seg_x = tx * kTileX;
    const unsigned bytes = (unsigned)(min(kTileX, W - seg_x) * C * (int)sizeof(T));
    const long rows_total = (long)c.BN * H;
    const long g0 = (long)bn * H + ty * kTileY;
    const int n_rows = (bn == 0 && ty == 0) ? P.pf_rows + kTileY : kTileY;
    for (int i = threadIdx.x; i < 2 * n_rows; i += blockDim.x) {
      const long g = (n_rows == kTileY ? g0 + P.pf_rows : 0) + (i >> 1);
      if (g < rows_total) {
        const T* base = reinterpret_cast<const T*>((i & 1) ? c.curr : c.prev) + ((size_t)g * W + seg_x) * C;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(base), "r"(bytes) : "memory");
      }
    }
  }
  const T* prev = reinterpret_cast<const T*>(c.prev) + (size_t)bn * H * W * C;
  const T* cur_px = reinterpret_cast<const T*>(c.curr) + ((size_t)(bn * H + y) * W + x) * C;

  float cur[NK][V];
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    const int ch = (gl + 8 * k) * V;
    if (ch < C) {
      Vec<T>::load(cur_px + ch, cur[k]);
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) cur[k][i] = 0.f;
    }
  }
  // the reference's "sample fell outside" test looks at the first channel of the LAST group of four channels
  const int flag_k = ((C - 4) / V) / 8, flag_gl = ((C - 4) / V) % 8, flag_i = (C - 4) % V;
  const float* cam = c.cam != nullptr ? c.cam + (size_t)bn * DHD_STEREO_CAM_FLOATS : nullptr;
  const float fu = c.grid == nullptr ? __ldg(c.frustum_u + x) : 0.f, fv = c.grid == nullptr ? __ldg(c.frustum_v + y) : 0.f;
  const float wm1 = c.img_w - 1.f, hm1 = c.img_h - 1.f;
  const float sx = (float)(W - 1), sy = (float)(H - 1);

  float cost[kMaxBatches];
#pragma unroll
  for (int b = 0; b < kMaxBatches; ++b) {
    cost[b] = 0.f;
    const int d0 = b * 32;
    if (d0 >= D) continue;                                   // warp-uniform
    const int dl = d0 + lane;
    // ---- this lane's hypothesis: four tap weights (0 where a tap is outside) and clamped tap coordinates
    float w_nw = 0.f, w_ne = 0.f, w_sw = 0.f, w_se = 0.f;
    int o_nw = 0, o_ne = 0, o_sw = 0, o_se = 0;
    if (dl < D) {
      float gx, gy;
      const size_t pt = ((size_t)dl * H + y) * W + x;
      if (c.grid != nullptr) {
        const float2 g = __ldg(reinterpret_cast<const float2*>(c.grid) + (size_t)bn * D * H * W + pt);
        gx = g.x;
        gy = g.y;
      } else {
        sampling_coordinate(cam, fu, fv, __ldg(c.frustum_d + dl), wm1, hm1, &gx, &gy);
      }
      if (c.grid_out != nullptr)
        reinterpret_cast<float2*>(c.grid_out)[(size_t)bn * D * H * W + pt] = make_float2(gx, gy);
      // grid_sample, align_corners=True: ((g + 1) / 2) * (size - 1)
      const float fx = __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.f), 0.5f), sx);
      const float fy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.f), 0.5f), sy);
      if (fx > -1.f && fx < (float)W && fy > -1.f && fy < (float)H) {     // false for NaN too
        const float x0f = floorf(fx), y0f = floorf(fy);
        const int x0 = (int)x0f, y0 = (int)y0f;
        const float wx1 = fx - x0f, wx0 = (x0f + 1.f) - fx;
        const float wy1 = fy - y0f, wy0 = (y0f + 1.f) - fy;
        const bool xl = x0 >= 0, xr = x0 + 1 < W, yt = y0 >= 0, yb = y0 + 1 < H;
        w_nw = (xl && yt) ? wx0 * wy0 : 0.f;
        w_ne = (xr && yt) ? wx1 * wy0 : 0.f;
        w_sw = (xl && yb) ? wx0 * wy1 : 0.f;
        w_se = (xr && yb) ? wx1 * wy1 : 0.f;
        const int xa = max(x0, 0), xb = min(x0 + 1, W - 1), ya = max(y0, 0), yb2 = min(y0 + 1, H - 1);
        o_nw = (ya * W + xa) * C;                          // element offsets inside one image (< 2^31, checked by the host)
        o_ne = (ya * W + xb) * C;
        o_sw = (yb2 * W + xa) * C;
        o_se = (yb2 * W + xb) * C;
      }
    }
    const int nd = min(32, D - d0);
#pragma unroll 1
    for (int j = 0; j * 4 < nd; ++j) {
      const int src = j * 4 + grp;
      const float a_nw = __shfl_sync(kFull, w_nw, src), a_ne = __shfl_sync(kFull, w_ne, src);
      const float a_sw = __shfl_sync(kFull, w_sw, src), a_se = __shfl_sync(kFull, w_se, src);
      const T* p_nw = prev + __shfl_sync(kFull, o_nw, src) + gl * V;
      const T* p_ne = prev + __shfl_sync(kFull, o_ne, src) + gl * V;
      const T* p_sw = prev + __shfl_sync(kFull, o_sw, src) + gl * V;
      const T* p_se = prev + __shfl_sync(kFull, o_se, src) + gl * V;
      float acc = 0.f;
      bool zero_flag = false;
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        if (FULL || (gl + 8 * k) * V < C) {
          float a[V], bq[V], cq[V], dq[V];
          Vec<T>::load(p_nw + 8 * k * V, a);
          Vec<T>::load(p_ne + 8 * k * V, bq);
          Vec<T>::load(p_sw + 8 * k * V, cq);
          Vec<T>::load(p_se + 8 * k * V, dq);
#pragma unroll
          for (int i = 0; i < V; ++i) {
            // cur - (a w_nw + b w_ne + c w_sw + d w_se) as four fused multiply-adds
            const float t = fmaf(-dq[i], a_se, fmaf(-cq[i], a_sw, fmaf(-bq[i], a_ne, fmaf(-a[i], a_nw, cur[k][i]))));
            acc += fabsf(t);
          }
          if (k == flag_k) {                             // the warped value itself, for the reference's `== 0` test
            const float v0 = a[0] * a_nw + bq[0] * a_ne + cq[0] * a_sw + dq[0] * a_se;
            const float v4 = a[V - 4] * a_nw + bq[V - 4] * a_ne + cq[V - 4] * a_sw + dq[V - 4] * a_se;
            zero_flag = (flag_i == 0 ? v0 : v4) == 0.f;
          }
        }
      }
      acc += __shfl_xor_sync(kFull, acc, 1);
      acc += __shfl_xor_sync(kFull, acc, 2);
      acc += __shfl_xor_sync(kFull, acc, 4);
      if (c.bias != 0.f) {
        if (__shfl_sync(kFull, (int)zero_flag, (lane & 24) | flag_gl) != 0) acc += c.bias;
      }
      // group g holds hypothesis 4j + g: hand it to the lane that owns that hypothesis for the softmax
      const float v = __shfl_sync(kFull, acc, (lane & 3) * 8);
      if ((lane >> 2) == j) cost[b] = v;
    }
  }

  // softmax over depth of the negated cost (depthnet.py:358-360)
  float m = -3.0e38f;
#pragma unroll
  for (int b = 0; b < kMaxBatches; ++b)
    if (b * 32 + lane < D) m = fmaxf(m, -cost[b]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(kFull, m, o));
  float e[kMaxBatches], s = 0.f;
#pragma unroll
  for (int b = 0; b < kMaxBatches; ++b) {
    e[b] = (b * 32 + lane < D) ? expf(-cost[b] - m) : 0.f;
    s += e[b];
  }
  s = warp_sum(s);
  const size_t pix = (size_t)(bn * H + y) * W + x;
#pragma unroll
  for (int b = 0; b < kMaxBatches; ++b) {
    const int dl = b * 32 + lane;
    const float p = e[b] / s;
    if (c.out_f32 != nullptr && dl < D)
      c.out_f32[(size_t)bn * c.f32_sN + (size_t)dl * c.f32_sD + (size_t)y * c.f32_sY + (size_t)x * c.f32_sX] = p;
    if (c.out_b16 != nullptr && dl < c.b16_cpad) {           // split-bf16 NHWC activation, pad channels zeroed
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(c.out_b16) + pix * c.b16_ld + c.b16_coff + dl;
      float r = dl < D ? p : 0.f;
      for (int q = 0; q < c.b16_parts; ++q) {
        const __nv_bfloat16 h = __float2bfloat16_rn(r);
        o[(size_t)q * c.b16_part_stride] = h;
        r -= __bfloat162float(h);
      }
    }
  }
}

template <typename OUT>
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* __restrict__ in, int C, int HW,
                                                           OUT* __restrict__ out) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z, c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const float* src = in + (size_t)n * C * HW;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int ch = c0 + r, p = p0 + tx;
    tile[r][tx] = (ch < C && p < HW) ? __ldg(src + (size_t)ch * HW + p) : 0.f;
  }
  __syncthreads();
  OUT* dst = out + (size_t)n * HW * C;
#pragma unroll
  for (int r = ty; r < 32; r += 8) {
    const int p = p0 + r, ch = c0 + tx;
    if (p < HW && ch < C) {
      if constexpr (sizeof(OUT) == 4) dst[(size_t)p * C + ch] = tile[tx][r];
      else dst[(size_t)p * C + ch] = __float2bfloat16_rn(tile[tx][r]);
    }
  }
}

template <typename T, bool FULL>
int launch_cost_volume(const StereoParams& P, int nk, int blocks, cudaStream_t st) {
  if (nk <= 1) stereo_cost_volume_kernel<T, 1, FULL><<<blocks, 256, 0, st>>>(P);
  else if (nk <= 2) stereo_cost_volume_kernel<T, 2, FULL><<<blocks, 256, 0, st>>>(P);
  else if (nk <= 4) stereo_cost_volume_kernel<T, 4, FULL><<<blocks, 256, 0, st>>>(P);
  else if (nk <= 8) stereo_cost_volume_kernel<T, 8, FULL><<<blocks, 256, 0, st>>>(P);
  else stereo_cost_volume_kernel<T, 16, FULL><<<blocks, 256, 0, st>>>(P);
  return 0;
}

}  // namespace

}  // namespace dhd

using namespace dhd;

extern "C" int dhd_nchw_to_nhwc(const float* in, int N, int C, int HW, void* out, int out_bf16, void* stream) {
  DHD_REQUIRE(in && out && N > 0 && C > 0 && HW > 0 && N <= 65535, "bad arguments");
  const dim3 grid((HW + 31) / 32, (C + 31) / 32, N);
  DHD_REQUIRE(grid.y <= 65535, "too many channels");
  if (out_bf16)
    nchw_to_nhwc_kernel<__nv_bfloat16><<<grid, 256, 0, (cudaStream_t)stream>>>(in, C, HW, (__nv_bfloat16*)out);
  else
    nchw_to_nhwc_kernel<float><<<grid, 256, 0, (cudaStream_t)stream>>>(in, C, HW, (float*)out);
  DHD_CUDA_LAUNCH_CHECK("nchw_to_nhwc");
  return DHD_OK;
}

extern "C" int dhd_stereo_cost_volume(const dhd_stereo_desc* d, void* stream) {
  DHD_REQUIRE(d != nullptr, "stereo desc is null");
  DHD_REQUIRE(d->prev && d->curr, "null feature pointer");
  DHD_REQUIRE(d->BN > 0 && d->H > 0 && d->W > 0 && (long)d->H * d->W * d->C < (1L << 31), "bad map shape");
  DHD_REQUIRE(d->C >= 4 && d->C % 4 == 0 && d->C <= 512, "C must be a multiple of 4, at most 512");
  DHD_REQUIRE(!d->feat_bf16 || d->C % 8 == 0, "bf16 features: C must be a multiple of 8");
  DHD_REQUIRE(d->D >= 1 && d->D <= 32 * kMaxBatches, "D must be in 1..128");
  DHD_REQUIRE(d->grid != nullptr || (d->frustum_u != nullptr && d->frustum_v != nullptr && d->frustum_d != nullptr &&
                                     d->cam != nullptr),
              "either the sampling grid or the frustum vectors + camera matrices must be given");
  DHD_REQUIRE(d->out_f32 != nullptr || d->out_b16 != nullptr, "no output");
  DHD_REQUIRE(((uintptr_t)d->prev & 15) == 0 && ((uintptr_t)d->curr & 15) == 0 && ((uintptr_t)d->grid & 7) == 0 &&
                  ((uintptr_t)d->grid_out & 7) == 0,
              "feature / grid pointers must be 16 / 8-byte aligned");
  if (d->out_b16 != nullptr)
    DHD_REQUIRE(d->b16_parts >= 1 && d->b16_parts <= 3 && d->b16_cpad >= d->D && d->b16_cpad <= 32 * kMaxBatches &&
                    d->b16_coff + d->b16_cpad <= d->b16_ld,
                "bad split-bf16 output description");
  StereoParams P;
  P.d = *d;
  P.tiles_x = (d->W + kTileX - 1) / kTileX;
  P.tiles_y = (d->H + kTileY - 1) / kTileY;
  static const int pf_default = [] {
    const char* e = getenv("DHD_STEREO_PF");
    return e != nullptr ? atoi(e) : 24;
  }();
  P.pf_rows = ((size_t)d->C * kTileX * (d->feat_bf16 ? 2 : 4)) % 16 == 0 ? pf_default : 0;
  const long blocks = (long)P.tiles_x * P.tiles_y * d->BN;
  DHD_REQUIRE(blocks < (1L << 31), "map too large");
  const int nk = d->feat_bf16 ? (d->C + 63) / 64 : (d->C + 31) / 32;      // 16-byte slots per lane of an 8-lane group
  const int slot = d->feat_bf16 ? 64 : 32;              // channels one k-step of an 8-lane group covers
  const int nk_pow2 = nk <= 1 ? 1 : nk <= 2 ? 2 : nk <= 4 ? 4 : nk <= 8 ? 8 : 16;
  const bool full = d->C == nk_pow2 * slot;             // every slot in range: the kernel drops its channel guards
  const cudaStream_t st = (cudaStream_t)stream;
  if (d->feat_bf16) {
    if (full) launch_cost_volume<__nv_bfloat16, true>(P, nk, (int)blocks, st);
    else launch_cost_volume<__nv_bfloat16, false>(P, nk, (int)blocks, st);
  } else {
    if (full) launch_cost_volume<float, true>(P, nk, (int)blocks, st);
    else launch_cost_volume<float, false>(P, nk, (int)blocks, st);
  }
  DHD_CUDA_LAUNCH_CHECK("stereo_cost_volume");
  return DHD_OK;
}
