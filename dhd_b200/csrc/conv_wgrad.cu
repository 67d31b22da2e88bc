// Weight gradient of the implicit-GEMM convolution on tcgen05 tensor cores (sm_100a).
//
//   dw[co][tap][ci] = sum over pixels p of  dy[p][co] * x[p + tap][ci]
//
// (torch: the grad_weight of F.conv2d; reference layers: every nn.Conv2d / nn.Linear of the hot
// path listed at dhd_conv2d_fwd).  As a GEMM this is M = co, N = ci, K = pixels, and both
// operands arrive PIXEL-major from the NHWC activations -- i.e. "MN-major" in UMMA terms: the
// 64 channels of one pixel are the contiguous 128 bytes.  A 4-D TMA box {64 ch, bw, bh, 1} with
// SWIZZLE_128B therefore lands in shared memory exactly as the canonical MN-major SW128 layout
// ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units: 8 pixel rows form a 1024-byte atom (SBO), the
// next 64-channel block is the next box (LBO = one box = 16 KB).  The tap shift of x and the image
// borders are the TMA box coordinates / OOB zero fill, as in the forward kernel.
//
// One CTA owns (128 output channels) x (<=256 input channels) x (one tap) x (a slice of the pixel
// tiles, streamed 64 pixels per stage through a 4-stage TMA ring); the fp32 accumulator lives in TMEM for the whole slice, is written once to a partial
// buffer, and a second small kernel sums the slices in a fixed order (deterministic, no atomics).
#include "common.cuh"
#include "tc_ptx.cuh"

namespace dhd {

constexpr int kWgPix = 64;                        // pixels (= GEMM K) per pipeline stage: half of a 128-pixel tile box
constexpr int kWgBoxBytes = kWgPix * 128;         // one {64 ch x 64 px} bf16 box
constexpr int kWgStages = 4;
constexpr int kWgThreads = 128;

struct WgradParams {
  dhd_wgrad_desc d;
  int sbw, sbh;                     // stage box = half of the (bw x bh) tile box: sbw * sbh == 64 pixels
  int tiles_w, tiles_h, tiles;      // stage boxes per image row / column / total
  int co_blocks, ci_blocks, ncols;  // ncols = input channels per CTA (<= 256, multiple of 64)
  int splits;
};
struct WgradMaps {
  CUtensorMap x, dy;
};

// MN-major, 128B-swizzled operand: LBO = bytes between 64-channel blocks, SBO = bytes between 8-row K atoms
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t umma_instr_desc_bf16_mn(int M, int N) {
  return umma_instr_desc_bf16(M, N) | (1u << 15) | (1u << 16);     // a_major = b_major = MN
}

__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_kernel(const __grid_constant__ WgradMaps M, const __grid_constant__ WgradParams P) {
  extern __shared__ uint8_t smem_raw[];
  const dhd_wgrad_desc& d = P.d;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const int nbx = P.ncols / 64;                                   // x boxes per stage
  const uint32_t stage_bytes = (uint32_t)(2 + nbx) * kWgBoxBytes;
  const uint32_t bar_base = base + kWgStages * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kWgStages + s); };
  const uint32_t done_bar = bar_base + 8u * (2 * kWgStages);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + (bar_base - raw) + 8u * (2 * kWgStages + 1));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // work item
  int w = blockIdx.x;
  const int split = w % P.splits;  w /= P.splits;
  const int cib = w % P.ci_blocks; w /= P.ci_blocks;
  const int tap = w % d.taps;      w /= d.taps;
  const int cob = w;
  const int t_lo = (int)((long)P.tiles * split / P.splits), t_hi = (int)((long)P.tiles * (split + 1) / P.splits);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&M.x) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&M.dy) : "memory");
    for (int s = 0; s < kWgStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(256)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int t, int& img, int& x0, int& y0) {
    const int tx = t % P.tiles_w;
    t /= P.tiles_w;
    const int ty = t % P.tiles_h;
    img = t / P.tiles_h;
    x0 = tx * P.sbw;
    y0 = ty * P.sbh;
  };

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      const int xs = d.x_stride > 1 ? d.x_stride : 1;
      for (int t = t_lo; t < t_hi; ++t, ++it) {
        int img, x0, y0;
        decode(t, img, x0, y0);
        const int s = it % kWgStages;
        const uint32_t ph = (it / kWgStages) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        const uint32_t sa = base + s * stage_bytes;
        mbar_expect_tx(full_bar(s), stage_bytes);
        for (int i = 0; i < 2; ++i)
          tma_load_4d(sa + i * kWgBoxBytes, &M.dy, full_bar(s), d.dy_coff + cob * 128 + i * 64, x0, y0, img);
        for (int i = 0; i < nbx; ++i)
          tma_load_4d(sa + (2 + i) * kWgBoxBytes, &M.x, full_bar(s), d.x_coff + cib * P.ncols + i * 64,
                      x0 * xs + d.tap_dx[tap], y0 * xs + d.tap_dy[tap], img);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_instr_desc_bf16_mn(128, P.ncols);
      int it = 0;
      for (int t = t_lo; t < t_hi; ++t, ++it) {
        const int s = it % kWgStages;
        const uint32_t ph = (it / kWgStages) & 1;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t sa = base + s * stage_bytes, sb = sa + 2 * kWgBoxBytes;
#pragma unroll
        for (int k = 0; k < kWgPix / kUmmaK; ++k) {
          const uint64_t da = umma_desc_mn_sw128(sa + k * 2048, kWgBoxBytes, 1024);
          const uint64_t db = umma_desc_mn_sw128(sb + k * 2048, kWgBoxBytes, 1024);
          umma_bf16(tmem_base, da, db, idesc, (it == 0 && k == 0) ? 0u : 1u);
        }
        umma_commit(empty_bar(s));
      }
      umma_commit(done_bar);
    }
    __syncwarp();
  }
  // ---- epilogue: TMEM lane = output channel, column = input channel -> partial[split][co][tap][ci]
  if (t_hi > t_lo) {
    mbar_wait(done_bar, 0);
    tc_fence_after();
  }
  const int co = cob * 128 + warp * 32 + lane;
  float* dst = d.partial + (((size_t)split * d.Cout + co) * d.taps + tap) * d.Cin + (size_t)cib * P.ncols;
  for (int cb = 0; cb < P.ncols / 32; ++cb) {
    float v[32];
    if (t_hi > t_lo) {
      tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(cb * 32), v);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = 0.f;
    }
    if (co < d.Cout) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        reinterpret_cast<float4*>(dst + cb * 32)[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(256) : "memory");
  }
}

// dw[i] (+)= scale[co] * sum_s partial[s][i], slices added in ascending order
__global__ void __launch_bounds__(256)
wgrad_reduce_kernel(const float* __restrict__ partial, int splits, long n, int per_co, const float* __restrict__ scale,
                    float* __restrict__ dw, int accumulate) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float a = 0.f;
  for (int s = 0; s < splits; ++s) a += partial[(size_t)s * n + i];
  if (scale != nullptr) a *= scale[i / per_co];
  dw[i] = accumulate ? dw[i] + a : a;
}

// dw_torch layout: dw is the parameter's gradient [Cout][cin_total][taps].  A block owns 32 (co, ci) pairs; thread =
// (slice group g, tap t, lane = ci): reads are 128-byte coalesced, every thread has four independent loads in flight
// and almost no state (the first version kept all taps of a pair in one thread: 64 registers, 38 % occupancy, and ncu
// showed the pass waiting on DRAM latency -- the partials are NOT L2 hits -- at 17 % of the HBM bandwidth).  Group g adds
// slices g, g+G, .. in order; the G group sums meet in shared memory, are added in order and leave in OUTPUT order with
// consecutive threads on consecutive addresses.  32-bit index arithmetic throughout.
__global__ void __launch_bounds__(320)
wgrad_reduce_torch_kernel(const float* __restrict__ partial, int splits, long n, const float* __restrict__ scale,
                          float* __restrict__ dw, int accumulate, int Cout, int Cin, int taps, int cin_total, int cin_lo,
                          int cin_used, int G) {
  extern __shared__ float grp[];                 // [G][32 * taps]
  const unsigned lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;      // warp = (g, t)
  const unsigned g = wrp / (unsigned)taps, t = wrp - g * (unsigned)taps;
  const unsigned p0 = blockIdx.x * 32u;
  const unsigned npairs = (unsigned)Cout * (unsigned)cin_used;
  const unsigned p = p0 + lane;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (p < npairs) {
    const unsigned co = p / (unsigned)cin_used, ci = p - co * (unsigned)cin_used;
    const float* src = partial + ((size_t)co * taps + t) * Cin + ci;
    int sl = (int)g;
    for (; sl + 3 * G < splits; sl += 4 * G) {
      a0 += __ldg(src + (size_t)sl * n);
      a1 += __ldg(src + (size_t)(sl + G) * n);
      a2 += __ldg(src + (size_t)(sl + 2 * G) * n);
      a3 += __ldg(src + (size_t)(sl + 3 * G) * n);
    }
    for (; sl < splits; sl += G) a0 += __ldg(src + (size_t)sl * n);
  }
  grp[g * 32u * taps + lane * taps + t] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  const unsigned cnt = min(32u, npairs - p0) * (unsigned)taps;
  for (unsigned k = threadIdx.x; k < cnt; k += blockDim.x) {
    float a = grp[k];
    for (int gg = 1; gg < G; ++gg) a += grp[gg * 32u * taps + k];
    const unsigned lp = k / (unsigned)taps, tt = k - lp * (unsigned)taps;
    const unsigned q = p0 + lp;
    const unsigned co = q / (unsigned)cin_used, ci = q - co * (unsigned)cin_used;
    if (scale != nullptr) a *= scale[co];
    const unsigned o = (co * (unsigned)cin_total + (unsigned)cin_lo + ci) * (unsigned)taps + tt;
    dw[o] = accumulate ? dw[o] + a : a;
  }
}

typedef CUresult (*EncodeTiledFnW)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);
void* conv_encode_fn();   // conv_igemm.cu

static void wgrad_plan(const dhd_wgrad_desc* d, WgradParams* P) {
  P->d = *d;
  P->sbw = d->bh >= 2 ? d->bw : d->bw / 2;
  P->sbh = d->bh >= 2 ? d->bh / 2 : 1;
  P->tiles_w = (d->W + P->sbw - 1) / P->sbw;
  P->tiles_h = (d->H + P->sbh - 1) / P->sbh;
  P->tiles = P->tiles_w * P->tiles_h * d->N;
  P->co_blocks = (d->Cout + 127) / 128;
  P->ncols = d->Cin % 256 == 0 ? 256 : (d->Cin % 128 == 0 ? 128 : 64);
  if (P->ncols > d->Cin) P->ncols = d->Cin;
  P->ci_blocks = d->Cin / P->ncols;
  const int items = P->co_blocks * d->taps * P->ci_blocks;
  // one CTA per SM (192 KB of shared memory each): as many slices as fill ONE wave, never a partial second one
  int splits = sm_count() / items;
  if (splits > P->tiles) splits = P->tiles;
  if (splits > 128) splits = 128;
  if (splits < 1) splits = 1;
  P->splits = splits;
}

}  // namespace dhd

using namespace dhd;

extern "C" size_t dhd_conv2d_wgrad_workspace_bytes(const dhd_wgrad_desc* d) {
  if (d == nullptr || d->bw <= 0 || d->bh <= 0 || d->Cin <= 0 || d->Cin % 64 != 0) return 0;
  WgradParams P;
  wgrad_plan(d, &P);
  return (size_t)P.splits * d->Cout * d->taps * d->Cin * sizeof(float);
}

extern "C" int dhd_conv2d_wgrad(const dhd_wgrad_desc* d, void* stream) {
  DHD_REQUIRE(d != nullptr, "wgrad desc is null");
  DHD_REQUIRE(d->x && d->dy && d->dw && d->partial, "null pointer");
  DHD_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0, "bad image shape");
  DHD_REQUIRE(d->Cin > 0 && d->Cin % 64 == 0, "Cin must be a multiple of 64");
  DHD_REQUIRE(d->Cout > 0, "bad Cout");
  DHD_REQUIRE(d->taps >= 1 && d->taps <= DHD_CONV_MAX_TAPS, "taps out of range");
  DHD_REQUIRE(d->bw > 0 && d->bh > 0 && d->bw * d->bh == 128 && d->bw <= 256 && d->bh <= 256,
              "tile box must cover exactly 128 pixels");
  DHD_REQUIRE(d->x_ld % 8 == 0 && d->x_coff % 8 == 0 && d->dy_ld % 8 == 0 && d->dy_coff % 8 == 0,
              "channel offsets must be multiples of 8 (16-byte TMA alignment)");
  DHD_REQUIRE(((uintptr_t)d->x & 15) == 0 && ((uintptr_t)d->dy & 15) == 0 && ((uintptr_t)d->dw & 3) == 0 &&
                  ((uintptr_t)d->partial & 15) == 0, "x / dy / partial must be 16-byte aligned");
  if (d->dw_torch != 0)
    DHD_REQUIRE(d->dw_cin_used > 0 && d->dw_cin_used <= d->Cin && d->dw_cin_lo >= 0 &&
                    d->dw_cin_lo + d->dw_cin_used <= d->dw_cin_total &&
                    (long)d->Cout * d->dw_cin_total * d->taps < (1L << 31), "dw_torch: bad input-channel window");
  DHD_REQUIRE(d->dy_coff + d->Cout <= d->dy_ld && d->x_coff + d->Cin <= d->x_ld, "channel range exceeds the row");
  DHD_REQUIRE(d->x_stride == 0 || d->x_stride == 1 || (d->x_stride == 2 && d->x_H >= 2 * d->H - 1 && d->x_W >= 2 * d->W - 1),
              "x_stride must be 1 or 2 (with the x_H x x_W grid covering the strided taps)");
  EncodeTiledFnW enc = (EncodeTiledFnW)conv_encode_fn();
  if (enc == nullptr) return fail(DHD_EUNSUPPORTED, "%s", "cuTensorMapEncodeTiled is unavailable");
  WgradParams P;
  wgrad_plan(d, &P);
  WgradMaps maps;
  cuuint32_t box[4] = {64, (cuuint32_t)P.sbw, (cuuint32_t)P.sbh, 1};
  cuuint32_t es[4] = {1, 1, 1, 1};
  {
    // the channel extent stops at the layer's last channel: a 64-channel box past it reads zeros
    const int xs = d->x_stride > 1 ? d->x_stride : 1;
    const cuuint64_t xw = xs > 1 ? d->x_W : d->W, xh = xs > 1 ? d->x_H : d->H;
    cuuint64_t dims[4] = {(cuuint64_t)(d->x_coff + d->Cin), xw, xh, (cuuint64_t)d->N};
    cuuint64_t strides[3] = {(cuuint64_t)d->x_ld * 2, xw * d->x_ld * 2, xh * xw * d->x_ld * 2};
    cuuint32_t xbox[4] = {64, (cuuint32_t)(P.sbw * xs), (cuuint32_t)(P.sbh * xs), 1};
    cuuint32_t xes[4] = {1, (cuuint32_t)xs, (cuuint32_t)xs, 1};
    CUresult r = enc(&maps.x, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)d->x, dims, strides, xbox, xes,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(DHD_EINVAL, "%s: %ld", "cuTensorMapEncodeTiled(x) failed", (long)r);
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)(d->dy_coff + d->Cout), (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
    cuuint64_t strides[3] = {(cuuint64_t)d->dy_ld * 2, (cuuint64_t)d->W * d->dy_ld * 2,
                             (cuuint64_t)d->H * d->W * d->dy_ld * 2};
    CUresult r = enc(&maps.dy, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)d->dy, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(DHD_EINVAL, "%s: %ld", "cuTensorMapEncodeTiled(dy) failed", (long)r);
  }
  const size_t smem = (size_t)kWgStages * (2 + P.ncols / 64) * kWgBoxBytes + 256 + 1024;
  static size_t smem_set = 0;
  if (smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail((int)e, "%s: %ld", "cudaFuncSetAttribute(conv_wgrad)", (long)e);
    smem_set = smem;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = P.co_blocks * d->taps * P.ci_blocks * P.splits;
  conv_wgrad_kernel<<<grid, kWgThreads, smem, st>>>(maps, P);
  DHD_CUDA_LAUNCH_CHECK("conv_wgrad");
  const long n = (long)d->Cout * d->taps * d->Cin;
  if (d->dw_torch != 0) {
    const long npairs = (long)d->Cout * d->dw_cin_used;
    int G = 8 / d->taps;                          // warps per block = G * taps <= 10
    if (G < 1) G = 1;
    if (G > P.splits) G = P.splits;
    const int threads = 32 * d->taps * G;
    wgrad_reduce_torch_kernel<<<(int)((npairs + 31) / 32), threads, (size_t)G * 32 * d->taps * sizeof(float), st>>>(
        d->partial, P.splits, n, d->scale, d->dw, d->accumulate, d->Cout, d->Cin, d->taps, d->dw_cin_total, d->dw_cin_lo,
        d->dw_cin_used, G);
  } else {
    wgrad_reduce_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(d->partial, P.splits, n, d->taps * d->Cin, d->scale,
                                                                d->dw, d->accumulate);
  }
  DHD_CUDA_LAUNCH_CHECK("wgrad_reduce");
  return DHD_OK;
}
