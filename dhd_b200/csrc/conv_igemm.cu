// Implicit-GEMM 2D convolution / linear layer on the sm_100a tensor cores (tcgen05 + TMEM + TMA).
//
// One kernel serves every dense layer of the DHD hot path (reference call sites:
//   MGHS.depth_net 1x1            models/necks/lss_heightmap.py:62, 482-485
//   HeightNet / DepthNet convs    models/model_utils/depthnet.py:172-243, 418-487 (3x3, dilated 3x3, 1x1)
//   SFA 1x1 / 3x3 convs           models/necks/mix.py:28-33, 74-85
//   predictor conv + MLP          models/dense_heads/occ_head.py:52-67
// which the reference runs as unfused cuDNN / cuBLAS launches).
//
// Formulation.  Activations are NHWC bf16 in HBM; a 128-pixel output tile is a (bh x bw) box
// of one image.  For every filter tap the A operand of the GEMM is that same box shifted by
// the tap offset, fetched by ONE 4-D TMA box load whose out-of-bounds rows / columns are
// zero-filled by the TMA unit -- padding, dilation and image borders cost no instructions and
// no im2col buffer exists.  Weights are [Cout][tap][part][Cin] bf16 (K-major) and arrive by
// 2-D TMA.  Both operands land in 128B-swizzled shared memory, tcgen05.mma (M=128, N=128,
// K=16, bf16 -> fp32) accumulates in TMEM, a 4-warp epilogue pulls the accumulator with
// tcgen05.ld and applies folded BatchNorm / bias / per-image bias / residual / activation /
// SE gate / channel softmax and writes bf16 and/or fp32 with arbitrary strides (NHWC, NCHW,
// transposed BEV).
//
// Precision.  `n_terms` > 1 selects split-bf16 arithmetic: an fp32 tensor is carried as up to
// three bf16 "parts" (x = x0 + x1 + x2, stacked on the channel axis) and the K loop issues
// one MMA per (activation part, weight part) pair listed in term_a/term_b.  1 term = plain
// bf16; 3 terms ~ 2^-16; 6 terms ~ fp32.  The tensor core never sees anything but bf16.
//
// Warp roles (192 threads): warps 0-3 epilogue (TMEM lanes 32w..32w+31), warp 4 TMA producer,
// warp 5 TMEM allocator + single-thread MMA issuer.  3-stage smem ring (A 16 KB + B 16 KB per
// stage) -> 2 CTAs per SM, so one CTA's epilogue overlaps the other's main loop.
#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace dhd {

constexpr int kBlockM = 128;
constexpr int kBlockN = 128;
constexpr int kBlockK = 64;  // bf16 elements = one 128-byte swizzle row
constexpr int kStages = 3;
constexpr int kConvThreads = 192;
constexpr uint32_t kABytes = kBlockM * kBlockK * 2;
constexpr uint32_t kBBytes = kBlockN * kBlockK * 2;
constexpr uint32_t kStageBytes = kABytes + kBBytes;
constexpr uint32_t kConvSmem = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int kTmemCols = 128;

struct ConvKernelParams {
  dhd_conv_desc d;
  int tiles_w, tiles_h;
};

constexpr uint32_t kInstrDesc = umma_instr_desc_bf16(kBlockM, kBlockN);

__device__ __forceinline__ float act_apply(float v, int act) {
  switch (act) {
    case DHD_ACT_RELU: return fmaxf(v, 0.f);
    case DHD_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    case DHD_ACT_SOFTPLUS: return v > 20.f ? v : log1pf(expf(v));  // torch Softplus(beta=1, threshold=20)
    default: return v;
  }
}

// is a filter tap entirely outside the image for this tile? (then it contributes only zeros)
__device__ __forceinline__ bool tap_dead(const dhd_conv_desc& d, int t, int x0, int y0) {
  const int xs = x0 + d.tap_dx[t], ys = y0 + d.tap_dy[t];
  return xs >= d.W || xs + d.bw <= 0 || ys >= d.H || ys + d.bh <= 0;
}

__global__ void __launch_bounds__(kConvThreads, 2)
conv_igemm_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  const __grid_constant__ ConvKernelParams P) {
  extern __shared__ uint8_t smem_raw[];
  const dhd_conv_desc& d = P.d;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = base + kStages * kStageBytes;
  // barriers: full[kStages], empty[kStages], tmem_full; then the TMEM base address word
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * kStages);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + (bar_base - smem_u32(smem_raw)) +
                                                    8u * (2 * kStages + 1));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int mt = blockIdx.x;
  const int tx = mt % P.tiles_w;
  mt /= P.tiles_w;
  const int ty = mt % P.tiles_h;
  const int img = mt / P.tiles_h;
  const int x0 = tx * d.bw, y0 = ty * d.bh;
  const int n0 = blockIdx.y * kBlockN;
  const int kchunks = d.Cin / kBlockK;

  if (warp == 4 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(tmem_slot)),
                 "n"(kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 4) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int it = 0;
      for (int t = 0; t < d.taps; ++t) {
        if (tap_dead(d, t, x0, y0)) continue;
        for (int kc = 0; kc < kchunks; ++kc) {
          for (int e = 0; e < d.n_terms; ++e, ++it) {
            const int s = it % kStages;
            const uint32_t ph = (it / kStages) & 1;
            mbar_wait(empty_bar(s), ph ^ 1);
            const uint32_t sa = base + s * kStageBytes, sb = sa + kABytes;
            mbar_expect_tx(full_bar(s), kStageBytes);
            tma_load_4d(sa, &map_a, full_bar(s),
                        d.in_coff + d.term_a[e] * d.in_part_stride + kc * kBlockK,
                        x0 + d.tap_dx[t], y0 + d.tap_dy[t], img);
            tma_load_2d(sb, &map_b, full_bar(s),
                        (t * d.w_parts + d.term_b[e]) * d.Cin + kc * kBlockK, n0);
          }
        }
      }
    }
  } else if (warp == 5) {
    // ===================================================== MMA issuer
    if (lane == 0) {
      int it = 0;
      for (int t = 0; t < d.taps; ++t) {
        if (tap_dead(d, t, x0, y0)) continue;
        for (int kc = 0; kc < kchunks; ++kc) {
          for (int e = 0; e < d.n_terms; ++e, ++it) {
            const int s = it % kStages;
            const uint32_t ph = (it / kStages) & 1;
            mbar_wait(full_bar(s), ph);
            tc_fence_after();
            const uint32_t sa = base + s * kStageBytes, sb = sa + kABytes;
            const uint64_t da = umma_desc_sw128(sa), db = umma_desc_sw128(sb);
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
              // +32 bytes along K inside the 128B swizzle row = +2 in the (addr >> 4) field
              umma_bf16(tmem_base, da + 2u * k, db + 2u * k, kInstrDesc, (it | k) != 0 ? 1u : 0u);
            }
            umma_commit(empty_bar(s));
          }
        }
      }
      if (it == 0) {
        // every tap was dead (cannot happen for a tile that overlaps the image, kept for safety)
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    // ===================================================== epilogue (warps 0..3)
    // does this tile have at least one live tap?  (the centre tap always is)
    bool any_live = false;
    for (int t = 0; t < d.taps; ++t) any_live |= !tap_dead(d, t, x0, y0);
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const int row = warp * 32 + lane;
    const int px = x0 + row % d.bw, py = y0 + row / d.bw;
    const bool valid = px < d.W && py < d.H;
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);

    for (int sgi = 0; sgi < d.n_seg; ++sgi) {
      const dhd_conv_seg& sg = d.seg[sgi];
      const int c_lo = max(sg.c_lo, n0), c_hi = min(sg.c_hi, min(d.Cout, n0 + kBlockN));
      if (c_lo >= c_hi) continue;
      float mx = 0.f, inv = 1.f;
      auto value = [&](float acc, int c) -> float {
        float v = any_live ? acc : 0.f;
        if (d.scale != nullptr) v *= __ldg(d.scale + c);
        if (d.bias != nullptr) v += __ldg(d.bias + c);
        if (d.img_bias != nullptr) v += __ldg(d.img_bias + (size_t)img * d.Cout + c);
        if (d.residual != nullptr && valid)
          v += __ldg(d.residual + (size_t)img * d.res_sN + (size_t)py * d.res_sY +
                     (size_t)px * d.res_sX + c);
        return v;
      };
      if (sg.act == DHD_ACT_SOFTMAX) {
        // whole softmax range lives in this thread's TMEM row: max pass, sum pass, write pass
        mx = -INFINITY;
        for (int cb = (c_lo - n0) / 32; cb * 32 < c_hi - n0; ++cb) {
          float v[32];
          tmem_ld32(taddr + cb * 32, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int c = n0 + cb * 32 + j;
            if (c >= c_lo && c < c_hi) mx = fmaxf(mx, value(v[j], c));
          }
        }
        float sum = 0.f;
        for (int cb = (c_lo - n0) / 32; cb * 32 < c_hi - n0; ++cb) {
          float v[32];
          tmem_ld32(taddr + cb * 32, v);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int c = n0 + cb * 32 + j;
            if (c >= c_lo && c < c_hi) sum += expf(value(v[j], c) - mx);
          }
        }
        inv = 1.f / sum;
      }
      for (int cb = (c_lo - n0) / 32; cb * 32 < c_hi - n0; ++cb) {
        float v[32];
        tmem_ld32(taddr + cb * 32, v);   // warp-collective: every lane executes it
        if (!valid) continue;
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int c = n0 + cb * 32 + j;
          float o = 0.f;
          if (c >= c_lo && c < c_hi) {
            o = value(v[j], c);
            if (sg.act == DHD_ACT_SOFTMAX) o = expf(o - mx) * inv;
            else o = act_apply(o, sg.act);
            if (d.img_gate != nullptr) o *= __ldg(d.img_gate + (size_t)img * d.Cout + c);
          }
          v[j] = o;
        }
        if (sg.out_f32 != nullptr) {
          float* o = sg.out_f32 + (size_t)img * sg.f32_sN + (size_t)py * sg.f32_sY +
                     (size_t)px * sg.f32_sX;
          const int cfirst = n0 + cb * 32;
          if (sg.f32_sC == 1 && cfirst >= c_lo && cfirst + 32 <= c_hi &&
              (((uintptr_t)(o + (cfirst - sg.c_lo))) & 15) == 0) {
            float4* o4 = reinterpret_cast<float4*>(o + (cfirst - sg.c_lo));
#pragma unroll
            for (int j = 0; j < 8; ++j) o4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int c = cfirst + j;
              if (c >= c_lo && c < c_hi) o[(size_t)(c - sg.c_lo) * sg.f32_sC] = v[j];
            }
          }
        }
        if (sg.out_b16 != nullptr) {
          __nv_bfloat16* ob = reinterpret_cast<__nv_bfloat16*>(sg.out_b16) +
                              ((size_t)img * d.H * d.W + (size_t)py * d.W + px) * sg.b16_ld + sg.b16_coff;
          const int cfirst = n0 + cb * 32;
          for (int p = 0; p < sg.b16_parts; ++p) {
            __nv_bfloat16* op = ob + (size_t)p * sg.b16_part_stride;
            if (cfirst >= c_lo && cfirst + 32 <= c_hi &&
                (((uintptr_t)(op + (cfirst - sg.c_lo))) & 15) == 0) {
              uint4* o4 = reinterpret_cast<uint4*>(op + (cfirst - sg.c_lo));
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                __nv_bfloat162 h[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  h[q] = __floats2bfloat162_rn(v[8 * j + 2 * q], v[8 * j + 2 * q + 1]);
                  v[8 * j + 2 * q] -= __low2float(h[q]);       // residue feeds the next part
                  v[8 * j + 2 * q + 1] -= __high2float(h[q]);
                }
                o4[j] = *reinterpret_cast<uint4*>(h);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int c = cfirst + j;
                const __nv_bfloat16 h = __float2bfloat16_rn(v[j]);
                if (c >= c_lo && c < c_hi) op[c - sg.c_lo] = h;
                v[j] -= __bfloat162float(h);
              }
            }
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 5) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "n"(kTmemCols)
                 : "memory");
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

int conv2_launch(const dhd_conv_desc* d, void* encode, void* stream);   // conv_igemm2.cu
void* conv_encode_fn() { return (void*)encode_fn(); }                  // used by conv_wgrad.cu

static int conv_version() {
  const char* v = getenv("DHD_CONV_V");
  return v != nullptr && *v != 0 ? atoi(v) : 2;
}

}  // namespace dhd

using namespace dhd;

extern "C" int dhd_conv2d_fwd(const dhd_conv_desc* d, void* stream) {
  DHD_REQUIRE(d != nullptr, "conv desc is null");
  DHD_REQUIRE(d->in != nullptr && d->weight != nullptr, "null input / weight pointer");
  DHD_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0, "bad image shape");
  DHD_REQUIRE(d->Cin > 0 && d->Cin % kBlockK == 0, "Cin must be a multiple of 64");
  DHD_REQUIRE(d->Cout > 0, "bad Cout");
  DHD_REQUIRE(d->taps >= 1 && d->taps <= DHD_CONV_MAX_TAPS, "taps out of range");
  DHD_REQUIRE(d->n_terms >= 1 && d->n_terms <= DHD_CONV_MAX_TERMS, "n_terms out of range");
  DHD_REQUIRE(d->w_parts >= 1 && d->w_parts <= 3, "w_parts out of range");
  DHD_REQUIRE(d->bw > 0 && d->bh > 0 && d->bw * d->bh == kBlockM && d->bw <= 256 && d->bh <= 256,
              "tile box must cover exactly 128 pixels");
  DHD_REQUIRE(d->in_ld % 8 == 0 && d->in_coff % 8 == 0 && d->in_part_stride % 8 == 0,
              "input channel offsets must be multiples of 8 (16-byte TMA alignment)");
  DHD_REQUIRE(((uintptr_t)d->in & 15) == 0 && ((uintptr_t)d->weight & 15) == 0,
              "input / weight must be 16-byte aligned");
  DHD_REQUIRE(d->n_seg >= 1 && d->n_seg <= DHD_CONV_MAX_SEGS, "n_seg out of range");
  for (int e = 0; e < d->n_terms; ++e)
    DHD_REQUIRE(d->term_a[e] >= 0 && d->term_a[e] < 3 && d->term_b[e] >= 0 && d->term_b[e] < d->w_parts,
                "bad split term");
  for (int s = 0; s < d->n_seg; ++s) {
    const dhd_conv_seg& sg = d->seg[s];
    DHD_REQUIRE(sg.c_lo >= 0 && sg.c_hi <= d->Cout && sg.c_lo < sg.c_hi, "bad output segment");
    DHD_REQUIRE(sg.out_f32 != nullptr || sg.out_b16 != nullptr, "segment without an output");
    if (sg.act == DHD_ACT_SOFTMAX)
      DHD_REQUIRE(sg.c_lo / kBlockN == (sg.c_hi - 1) / kBlockN, "softmax range must sit in one 128-channel tile");
    if (sg.out_b16 != nullptr) DHD_REQUIRE(sg.b16_parts >= 1 && sg.b16_parts <= 3, "bad b16_parts");
  }
  EncodeTiledFn enc = encode_fn();
  if (enc == nullptr) return fail(DHD_EUNSUPPORTED, "%s", "cuTensorMapEncodeTiled is unavailable");

  bool centre = false;
  for (int t = 0; t < d->taps; ++t) centre |= d->tap_dx[t] == 0 && d->tap_dy[t] == 0;
  DHD_REQUIRE(centre, "the filter must contain the (0, 0) tap");
  if (d->residual != nullptr)
    DHD_REQUIRE(d->Cout % 32 == 0 && ((uintptr_t)d->residual & 15) == 0 && d->res_sN % 4 == 0 &&
                    d->res_sY % 4 == 0 && d->res_sX % 4 == 0,
                "residual needs Cout % 32 == 0 and 16-byte aligned rows");
  DHD_REQUIRE(d->stride == 0 || d->stride == 1 || d->stride == 2, "stride must be 1 or 2");
  if (d->stride == 2) {
    DHD_REQUIRE(d->in_H > 0 && d->in_W > 0 && d->H <= (d->in_H + 1) / 2 && d->W <= (d->in_W + 1) / 2,
                "stride 2: output size must be at most ceil(input / 2)");
    DHD_REQUIRE(d->bw * 2 <= 256 && d->bh * 2 <= 256, "stride 2: tile box too large for the strided TMA box");
  }
  if (d->mix_x != nullptr) {
    DHD_REQUIRE(d->mix_a1 != nullptr && d->img_gate == nullptr && d->Cout % 32 == 0 && d->mix_ld % 8 == 0 &&
                    d->mix_coff % 8 == 0 && d->mix_part_stride % 8 == 0 && ((uintptr_t)d->mix_x & 15) == 0 &&
                    d->mix_parts >= 1 && d->mix_parts <= 3 && (d->stride == 0 || d->stride == 1),
                "sfa mix epilogue: Cout % 32 == 0, 16-byte aligned [bev | vox] rows, no img_gate");
    DHD_REQUIRE(conv_version() != 1, "the sfa mix epilogue needs the second-generation kernel");
  }
  if (d->stride != 2 && (d->in_H > 0 || d->in_W > 0))
    DHD_REQUIRE(d->in_H >= d->H && d->in_W >= d->W && conv_version() != 1,
                "in_H / in_W (input grid larger than the output grid) need the second-generation kernel");
  bool strided_out = false;
  for (int s = 0; s < d->n_seg; ++s) strided_out |= d->seg[s].b16_sX != 0;
  if (conv_version() != 1) return conv2_launch(d, (void*)enc, stream);
  DHD_REQUIRE(d->stride != 2 && !strided_out, "stride 2 / strided outputs need the second-generation kernel");

  CUtensorMap map_a, map_b;
  {
    cuuint64_t dims[4] = {(cuuint64_t)d->in_ld, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
    cuuint64_t strides[3] = {(cuuint64_t)d->in_ld * 2, (cuuint64_t)d->W * d->in_ld * 2,
                             (cuuint64_t)d->H * d->W * d->in_ld * 2};
    cuuint32_t box[4] = {(cuuint32_t)kBlockK, (cuuint32_t)d->bw, (cuuint32_t)d->bh, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&map_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)d->in, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(DHD_EINVAL, "%s: %ld", "cuTensorMapEncodeTiled(A) failed", (long)r);
  }
  {
    const cuuint64_t ktot = (cuuint64_t)d->taps * d->w_parts * d->Cin;
    cuuint64_t dims[2] = {ktot, (cuuint64_t)d->Cout};
    cuuint64_t strides[1] = {ktot * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)kBlockN};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&map_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)d->weight, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(DHD_EINVAL, "%s: %ld", "cuTensorMapEncodeTiled(B) failed", (long)r);
  }
  ConvKernelParams P;
  P.d = *d;
  P.tiles_w = (d->W + d->bw - 1) / d->bw;
  P.tiles_h = (d->H + d->bh - 1) / d->bh;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)kConvSmem);
    if (e != cudaSuccess) return fail((int)e, "%s: %ld", "cudaFuncSetAttribute(conv_igemm)", (long)e);
    attr_set = true;
  }
  dim3 grid(P.tiles_w * P.tiles_h * d->N, (d->Cout + kBlockN - 1) / kBlockN);
  conv_igemm_kernel<<<grid, kConvThreads, kConvSmem, (cudaStream_t)stream>>>(map_a, map_b, P);
  DHD_CUDA_LAUNCH_CHECK("conv_igemm");
  return DHD_OK;
}
