// Host side of the implicit-GEMM 2D convolution / linear layer on the sm_100a tensor cores: argument validation of
// dhd_conv2d_fwd and the cuTensorMapEncodeTiled entry point; the kernel itself is conv_igemm2.cu (the first-generation
// kernel that lived here -- one 128x128 tile per CTA, 4 epilogue warps -- was removed in round 2; git history keeps it).
//
// One kernel serves every dense layer of the DHD hot path (reference call sites:
//   MGHS.depth_net 1x1            models/necks/lss_heightmap.py:62, 482-485
//   HeightNet / DepthNet convs    models/model_utils/depthnet.py:172-243, 418-487 (3x3, dilated 3x3, 1x1)
//   SFA 1x1 / 3x3 convs           models/necks/mix.py:28-33, 74-85
//   predictor conv + MLP          models/dense_heads/occ_head.py:52-67
// which the reference runs as unfused cuDNN / cuBLAS launches).
//
// Formulation.  Activations are NHWC bf16 in HBM; a 128-pixel output tile is a (bh x bw) box of one image.  For every
// filter tap the A operand of the GEMM is that same box shifted by the tap offset, fetched by ONE 4-D TMA box load whose
// out-of-bounds rows / columns are zero-filled by the TMA unit -- padding, dilation and image borders cost no
// instructions and no im2col buffer exists.  Weights are [Cout][tap][part][Cin] bf16 (K-major) and arrive by 2-D TMA.
//
// Precision.  `n_terms` > 1 selects split-bf16 arithmetic: an fp32 tensor is carried as up to three bf16 "parts"
// (x = x0 + x1 + x2, stacked on the channel axis) and the K loop issues one MMA per (activation part, weight part) pair
// listed in term_a/term_b.  1 term = plain bf16; 3 terms ~ 2^-16; 6 terms ~ fp32.
#include <stdlib.h>

#include <cuda.h>

#include "common.cuh"

namespace dhd {

constexpr int kBlockM = 128;   // pixels per tile
constexpr int kBlockN = 128;   // softmax segments must sit inside one 128-channel tile
constexpr int kBlockK = 64;    // bf16 elements = one 128-byte swizzle row

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
int conv2_launch_batch(const dhd_conv_desc* const* descs, int n, void* encode, void* stream);
int conv2_pair_mode(int set);
void* conv_encode_fn() { return (void*)encode_fn(); }                  // used by conv_wgrad.cu

}  // namespace dhd

using namespace dhd;

static int validate_conv(const dhd_conv_desc* d) {
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
  bool centre = false;
  for (int t = 0; t < d->taps; ++t) centre |= d->tap_dx[t] == 0 && d->tap_dy[t] == 0;
  DHD_REQUIRE(centre, "the filter must contain the (0, 0) tap");
  if (d->residual != nullptr)
    DHD_REQUIRE(d->Cout % 32 == 0 && ((uintptr_t)d->residual & 15) == 0 && d->res_sN % 4 == 0 &&
                    d->res_sY % 4 == 0 && d->res_sX % 4 == 0,
                "residual needs Cout % 32 == 0 and 16-byte aligned rows");
  if (d->res_b16 != nullptr)
    DHD_REQUIRE(d->Cout % 32 == 0 && ((uintptr_t)d->res_b16 & 15) == 0 && d->res_b16_ld % 8 == 0 &&
                    d->res_b16_coff % 8 == 0 && d->residual == nullptr && (d->stride == 0 || d->stride == 1),
                "bf16 residual needs Cout % 32 == 0, 16-byte aligned rows, stride 1 and no fp32 residual");
  DHD_REQUIRE(d->w_image_rows == 0 || d->w_image_rows >= d->Cout, "w_image_rows must be 0 or >= Cout");
  if (d->stat_partial != nullptr) {
    const dhd_conv_seg& sg = d->seg[0];
    DHD_REQUIRE(d->n_seg == 1 && sg.c_lo == 0 && sg.c_hi == d->Cout && sg.out_b16 != nullptr && sg.out_f32 == nullptr &&
                    sg.b16_parts == 1 && sg.act != DHD_ACT_SOFTMAX && d->mix_x == nullptr && d->Cout % 2 == 0 &&
                    sg.b16_ld % 8 == 0 && sg.b16_coff % 8 == 0 && sg.b16_sX % 8 == 0 && sg.b16_sY % 8 == 0 &&
                    sg.b16_sN % 8 == 0 && ((uintptr_t)sg.out_b16 & 15) == 0 && ((uintptr_t)d->stat_partial & 7) == 0,
                "fused statistics: one single-part bf16 output over all channels (16-byte aligned rows), even Cout");
  }
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
  }
  if (d->stride != 2 && (d->in_H > 0 || d->in_W > 0))
    DHD_REQUIRE(d->in_H >= d->H && d->in_W >= d->W, "in_H / in_W must be at least the output grid");
  return DHD_OK;
}

extern "C" int dhd_conv2d_fwd(const dhd_conv_desc* d, void* stream) {
  const int rc = validate_conv(d);
  if (rc != DHD_OK) return rc;
  EncodeTiledFn enc = encode_fn();
  if (enc == nullptr) return fail(DHD_EUNSUPPORTED, "%s", "cuTensorMapEncodeTiled is unavailable");
  return conv2_launch(d, (void*)enc, stream);
}

extern "C" int dhd_conv2d_fwd_batch(const dhd_conv_desc* descs, int n, void* stream) {
  DHD_REQUIRE(descs != nullptr && n >= 1 && n <= DHD_CONV_MAX_BATCH, "batch of 1..DHD_CONV_MAX_BATCH descriptors");
  const dhd_conv_desc* ptrs[DHD_CONV_MAX_BATCH];
  for (int i = 0; i < n; ++i) {
    const int rc = validate_conv(descs + i);
    if (rc != DHD_OK) return rc;
    ptrs[i] = descs + i;
  }
  EncodeTiledFn enc = encode_fn();
  if (enc == nullptr) return fail(DHD_EUNSUPPORTED, "%s", "cuTensorMapEncodeTiled is unavailable");
  if (n == 1) return conv2_launch(ptrs[0], (void*)enc, stream);
  return conv2_launch_batch(ptrs, n, (void*)enc, stream);
}

extern "C" int dhd_conv_pair_mode(int mode) { return conv2_pair_mode(mode); }

extern "C" int dhd_conv2d_stat_rows(const dhd_conv_desc* d) {
  if (d == nullptr || d->bw <= 0 || d->bh <= 0) return 0;
  return ((d->W + d->bw - 1) / d->bw) * ((d->H + d->bh - 1) / d->bh) * d->N;
}
