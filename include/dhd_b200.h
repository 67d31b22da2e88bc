/* dhd_b200 -- C-ABI of the B200-native DHD view-transform / voxel-pool hot path.
 *
 * Plain pointers and sizes only (no torch types).  Every pointer is a DEVICE pointer
 * unless its name ends in _host.  Every function returns 0 on success or a negative
 * DHD_E* / positive cudaError_t code; dhd_last_error() returns a static message.
 * All work is enqueued on `stream` (a cudaStream_t passed as void*), nothing
 * synchronises, nothing allocates: the caller owns outputs and the workspace
 * (reference ownership model: projects/mmdet3d_plugin/ops/bev_pool_v2/bev_pool.py:27,67-68).
 *
 * Reference interfaces replaced (paths relative to the reference root):
 *   dhd_bev_pool_v2_fwd   <- bev_pool_v2_forward  ops/bev_pool_v2/src/bev_pool.cpp:30-57
 *                            (kernel ops/bev_pool_v2/src/bev_pool_cuda.cu:21-50, launch 127-133)
 *   dhd_bev_pool_v2_bwd   <- bev_pool_v2_backward ops/bev_pool_v2/src/bev_pool.cpp:74-104
 *                            (kernel bev_pool_cuda.cu:69-123, launch 135-142)
 *   dhd_height_to_mask    <- MGHS.height_feature_to_height_map + create_mask_3
 *                            models/necks/lss_heightmap.py:528-564
 *   dhd_mghs_prepare      <- MGHS.get_ego_coor + voxel_pooling_prepare_v2 (x4 passes)
 *                            models/necks/lss_heightmap.py:179-231, 303-371
 *   dhd_mghs_pool_fwd     <- masked features + 4x (new_zeros + bev_pool_v2 + permute + collapse-Z cat)
 *                            models/necks/lss_heightmap.py:261-300, 407-459; bev_pool.py:17-41,105
 *   dhd_mghs_pool_bwd     <- QuickCumsumCuda.backward x4 + mask product backward
 *                            ops/bev_pool_v2/bev_pool.py:44-83
 *   dhd_conv2d_fwd        <- nn.Conv2d / nn.Linear of depth_net, HeightNet, SFA, predictor
 *                            (see the dense-layer section below)
 *   dhd_stereo_cost_volume <- DepthNet.gen_grid + calculate_cost_volumn
 *                            models/model_utils/depthnet.py:245-361
 *   dhd_predictor_tail    <- predictor.predicter (Linear, Softplus, Linear) + permute + get_occ argmax
 *                            models/dense_heads/occ_head.py:63-67, 84-100, 141-153
 *   dhd_mghs_voxel_index  <- the (kept, ranks_bev) part of voxel_pooling_prepare_v2
 *                            models/necks/lss_heightmap.py:331-354 (bit-exact parity hook)
 */
#ifndef DHD_B200_H_
#define DHD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DHD_ABI_VERSION 1
#define DHD_MAX_PASSES 4
#define DHD_MAX_PLANES 32   /* sum of dz over passes */

enum {
  DHD_OK = 0,
  DHD_EINVAL = -1,      /* bad argument (shape, null pointer, unsupported size) */
  DHD_EUNSUPPORTED = -2 /* valid request this build does not implement */
};

/* Output memory layouts of the fused pool. Logical result of pass p is always the
 * reference's (B, dz*C, Dy, Dx) [collapse_z=True, channel = z*C + c, lss_heightmap.py:298-299]
 * or (B, C, dz, Dy, Dx) [collapse_z=False]; the layout says how it sits in HBM. */
enum {
  DHD_LAYOUT_NHWC = 0,          /* (b, y, x, z, c): channels_last of the collapsed tensor; fastest */
  DHD_LAYOUT_NCHW_COLLAPSE = 1, /* (b, z, c, y, x): contiguous (B, dz*C, Dy, Dx), the reference layout */
  DHD_LAYOUT_NCDHW = 2,         /* (b, c, z, y, x): contiguous (B, C, dz, Dy, Dx), collapse_z=False */
  DHD_LAYOUT_NCDHW_CAT = 3,     /* pass 0 as NCDHW into out[0]; passes 1.. stacked on z in ONE tensor out[1] of
                                   shape (B, C, sum dz[1..], Dy, Dx): MGHS_Depth's torch.cat(dim=2), lss_heightmap.py:845 */
  DHD_LAYOUT_NHWC_BF16 = 4      /* DHD_LAYOUT_NHWC with bf16 elements (out_host[p] point at bf16): the activation layout
                                   the encoder convolutions read -- half the bytes, no conversion pass; forward only */
};

/* One fused view-transform problem: a frustum of B*N*D*fH*fW points pooled into
 * n_pass voxel grids that share the x/y grid and differ in z range and pixel mask. */
typedef struct dhd_mghs_cfg {
  int32_t B, N, D, fH, fW, C;        /* C must be 64 (DHD numC_Trans) in this build */
  int32_t Dx, Dy;                    /* int(grid_size) of the shared BEV grid */
  float x_lower, x_interval, x_size; /* fp32 values of create_grid_infos (lss_heightmap.py:99-102) */
  float y_lower, y_interval, y_size;
  int32_t n_pass;                    /* 1..DHD_MAX_PASSES */
  float z_lower[DHD_MAX_PASSES];
  float z_interval[DHD_MAX_PASSES];
  float z_size[DHD_MAX_PASSES];      /* fp32 grid_size[2]; kept test is (float)idx < z_size */
  int32_t dz[DHD_MAX_PASSES];        /* int(z_size) */
  int32_t mask_id[DHD_MAX_PASSES];   /* 0: every pixel contributes; k>0: only pixels with mask == k */
} dhd_mghs_cfg;

const char* dhd_last_error(void);
int dhd_abi_version(void);
/* sizeof(dhd_mghs_cfg | dhd_conv_seg | dhd_conv_desc | dhd_wgrad_desc | dhd_stereo_desc | dhd_predictor_tail_desc) for which = 0..5: lets a binding check its
 * own struct layout against the library it loaded */
size_t dhd_abi_sizeof(int which);

/* ---- drop-in operator (reference tensor contracts, fp32 / int32) ------------------
 * depth (B,N,D,fH,fW), feat (B,N,fH,fW,C), out (B,Dz,Dy,Dx,C) PRE-ZEROED by the caller;
 * one interval = one run of points with the same ranks_bev. */
int dhd_bev_pool_v2_fwd(int c, int n_intervals, const float* depth, const float* feat,
                        const int32_t* ranks_depth, const int32_t* ranks_feat,
                        const int32_t* ranks_bev, const int32_t* interval_starts,
                        const int32_t* interval_lengths, float* out, void* stream);
/* intervals here are runs of equal ranks_feat over points re-sorted by ranks_feat
 * (bev_pool.py:47-57); depth_grad / feat_grad PRE-ZEROED by the caller. */
int dhd_bev_pool_v2_bwd(int c, int n_intervals, const float* out_grad, const float* depth,
                        const float* feat, const int32_t* ranks_depth, const int32_t* ranks_feat,
                        const int32_t* ranks_bev, const int32_t* interval_starts,
                        const int32_t* interval_lengths, float* depth_grad, float* feat_grad,
                        void* stream);

/* ---- height distribution -> per-pixel mask id -------------------------------------
 * height (BN, H, fH*fW) fp32; height_range (H) fp32; thresholds (n_mask+1) fp32:
 * mask k (1-based) iff thresholds[k-1] <= height_range[argmax] < thresholds[k]; else 0. */
int dhd_height_to_mask(const float* height, int BN, int H, int HW, const float* height_range,
                       const float* thresholds, int n_mask, int8_t* pixmask, void* stream);

/* ---- fused MGHS pool -----------------------------------------------------------------
 * Workspace: dhd_mghs_workspace_bytes() bytes of device memory, 256B-aligned, owned by the
 * caller; filled by dhd_mghs_prepare, read by pool_fwd / pool_bwd / voxel_index. It depends
 * only on the camera geometry, so it may be cached across calls (MGHS `accelerate`). */
size_t dhd_mghs_workspace_bytes(const dhd_mghs_cfg* cfg);

/* Geometry source: if `coor` is non-null it is the (B,N,D,fH,fW,3) fp32 ego coordinates
 * (the reference's get_ego_coor result) and the camera arguments are ignored; otherwise the
 * coordinates are computed in-kernel from frustum_{u,v,d} (fW / fH / D fp32 values of
 * create_frustum), inv_post_rot (BN,3,3), post_tran (BN,3), combine (BN,3,3) =
 * sensor2ego[:3,:3] @ inv(K), trans (BN,3), bda (B,3,3) with the reference's fp32 operation
 * order (separately rounded multiplies and adds). */
int dhd_mghs_prepare(const dhd_mghs_cfg* cfg, const float* coor, const float* frustum_u,
                     const float* frustum_v, const float* frustum_d, const float* inv_post_rot,
                     const float* post_tran, const float* combine, const float* trans,
                     const float* bda, void* workspace, int deterministic, void* stream);

/* depth (B,N,D,fH,fW) fp32; feat (B,N,fH,fW,C) fp32; pixmask (B*N*fH*fW) int8 or NULL when no
 * pass is masked; out[p] written completely (zeros included) in `layout`. */
int dhd_mghs_pool_fwd(const dhd_mghs_cfg* cfg, const float* depth, const float* feat,
                      const int8_t* pixmask, const void* workspace, float* const* out_host,
                      int layout, void* stream);

/* gout[p] in `layout`; depth_grad (B,N,D,fH,fW) and feat_grad (B,N,fH,fW,C) are written
 * completely (no pre-zeroing needed). Only DHD_LAYOUT_NHWC in this build. */
int dhd_mghs_pool_bwd(const dhd_mghs_cfg* cfg, const float* depth, const float* feat,
                      const int8_t* pixmask, const void* workspace, const float* const* gout_host,
                      int layout, float* depth_grad, float* feat_grad, void* stream);

/* ranks_out (n_pass, B*N*D*fH*fW) int32: voxel rank b*(dz*Dy*Dx)+z*(Dy*Dx)+y*Dx+x of every
 * frustum point in every pass, or -1 where the reference's `kept` is false. */
int dhd_mghs_voxel_index(const dhd_mghs_cfg* cfg, const void* workspace, int32_t* ranks_out,
                         void* stream);

/* number of binned entries after prepare (device int32[1] copied by caller) lives at this
 * byte offset of the workspace */
size_t dhd_mghs_workspace_count_offset(const dhd_mghs_cfg* cfg);

/* ---- dense layers: implicit-GEMM convolution on tcgen05 tensor cores ------------------
 * Replaces the cuDNN / cuBLAS calls behind nn.Conv2d / nn.Linear in
 *   MGHS.depth_net                     models/necks/lss_heightmap.py:62, 482-485
 *   HeightNet / DepthNet / ASPP / SE   models/model_utils/depthnet.py:10-116, 150-243, 418-487
 *   SFA                                models/necks/mix.py:28-33, 74-85
 *   predictor.final_conv / predicter   models/dense_heads/occ_head.py:52-67
 * Input: NHWC bf16, `in_ld` channels per pixel; the layer reads Cin channels starting at
 * in_coff (+ part * in_part_stride for split-bf16 part `part`).  Weight: bf16
 * [Cout][taps][w_parts][Cin].  out = act(scale[c]*acc + bias[c] + img_bias[n,c] + residual)
 * * img_gate[n,c], written per output segment (a channel range of this layer). */
#define DHD_CONV_MAX_TAPS 9
#define DHD_CONV_MAX_TERMS 6
#define DHD_CONV_MAX_SEGS 2
#define DHD_CONV_MAX_BATCH 4   /* convolutions per dhd_conv2d_fwd_batch launch */

enum {
  DHD_ACT_NONE = 0,
  DHD_ACT_RELU = 1,
  DHD_ACT_SIGMOID = 2,
  DHD_ACT_SOFTPLUS = 3, /* torch.nn.Softplus(beta=1, threshold=20) */
  DHD_ACT_SOFTMAX = 4   /* over the channels of the segment (must sit in one 128-channel tile) */
};

typedef struct dhd_conv_seg {
  int32_t c_lo, c_hi;          /* output channels [c_lo, c_hi) of this layer */
  int32_t act;                 /* DHD_ACT_* */
  float* out_f32;              /* or NULL; element (n, y, x, c) at n*sN + y*sY + x*sX + (c-c_lo)*sC */
  int64_t f32_sN, f32_sY, f32_sX, f32_sC;
  void* out_b16;               /* or NULL; NHWC bf16, b16_ld channels per pixel, channel c-c_lo at
                                  b16_coff + part*b16_part_stride + (c-c_lo) */
  int32_t b16_ld, b16_coff, b16_parts, b16_part_stride;
  int64_t b16_sN, b16_sY, b16_sX; /* output pixel strides in bf16 elements; all 0 = dense NHWC (H*W*ld, W*ld, ld).
                                     A strided view lets a layer write every 2nd pixel of a larger map: the four
                                     phases of ConvTranspose2d(kernel 2, stride 2) (backbones/unet.py:86) */
} dhd_conv_seg;

typedef struct dhd_conv_desc {
  int32_t N, H, W;             /* images and OUTPUT spatial size (== input size unless stride == 2) */
  int32_t Cin, Cout;           /* Cin % 64 == 0 */
  int32_t taps;                /* 1 (1x1 / linear) .. 9 */
  int32_t tap_dy[DHD_CONV_MAX_TAPS], tap_dx[DHD_CONV_MAX_TAPS]; /* input offset of tap t: (k-1)*dilation */
  int32_t bw, bh;              /* output tile box, bw*bh == 128 */
  const void* in;              /* bf16 NHWC */
  int32_t in_ld, in_coff, in_part_stride;
  const void* weight;          /* bf16 [Cout][taps][w_parts][Cin] */
  int32_t w_parts;
  int32_t n_terms;             /* MMAs per (tap, 64-channel chunk): 1 = bf16, 3 / 6 = split-bf16 */
  int32_t term_a[DHD_CONV_MAX_TERMS], term_b[DHD_CONV_MAX_TERMS];
  const float* scale;          /* [Cout] or NULL (folded BatchNorm scale) */
  const float* bias;           /* [Cout] or NULL */
  const float* img_bias;       /* [N][Cout] or NULL */
  const float* img_gate;       /* [N][Cout] or NULL, applied after the activation */
  const float* residual;       /* fp32 (n,y,x,c) at n*res_sN + y*res_sY + x*res_sX + c, or NULL */
  int64_t res_sN, res_sY, res_sX;
  int32_t n_seg;
  dhd_conv_seg seg[DHD_CONV_MAX_SEGS];
  int32_t stride;              /* 0 / 1: same-size convolution; 2: N,H,W describe the OUTPUT, the input is in_H x in_W
                                  and tap t reads input pixel (stride*y + tap_dy, stride*x + tap_dx)
                                  (CustomResNet's stride-2 blocks, backbones/resnet.py:47-52) */
  int32_t in_H, in_W;          /* input grid when it differs from the output grid: required for stride == 2; for
                                  stride 1 an optional larger input grid (the output covers its top-left H x W) */
  /* SFA spatial-gate blend fused into the epilogue (mix.py:52-57): with g = the activated output (the sigmoid
   * gate a2), the layer writes  g * a1*bev + (1 - g) * (1 - a1)*vox  instead of g, where [bev | vox] = mix_x
   * (bf16 NHWC, mix_parts split parts, bev at mix_coff, vox at mix_coff + Cout) and a1 = mix_a1 [N][Cout].
   * NULL = off.  Saves the round trip of the gate tensor through HBM and the separate blend pass. */
  const void* mix_x;
  int32_t mix_ld, mix_coff, mix_parts, mix_part_stride;
  const float* mix_a1;
  /* Per-image weights: 0 = one weight matrix for every image; > 0 = `weight` holds N matrices of w_image_rows
   * (>= Cout) rows each, image n convolves with rows [n*w_image_rows, n*w_image_rows + Cout).  This is how SFA's
   * channel gate (mix.py:41-50: u = a1*bev + (1-a1)*vox, a1 per image and channel) is folded into the 1x1
   * convolution that consumes u: W*u = [W diag(a1) | W diag(1-a1)] * [bev ; vox] (dhd_sfa_fold_gate). */
  int32_t w_image_rows;
  /* bf16 residual (same pixel grid as the output, channel c at res_b16_coff + c of res_b16_ld), added to the
   * pre-activation value like `residual`; the bf16 speed mode's identity paths.  NULL = off. */
  int32_t res_b16_ld, res_b16_coff;
  const void* res_b16;
  /* BatchNorm batch statistics fused into the epilogue (training mode: the convolution writes its raw bf16 output and
   * torch's BatchNorm needs sum / sum of squares per channel over all pixels).  NULL = off.  Otherwise fp32
   * [dhd_conv2d_stat_rows(desc)][2][Cout]: per-tile partial sums of the bf16-rounded output, [.][0][c] = sum,
   * [.][1][c] = sum of squares, written completely; dhd_colsum_finish adds the rows in fixed order.  Requires one
   * single-part bf16 output segment over all channels, no activation-dependent consumers, Cout even. */
  float* stat_partial;
} dhd_conv_desc;

int dhd_conv2d_fwd(const dhd_conv_desc* desc, void* stream);
/* n (<= DHD_CONV_MAX_BATCH) independent convolutions in ONE persistent launch: `descs` is an array of n descriptors
 * (same contract as dhd_conv2d_fwd each).  For layers that are individually too small to fill the GPU and do not
 * depend on each other: the four ASPP branches (depthnet.py:87-106), the groups of the deformable convolution,
 * depth_net next to HeightNet's reduce_conv (lss_heightmap.py:482-487 -- both read the image feature).  Outputs of
 * one problem must not be inputs of another. */
/* rows of dhd_conv_desc.stat_partial for this layer (one per 128-pixel tile) */
int dhd_conv2d_stat_rows(const dhd_conv_desc* desc);
/* sums[i] = sum over rows of partial[row][i], i < n, rows added in a fixed order (deterministic) */
int dhd_colsum_finish(const float* partial, int rows, int n, float* sums, void* stream);
int dhd_conv2d_fwd_batch(const dhd_conv_desc* descs, int n, void* stream);
/* Layers with Cout > 128 and shared weights run on CTA PAIRS (clusters of two CTAs, tcgen05.mma.cta_group::2, M = 256:
 * each CTA stages its own 128 pixels and half of the weight tile) where that pays: K = taps*Cin >= 1024 and at
 * least four tiles per SM (mode 1, the default).  mode 0: never; mode 2: wherever the kernel supports it; mode < 0
 * only queries.  Returns the previous setting (initially 1, or the DHD_CONV_CTA2 environment variable).  A/B switch
 * for tests and benchmarks -- results are bit-identical either way. */
int dhd_conv_pair_mode(int mode);

/* ---- backward of the dense layers ------------------------------------------------------
 * (torch autograd of nn.Conv2d / nn.Linear in the modules listed above; the reference trains them
 * through cuDNN's backward-data / backward-filter.)
 * Data gradient: dx = conv(dy, w') with w'[ci][flipped tap][co] = w[co][tap][ci] -- the same
 * dhd_conv2d_fwd kernel on a weight tensor repacked once per step by the host.
 * Weight gradient: dw[co][tap][ci] = scale[co] * sum_pixels dy[p][co] * x[p + tap][ci], bf16
 * operands, fp32 accumulation in TMEM.  `partial` is a caller-owned workspace of
 * dhd_conv2d_wgrad_workspace_bytes() bytes; dw is fp32 [Cout][taps][Cin] (the layout of the packed
 * forward weight), written completely, or added to when accumulate != 0. */
typedef struct dhd_wgrad_desc {
  int32_t N, H, W;
  int32_t Cin, Cout;           /* Cin % 64 == 0 */
  int32_t taps;
  int32_t tap_dy[DHD_CONV_MAX_TAPS], tap_dx[DHD_CONV_MAX_TAPS];
  int32_t bw, bh;              /* pixel box, bw*bh == 128 */
  const void* x;               /* layer input, bf16 NHWC: channel c at x_coff + c of x_ld */
  int32_t x_ld, x_coff;
  const void* dy;              /* gradient w.r.t. the convolution output, bf16 NHWC */
  int32_t dy_ld, dy_coff;
  const float* scale;          /* [Cout] or NULL (folded BatchNorm scale of the forward epilogue) */
  float* dw;
  float* partial;
  int32_t accumulate;
  int32_t x_stride;            /* 0 / 1, or 2: x is sampled at (2*y + tap_dy, 2*x + tap_dx) of an x_H x x_W grid
                                  (stride-2 layers; ConvTranspose2d(2,2) with the operand roles swapped) */
  int32_t x_H, x_W;
  /* Output layout.  dw_torch == 0: dw is [Cout][taps][Cin] (the packed forward layout).  dw_torch != 0: dw is the
   * nn.Conv2d / nn.Linear gradient tensor itself, [Cout][dw_cin_total][taps]; input channel ci < dw_cin_used lands
   * in column dw_cin_lo + ci and the zero-padded channels ci >= dw_cin_used are dropped -- with accumulate != 0 the
   * kernel adds straight into param.grad (a view into the data-parallel gradient bucket), no repack / add launches. */
  int32_t dw_torch, dw_cin_total, dw_cin_lo, dw_cin_used;
} dhd_wgrad_desc;

size_t dhd_conv2d_wgrad_workspace_bytes(const dhd_wgrad_desc* desc);
int dhd_conv2d_wgrad(const dhd_wgrad_desc* desc, void* stream);

/* ---- streaming kernels of the training path (csrc/train.cu) ----------------------------------
 * dz = dy * act'(y) on bf16 NHWC rows (y = the layer's saved OUTPUT: relu y>0, sigmoid y(1-y),
 * softplus 1-exp(-y)); out may alias dy or be NULL.  colsum (2*C floats, or NULL): [0,C) = sum_rows dz
 * (bias / BatchNorm beta gradient), [C,2C) = sum_rows dz*y (BatchNorm gamma gradient after the
 * affine fix-up); needs workspace of dhd_act_bwd_workspace_bytes(C); sums are deterministic. */
size_t dhd_act_bwd_workspace_bytes(int C);
int dhd_act_bwd(const void* dy, int dy_ld, int dy_coff, const void* y, int y_ld, int y_coff, long rows,
                int C, int act, void* out, int out_ld, int out_coff, float* colsum, float* workspace,
                const void* add, int add_ld, int add_coff, void* stream);
/* (add, optional: a second bf16 gradient summed into dy first -- the identity path of a residual block) */
/* predictor.loss (occ_head.py:102-131): class-weighted masked cross-entropy (mmdet CrossEntropyLoss with
 * class_weight, weight = mask_camera, avg_factor = sum mask*class_weight[label]) and, when weight_sem / weight_geo
 * are non-zero, sem_scal_loss_with_mask + geo_scal_loss_with_mask (losses/semkitti_loss.py:136-225; needs
 * workspace of dhd_occ_loss_workspace_bytes()).  logits (B,Dx,Dy,Dz,ncls) fp32, labels / mask (B,Dx,Dy,Dz) uint8
 * (mask or class_weight may be NULL).  losses[0] = loss_occ, [1] = avg_factor, [2] = loss_voxel_sem_scal,
 * [3] = loss_voxel_geo_scal; dlogits = d(sum of the terms)/d logits as bf16 NHWC rows [(b*Dy+y)*Dx+x][dl_ld],
 * channel z*ncls+k -- the input layout of the last Linear's backward GEMMs. */
size_t dhd_occ_loss_workspace_bytes(void);
int dhd_occ_ce_loss(const float* logits, const uint8_t* labels, const uint8_t* mask,
                    const float* class_weight, int ncls, int ignore_index, int B, int Dx, int Dy, int Dz,
                    float loss_weight, float weight_sem, float weight_geo, int non_empty_idx,
                    float* losses, void* dlogits, int dl_ld, float* workspace, void* stream);
/* backward of MGHS.depth_net's output head (lss_heightmap.py:482-489): depth = softmax over D (NCHW),
 * depth_grad / feat_grad from dhd_mghs_pool_bwd -> gradient w.r.t. the 1x1 convolution output, one bf16
 * NHWC row per pixel: [0,D) softmax backward, [D,D+C) feat_grad, [D+C,out_ld) zeros. */
int dhd_depth_head_bwd(const float* depth, const float* depth_grad, const float* feat_grad, int BN, int D,
                       int HW, int C, void* out, int out_ld, void* stream);

/* backward of the SFA gates (mix.py:37-59; formulas at sfa_gate_bwd_kernel in csrc/train.cu).
 * mode 0: g = d fuse -> dpre2 (bf16 NHWC, gradient at the spatial gate's sigmoid input), dx (fp32
 * [pix][2C], written), a1_sums [N][C] (d a1 from this blend); mode 1: g = d u -> dx accumulated,
 * a1_sums (accumulated when accumulate_sums != 0).  x = [bev | vox] bf16 NHWC (2C channels at x_coff). */
size_t dhd_sfa_gate_bwd_workspace_bytes(int N, int HW, int C);
int dhd_sfa_gate_bwd(int mode, const void* g, int g_ld, int g_coff, const void* x, int x_ld, int x_coff,
                     int C, int N, int HW, const float* a1, const float* a2, void* dpre2, int d_ld,
                     int d_coff, float* dx, float* a1_sums, int accumulate_sums, float* workspace,
                     void* stream);
/* bf16 form of the same (the training step's): mode 0 writes dx as a bf16 activation ([pix][dx_ld], bev at dx_coff,
 * vox at dx_coff + C) instead of 2C fp32 values per pixel; mode 1 only reduces a1_sums (dx may be NULL) -- its data
 * gradient a1*du | (1-a1)*du needs no x and is added by dhd_sfa_dx_combine:
 *   out = dxb + [a1 * du | (1 - a1) * du] + ds[n]   (ds: [N][2C] per-image row vector or NULL; out may alias dxb) */
int dhd_sfa_gate_bwd_b16(int mode, const void* g, int g_ld, int g_coff, const void* x, int x_ld, int x_coff, int C, int N,
                         int HW, const float* a1, const float* a2, void* dpre2, int d_ld, int d_coff, void* dx, int dx_ld,
                         int dx_coff, float* a1_sums, int accumulate_sums, float* workspace, void* stream);
int dhd_sfa_dx_combine(const void* dxb, int b_ld, int b_coff, const void* du, int u_ld, int u_coff, const float* a1,
                       const float* ds, int C, int N, int HW, void* out, int o_ld, int o_coff, void* stream);
/* out (bf16 NHWC) = in (fp32 [N*HW][C]) + v[n][c] (v may be NULL) */
int dhd_add_rowvec(const float* in, const float* v, int N, int HW, int C, void* out, int out_ld,
                   int out_coff, void* stream);

/* SE gate backward (depthnet.py:150-169, 624-629): h = relu(bn(conv x)) * gate[n][c] saved as bf16;
 * dpre = dh * gate * (h > 0); gate_sums [N][C] = sum_pixels dh * h / gate (= d gate); workspace as
 * dhd_sfa_gate_bwd_workspace_bytes(N, HW, C). */
int dhd_se_gate_bwd(const void* dh, int g_ld, int g_coff, const void* h, int h_ld, int h_coff, int C, int N,
                    int HW, const float* gate, void* dpre, int d_ld, int d_coff, float* gate_sums,
                    float* workspace, void* stream);
/* MGHS.get_downsampled_gt_depth / get_downsampled_gt_height (lss_heightmap.py:625-701): gt (BN, H, W) sparse fp32 map
 * -> per ds x ds block the minimum of the non-zero values, binned as (min - lo) / interval (depth: lo = d_min - d_step,
 * interval = d_step; height: lo = height_range[0], interval = height_interval), label = bin - 1 in [-1, nbins)
 * (-1: no valid return, the reference's all-zero one-hot row), valid (optional) = label >= 0 (the foreground test
 * of get_height_loss when applied to the depth map). */
int dhd_gt_downsample(const float* gt, int BN, int H, int W, int ds, float lo, float interval, int nbins,
                      int32_t* label, uint8_t* valid, void* stream);
/* MGHS.get_height_loss (lss_heightmap.py:595-622) on already binned labels: height (BN,H,HW) softmax
 * probabilities, label[pix] = GT height bin or -1 (all-zero one-hot row), fg[pix] = valid GT depth,
 * n_fg device scalar.  loss[0] = weight * sum_fg BCE / max(1, n_fg); dz = d loss / d logits (through
 * the softmax) as bf16 NHWC rows of dz_ld channels (zero beyond H and for background pixels). */
int dhd_height_loss(const float* height, const int32_t* label, const uint8_t* fg, int BN, int H, int HW,
                    float weight, const float* n_fg, float* loss, void* dz, int dz_ld, void* stream);
/* sampling part of the deformable convolution backward (mmcv DeformConv2dPack, depthnet.py:466-477):
 * dcol in the layout dhd_dcn_im2col writes; dx fp32 [N*H*W][C] (zeroed here, accumulated with atomics),
 * doff fp32 [N*H*W][off_ld] gradient of the offsets. */
int dhd_dcn_col2im_bwd(const void* dcol, int col_ld, const void* x, int x_ld, int x_coff, int C, int N, int H,
                       int W, const float* offset, int off_ld, int ksize, int pad, int dilation, int groups,
                       float* dx, float* doff, void* stream);

/* re-pack one layer's fp32 master weight w[Cout][cin_total][taps] (input columns [col_lo, col_lo+Cin)) after
 * an optimizer step: fwd = bf16 [Cout][taps][cin_pad] (dhd_conv2d_fwd operand), bwd = the data-gradient operand
 * with scale[co] folded in: mode 0 [Cin][mirrored tap][cout_pad], mode 1 [tap][Cin][cout_pad] (grouped 1x1 view
 * of the DCN weight).  Either output may be NULL. */
int dhd_pack_conv_weights(const float* w, int Cout, int cin_total, int taps, int col_lo, int Cin,
                          const float* scale, void* fwd, int cin_pad, void* bwd, int cout_pad, int bwd_mode,
                          void* stream);

/* the same for up to DHD_PACK_MAX_BATCH layers in ONE launch (the re-pack after an optimizer step is ~30 independent
 * few-microsecond kernels otherwise) */
#define DHD_PACK_MAX_BATCH 32
typedef struct dhd_pack_desc {
  const float* w;
  const float* scale;
  void* fwd;
  void* bwd;
  int32_t Cout, cin_total, taps, col_lo, Cin, cin_pad, cout_pad, bwd_mode;
} dhd_pack_desc;
int dhd_pack_conv_weights_batch(const dhd_pack_desc* descs, int n, void* stream);

/* AdamW (torch.optim.AdamW: decoupled weight decay, bias-corrected moments; the reference's optimizer,
 * projects/configs/DHD/DHD-S.py:262) over FLAT fp32 buffers: p, m, v updated in place from g, all of n elements.
 * bc1 = 1 - beta1^t, bc2 = 1 - beta2^t for this step t.  grad_scale: optional DEVICE scalar multiplied into g on
 * the fly (the gradient-clipping coefficient min(1, max_norm / (norm + 1e-6)): no separate scaling pass). */
int dhd_adamw_flat(float* p, const float* g, float* m, float* v, long n, float lr, float beta1, float beta2, float eps,
                   float weight_decay, float bc1, float bc2, const float* grad_scale, void* stream);

/* batch-statistics BatchNorm of the training path (torch BatchNorm2d in training mode after every convolution of the
 * reference's modules).  Forward: out = act(scale[c]*raw + shift[c] [+ residual]) [* gate[n][c]] on the convolution's
 * bf16 output `raw` (scale / shift from its batch statistics; the sums come from dhd_act_bwd(raw, raw, act none)).
 * Backward: out = k1[c]*a + k2[c]*b + k3[c] (a = dz, b = raw; out may alias a): the BatchNorm backward as a per-channel affine
 * combination (coefficients at bn_apply_kernel in csrc/train.cu). */
int dhd_bn_apply(const void* raw, int raw_ld, int raw_coff, long rows, int C, const float* scale,
                 const float* shift, int act, const float* residual, long res_ld, const float* gate,
                 int rows_per_img, void* out_b16, int o_ld, int o_coff, float* out_f32, long f_ld, void* stream);
/* the same forward pass with a bf16 identity path: out = act(scale[c]*raw + shift[c] + res16) (the Bottleneck blocks of
 * the image backbone keep their residual stream in bf16) */
int dhd_bn_apply_res16(const void* raw, int raw_ld, int raw_coff, long rows, int C, const float* scale, const float* shift,
                       int act, const void* res16, int res16_ld, int res16_coff, void* out_b16, int o_ld, int o_coff,
                       void* stream);
/* per-channel coefficients of the two passes from the reduced sums (one launch each): forward -- scale / shift (and
 * mean / invstd for the backward, running statistics updated with `momentum` when given) from sums = [sum raw,
 * sum raw^2]; backward -- k1 / k2 / k3 from sums = [sum dz, sum dz*raw] (second half at sums_stride), d gamma / d beta
 * accumulated when given. */
int dhd_bn_fwd_coeffs(const float* sums, int C, float M, const float* gamma, const float* beta, float eps,
                      float momentum, float* running_mean, float* running_var, float* scale, float* shift,
                      float* mean, float* invstd, void* stream);
int dhd_bn_bwd_coeffs(const float* sums, int C, int sums_stride, float M, const float* mean, const float* invstd,
                      const float* gamma, float* k1, float* k2, float* k3, float* dgamma, float* dbeta, void* stream);
/* the same coefficients straight from per-block partial sums (one launch = fixed-order finish + formulas):
 * forward from the convolution epilogue's statistics (dhd_conv_desc.stat_partial, rows = dhd_conv2d_stat_rows);
 * backward = the reduction pass over (dy, raw) + finish + formulas (workspace: dhd_act_bwd_workspace_bytes(C)). */
int dhd_bn_fwd_coeffs_partial(const float* partial, int rows, int C, float M, const float* gamma, const float* beta,
                              float eps, float momentum, float* running_mean, float* running_var, float* scale,
                              float* shift, float* mean, float* invstd, void* stream);
int dhd_bn_bwd_sums_coeffs(const void* dy, int dy_ld, int dy_coff, const void* raw, int raw_ld, int raw_coff, long rows,
                           int C, float* workspace, float M, const float* mean, const float* invstd, const float* gamma,
                           float* k1, float* k2, float* k3, float* dgamma, float* dbeta, void* stream);
/* nn.Dropout behind the ASPP (depthnet.py:81, 106) in training mode, in place on a bf16 NHWC activation (or on the
 * gradient at the same place in the backward): x[r][c] *= keep(r, c) / (1 - p), keep from Philox-4x32-10 keyed by
 * rng[0] (seed) with counter (element, rng[1] (step), salt).  rng is a DEVICE pointer to two int64: the caller bumps
 * the step once per iteration, also inside a captured graph.  Same (seed, step, salt) => same mask. */
int dhd_dropout(void* x, int ld, int coff, long rows, int C, float p, const long long* rng, unsigned salt,
                void* stream);
int dhd_affine_combine(const void* a, int a_ld, int a_coff, const void* b, int b_ld, int b_coff, long rows, int C,
                       const float* k1, const float* k2, const float* k3, void* out, int o_ld, int o_coff,
                       void* stream);

/* ---- plane-sweep stereo cost volume of the camera-aware DepthNet (csrc/stereo.cu) ----------
 * Replaces DepthNet.gen_grid + DepthNet.calculate_cost_volumn (models/model_utils/depthnet.py:245-308, 310-361):
 * the previous frame's 1/4-resolution stereo feature is warped to every depth hypothesis of every current pixel
 * (bilinear, zeros outside, align_corners=True), the L1 distance to the current feature is summed over the channels,
 * `bias` is added where the reference's "warped channel C-4 == 0" test fires, and softmax(-cost) over depth is
 * written.  Features are NHWC (BN, H, W, C), fp32 or bf16 (dhd_nchw_to_nhwc converts the reference's NCHW fp32).
 * Per image `cam` holds DHD_STEREO_CAM_FLOATS fp32 values, all row-major:
 *   [0..8]  inverse(post_rots)      [9..11]  post_trans      [12..20] k2s_sensor[:3,:3] @ inverse(intrins)
 *   [21..23] k2s_sensor[:3,3]       [24..32] intrins         [33..36] post_rots[:2,:2]     [37..38] post_trans[:2]
 * (the small inverses / products stay on the host side in torch, as in the reference, so they round identically). */
#define DHD_STEREO_CAM_FLOATS 40
typedef struct dhd_stereo_desc {
  int32_t BN, C, H, W, D;      /* images, stereo channels (multiple of 4, <= 512), map size, depth hypotheses (<= 128) */
  int32_t feat_bf16;           /* element type of prev / curr: 0 fp32, 1 bf16 */
  const void* prev;            /* (BN, H, W, C) previous frame, already aligned to the same camera order */
  const void* curr;            /* (BN, H, W, C) current frame */
  const float* frustum_u;      /* (W), (H), (D): the three axes of the (D, H, W, 3) = (u, v, d) template MGHS_Stereo.cv_frustum */
  const float* frustum_v;      /*   is built from (create_frustum, lss_heightmap.py:105-134: frustum[d][y][x] = */
  const float* frustum_d;      /*   (u[x], v[y], d[d])); unused when grid != NULL */
  const float* cam;            /* (BN, DHD_STEREO_CAM_FLOATS); unused when grid != NULL */
  const float* grid;           /* optional (BN, D*H, W, 2): normalised sampling coordinates computed elsewhere */
  float img_w, img_h;          /* wi, hi of gen_grid: 4 * W, 4 * H */
  float bias;                  /* DepthNet.bias (DHD-L: 5.0); 0 disables the test */
  float* out_f32;              /* optional fp32 result, element (bn, d, y, x) at bn*sN + d*sD + y*sY + x*sX */
  int64_t f32_sN, f32_sD, f32_sY, f32_sX;
  void* out_b16;               /* optional split-bf16 NHWC activation (input of cost_volumn_net) */
  int32_t b16_ld, b16_coff, b16_parts, b16_part_stride;
  int32_t b16_cpad;            /* channels written per part: D rounded up by the caller, the tail is zeroed */
  float* grid_out;             /* optional (BN, D*H, W, 2): the coordinates the kernel sampled at (parity hook) */
} dhd_stereo_desc;
int dhd_stereo_cost_volume(const dhd_stereo_desc* d, void* stream);
/* fp32 (N, C, HW) -> (N, HW, C) as fp32 (out_bf16 = 0) or bf16 (1) */
int dhd_nchw_to_nhwc(const float* in, int N, int C, int HW, void* out, int out_bf16, void* stream);

/* ---- streaming layout / elementwise helpers of the dense path (csrc/layout.cu) -----------
 * "split-bf16 NHWC": bf16, `ld` channels per pixel, logical channel c of part p at
 * coff + p*part_stride + c; the fp32 value is the sum of the parts. */
/* fp32 NCHW -> split-bf16 NHWC (reference modules exchange fp32 NCHW tensors) */
int dhd_pack_nchw_to_nhwc(const float* in, int N, int C, int H, int W, void* out, int out_ld,
                          int out_coff, int part_stride, int parts, void* stream);
/* class map of predictor.get_occ (occ_head.py:141-153): out[v] = argmax_k logits[v][k] (uint8) */
int dhd_occ_argmax(const float* logits, long nvox, int ncls, uint8_t* out, void* stream);

/* ---- fused occupancy-head tail ------------------------------------------------------
 * predictor.predicter = Linear(K1 -> N1) + Softplus + Linear(N1 -> Dz * n_cls) on every BEV pixel
 * (models/dense_heads/occ_head.py:63-67), the permute(0, 3, 2, 1) of predictor.forward (occ_head.py:84-100)
 * and, when `occ` is given, predictor.get_occ's softmax(-1).argmax(-1) -> uint8 (occ_head.py:141-153), as ONE
 * back-to-back tcgen05 GEMM kernel: the hidden layer stays in TMEM / shared memory, the logits leave the SM only
 * when `logits` is given.  bf16 operands (part 0 of the input activation, bf16 weights), fp32 accumulation.
 * in: bf16 NHWC rows [N*H*W][in_ld], channels [in_coff, in_coff + K1) = ReLU(final_conv(x)).
 * w1: bf16 [N1][K1]; w2: bf16 [Dz*n_cls][N1] (row-major nn.Linear weights); b1 / b2 fp32 or NULL.
 * logits: fp32 [N][W][H][Dz*n_cls] when transpose_xy (the reference's (B, Dx, Dy, Dz, n_cls)), else [N][H][W][..];
 * occ: uint8, same pixel order, [Dz] per pixel.  This build: K1 = 256, N1 = 512, Dz = 16, n_cls = 18. */
typedef struct dhd_predictor_tail_desc {
  int32_t N, H, W;
  int32_t K1, N1, Dz, n_cls;
  int32_t in_ld, in_coff;
  int32_t transpose_xy;
  const void* in;
  const void* w1;
  const float* b1;
  const void* w2;
  const float* b2;
  float* logits;
  uint8_t* occ;
  /* training: the Softplus output (bf16 [N*H*W][hidden_ld], channel c at hidden_coff + c, input pixel order) is also
   * written, for the backward of the two Linear layers; NULL = the hidden layer never leaves the SM */
  void* hidden;
  int32_t hidden_ld, hidden_coff;
} dhd_predictor_tail_desc;
int dhd_predictor_tail(const dhd_predictor_tail_desc* desc, void* stream);
/* encoder helpers, bf16 NHWC -> bf16 NHWC (the output may be a channel slice of a concatenation buffer):
 * MaxPool2d(2) (backbones/unet.py:65-74) and bilinear Upsample(align_corners=True) (necks/lss_fpn.py:27-28, 41-42) */
int dhd_maxpool2(const void* in, int in_ld, int in_coff, int in_part_stride, int N, int H, int W, int C,
                 void* out, int out_ld, int out_coff, int out_part_stride, int parts, void* stream);
int dhd_upsample_bilinear(const void* in, int in_ld, int in_coff, int in_part_stride, int N, int H, int W,
                          int C, int out_H, int out_W, void* out, int out_ld, int out_coff,
                          int out_part_stride, int parts, void* stream);
/* image backbone helpers (the §8(f)-4 widening: mmdet ResNet-50 / -101 `img_backbone` + necks/fpn.py CustomFPN):
 * stem im2col -- conv1 (7x7, stride 2, pad 3, 3-channel fp32 NCHW images) becomes a 1x1 tcgen05 GEMM over rows of
 * K = ksize*ksize*Cin (<= 512) values ordered (ky, kx, c), zero-padded to out_part_stride (split-bf16 parts at
 * out_part_stride; the padding is written by the kernel);  MaxPool2d(3, 2, 1);  io += F.interpolate(lo, size=(H, W), 'nearest')
 * (fpn.py:166-176) */
int dhd_stem_im2col(const float* img, int N, int Cin, int H, int W, int ksize, int stride, int pad, void* out, int out_ld,
                    int out_part_stride, int parts, void* stream);
int dhd_maxpool3s2(const void* in, int in_ld, int in_coff, int in_part_stride, int N, int H, int W, int C, void* out,
                   int out_ld, int out_coff, int out_part_stride, int parts, void* stream);
int dhd_upsample_nearest_add(const void* lo, int lo_ld, int lo_coff, int lo_part_stride, int h, int w, void* io, int io_ld,
                             int io_coff, int io_part_stride, int N, int H, int W, int C, int parts, void* stream);
/* backward of the encoder helpers: MaxPool2d(2) (gradient to the first maximum of each window, as torch) and
 * bilinear Upsample(align_corners=True) (dx fp32 [N*H*W][C], zeroed here, accumulated with atomics) */
int dhd_maxpool2_bwd(const void* x, int x_ld, int x_coff, const void* dy, int dy_ld, int dy_coff, int N, int H, int W,
                     int C, void* dx, int dx_ld, int dx_coff, void* stream);
int dhd_upsample_bilinear_bwd(const void* dy, int dy_ld, int dy_coff, int N, int H, int W, int C, int out_H, int out_W,
                              float* dx, void* stream);
/* the same backward in gather form (one thread per input pixel, fixed summation order): deterministic, no atomics, no
 * zero-fill; dx written once as bf16 rows (dx_b16, dx_ld, dx_coff) and / or fp32 [N*H*W][C] (dx_f32) */
int dhd_upsample_bilinear_bwd_gather(const void* dy, int dy_ld, int dy_coff, int N, int H, int W, int C, int out_H,
                                     int out_W, void* dx_b16, int dx_ld, int dx_coff, float* dx_f32, void* stream);
/* MaxPool2d(kernel 3, stride 2, padding 1) backward (mmdet ResNet.maxpool when the image backbone trains, DHD-S.py:44-55
 * `norm_eval=False, frozen_stages=-1`): x (N, H, W, C) is the pool's bf16 input, dy (N, oH, oW, C) the gradient of its
 * output, dx (N, H, W, C) is written completely.  The gradient of a window goes to its first maximum in scan order
 * (torch's rule), gathered per input pixel: deterministic.  relu_mask != 0: dx is also multiplied by [x > 0] (x is the
 * output of the stem's ReLU). */
int dhd_maxpool3s2_bwd(const void* x, int x_ld, int x_coff, const void* dy, int dy_ld, int dy_coff, int N, int H, int W,
                       int C, void* dx, int dx_ld, int dx_coff, int relu_mask, void* stream);
/* number of kernels this library has enqueued since it was loaded (bench bookkeeping) */
long dhd_launch_count(void);
/* fp32 rows [rows][C] (NHWC) -> split-bf16 rows */
int dhd_split_nhwc(const float* in, long rows, int C, void* out, int out_ld, int out_coff,
                   int part_stride, int parts, void* stream);
/* same for an (N, HW, C) tensor, optionally also writing the per-image channel means
 * (mean_out [N][C] or NULL): the SFA squeeze (mix.py:41) fused into the one read of the tensor */
int dhd_split_nhwc_mean(const float* in, int N, int HW, int C, void* out, int out_ld, int out_coff,
                        int part_stride, int parts, float* mean_out, float* workspace, void* stream);
int dhd_unpack_nhwc_to_nchw(const void* in, int in_ld, int in_coff, int part_stride, int parts,
                            int N, int C, int H, int W, float* out, void* stream);
/* out[n][c] = mean over the HW pixels (AdaptiveAvgPool2d(1): depthnet.py:77-82; mix.py:41) */
int dhd_mean_hw(const void* in, int in_ld, int in_coff, int part_stride, int parts, int N, int C,
                int HW, float* out, float* workspace, void* stream);
/* bytes of the caller-owned scratch the two mean producers need (block partial sums, reduced in fixed order:
 * the means are bit-identical from run to run) */
size_t dhd_mean_workspace_bytes(int N, int C, int HW);
/* y[r][o] = act(sum_k (x[r][k]*in_scale[k]+in_shift[k]) * w[o][k] + b[o]); one_minus: y = 1 - y.
 * fp32 CUDA-core path for the M = B*N row MLP / SE / fc chains (depthnet.py:119-169, mix.py:20-25) */
int dhd_linear_rows(const float* x, int R, int K, const float* w, const float* b, int O, int act,
                    const float* in_scale, const float* in_shift, int one_minus, float* y,
                    void* stream);
/* out = x * gate[n][c]: SELayer's multiply (depthnet.py:169) for a feature map that feeds two
 * differently gated branches (DepthNet.forward, depthnet.py:388-396); x fp32 NHWC [N][HW][C];
 * out split-bf16 NHWC, out32 optional fp32 NHWC copy (residual input of the next block) */
int dhd_gate_channels(const float* x, int N, int HW, int C, const float* gate, void* out, int o_ld,
                      int o_coff, int o_part_stride, int o_parts, float* out32, void* stream);
/* channel_spatial_stage blends (mix.py:44-57): x holds bev channels [0,C) and voxel channels
 * [C,2C); a1 [N][C]; a2 NULL -> a1*bev + (1-a1)*vox, else a2*(a1*bev) + (1-a2)*((1-a1)*vox) with
 * a2 fp32 NHWC [pix][C] (already sigmoid-ed) */
int dhd_sfa_mix(const void* x, int x_ld, int x_coff, int x_part_stride, int x_parts, int C, int N,
                int HW, const float* a1, const float* a2, void* out, int o_ld, int o_coff,
                int o_part_stride, int o_parts, void* stream);
/* bf16 speed mode of the second blend: a2 is a single-part bf16 NHWC activation (channel c at a2_coff + c of a2_ld),
 * x and out single-part bf16; same arithmetic and operation order as dhd_sfa_mix */
int dhd_sfa_blend_b16(const void* x, int x_ld, int x_coff, int C, int N, int HW, const float* a1, const void* a2,
                      int a2_ld, int a2_coff, void* out, int o_ld, int o_coff, void* stream);
/* The first blend folded into the weights of the convolution that consumes it (mix.py:41-50, 28-29): for a 1x1
 * weight w [Cout][C] fp32 and the channel gate a1 [N][C], writes out [N][Cout][2C] bf16 with
 * out[n][co][c] = w[co][c]*a1[n][c], out[n][co][C+c] = w[co][c]*(1-a1[n][c]) -- the per-image weights
 * (dhd_conv_desc.w_image_rows = Cout) of a 1x1 convolution over x = [bev | vox] that equals conv(w, a1*bev+(1-a1)*vox) */
int dhd_sfa_fold_gate(const float* w, const float* a1, int N, int Cout, int C, void* out, void* stream);
/* deformable bilinear im2col of mmcv DeformConv2dPack (depthnet.py:225-236, 466-477), stride 1,
 * deform_groups 1; offset fp32 [pix][off_ld] with (dy, dx) per tap; out channels ordered
 * [group][tap][C/groups] */
int dhd_dcn_im2col(const void* x, int x_ld, int x_coff, int x_part_stride, int x_parts, int C, int N,
                   int H, int W, const float* offset, int off_ld, int ksize, int pad, int dilation,
                   int groups, void* out, int o_ld, int o_part_stride, int o_parts, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DHD_B200_H_ */
