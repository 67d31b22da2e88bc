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
  DHD_LAYOUT_NCDHW = 2          /* (b, c, z, y, x): contiguous (B, C, dz, Dy, Dx), collapse_z=False */
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

#ifdef __cplusplus
}
#endif
#endif /* DHD_B200_H_ */
