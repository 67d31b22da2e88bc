// Streaming kernels of the training path of the dense hot-path layers (everything between the
// tensor-core GEMMs of the backward pass): activation derivatives with the per-channel bias /
// BatchNorm reductions fused, the class-weighted masked cross-entropy of the occupancy head
// (reference: models/dense_heads/occ_head.py:102-139 with mmdet's CrossEntropyLoss,
// models/losses/cross_entropy_loss.py:11-62), and the softmax backward of MGHS.depth_net
// (models/necks/lss_heightmap.py:482-489).  All HBM-bound: every tensor is read once and
// written once, 16 bytes per thread.
#include <cuda_bf16.h>

#include "common.cuh"

namespace dhd {

constexpr int kMaxSumBlocks = 592;    // 4 resident blocks (64 registers, 17 KB of shared memory each) x 148 SMs: one full wave

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 q = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v[2 * j] = __low2float(h[j]);
    v[2 * j + 1] = __high2float(h[j]);
  }
}
__device__ __forceinline__ void unpack8(const uint4& q, float (&v)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    v[2 * j] = __low2float(h[j]);
    v[2 * j + 1] = __high2float(h[j]);
  }
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&v)[8]) {
  uint4 q;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
  for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
  *reinterpret_cast<uint4*>(p) = q;
}

// dz = dy * act'(y) on bf16 NHWC rows; per-block partial column sums of dz and dz*y.
// act: 0 none, 1 relu (y > 0), 2 sigmoid (y (1 - y)), 3 softplus (1 - exp(-y)).
// gate (optional, [N][C], rows_per_img rows per image): dz *= gate (a per-image channel gate that was
// applied AFTER the activation in the forward epilogue, e.g. the SE gate).
__global__ void __launch_bounds__(256, 4)
act_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int dy_ld, int dy_coff, const __nv_bfloat16* __restrict__ y,
               int y_ld, int y_coff, long rows, int C, int act, __nv_bfloat16* __restrict__ out, int out_ld,
               int out_coff, float* __restrict__ partial, long rows_per_block,
               const __nv_bfloat16* __restrict__ add, int add_ld, int add_coff) {
  __shared__ float red[256][17];
  const int cg = C / 8, rpb = 256 / cg;
  const int gi = threadIdx.x % cg, ri = threadIdx.x / cg;
  const long r0 = (long)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float s1[8], s2[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
  if (ri < rpb) {
    // two rows per trip: every load of both rows is issued before the first use (the kernel is a pure HBM stream; with
    // one row per trip a thread had 32-48 bytes in flight and the launch reached 0.66 of the copy bandwidth)
    const bool need_y = act != 0 || partial != nullptr;
    for (long r = r0 + ri; r < r1; r += 2 * rpb) {
      const long rb = r + rpb;
      const bool two = rb < r1;
      uint4 qg[2], qa[2], qv[2];
      qg[0] = ld_nc_u4(dy + r * dy_ld + dy_coff + gi * 8);
      if (two) qg[1] = ld_nc_u4(dy + rb * dy_ld + dy_coff + gi * 8);
      if (add != nullptr) {            // a second gradient path into the same tensor (residual connection)
        qa[0] = ld_nc_u4(add + r * add_ld + add_coff + gi * 8);
        if (two) qa[1] = ld_nc_u4(add + rb * add_ld + add_coff + gi * 8);
      }
      if (need_y) {
        qv[0] = ld_nc_u4(y + r * y_ld + y_coff + gi * 8);
        if (two) qv[1] = ld_nc_u4(y + rb * y_ld + y_coff + gi * 8);
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        if (h == 1 && !two) break;
        float g[8], v[8];
        unpack8(qg[h], g);
        if (add != nullptr) {
          float a[8];
          unpack8(qa[h], a);
#pragma unroll
          for (int j = 0; j < 8; ++j) g[j] += a[j];
        }
        if (need_y) unpack8(qv[h], v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float d = 1.f;
          if (act == 1) d = v[j] > 0.f ? 1.f : 0.f;
          else if (act == 2) d = v[j] * (1.f - v[j]);
          else if (act == 3) d = 1.f - __expf(-v[j]);
          g[j] *= d;
          s1[j] += g[j];
          s2[j] += g[j] * v[j];
        }
        if (out != nullptr) store8(out + (h == 0 ? r : rb) * out_ld + out_coff + gi * 8, g);
      }
    }
  }
  if (partial == nullptr) return;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[threadIdx.x][j] = s1[j];
    red[threadIdx.x][8 + j] = s2[j];
  }
  __syncthreads();
  if (ri == 0) {
    for (int k = 1; k < rpb; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s1[j] += red[threadIdx.x + k * cg][j];
        s2[j] += red[threadIdx.x + k * cg][8 + j];
      }
    float* p = partial + (size_t)blockIdx.x * 2 * C;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      p[gi * 8 + j] = s1[j];
      p[C + gi * 8 + j] = s2[j];
    }
  }
}

// sums[i] = sum over blocks of partial[b][i]: one warp per column, lanes stride over the blocks, fixed
// shuffle tree (deterministic)
__global__ void __launch_bounds__(256)
colsum_reduce_kernel(const float* __restrict__ partial, int nblocks, int n, float* __restrict__ sums) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= n) return;
  float a = 0.f;
  for (int b = lane; b < nblocks; b += 32) a += partial[(size_t)b * n + i];
  a = warp_sum(a);
  if (lane == 0) sums[i] = a;
}

// sums[i] = sum over rows of partial[row][i], rows in a fixed order: a block owns 32 columns (lane = column: 128-byte
// coalesced row reads), its 32 warps stride over the rows with four independent loads in flight each, and the 32
// partial sums meet in shared memory
__global__ void __launch_bounds__(1024)
colsum_finish_kernel(const float* __restrict__ partial, int rows, int n, float* __restrict__ sums) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  if (i < n) {
    int r = w;
    for (; r + 96 < rows; r += 128) {
      a0 += __ldg(partial + (size_t)r * n + i);
      a1 += __ldg(partial + (size_t)(r + 32) * n + i);
      a2 += __ldg(partial + (size_t)(r + 64) * n + i);
      a3 += __ldg(partial + (size_t)(r + 96) * n + i);
    }
    for (; r < rows; r += 32) a0 += __ldg(partial + (size_t)r * n + i);
  }
  __shared__ float sh[32][33];
  sh[w][lane] = (a0 + a1) + (a2 + a3);
  __syncthreads();
  if (w == 0 && i < n) {
    float t = sh[0][lane];
#pragma unroll
    for (int k = 1; k < 32; ++k) t += sh[k][lane];
    sums[i] = t;
  }
}

// ---- occupancy cross-entropy ---------------------------------------------------------------
// norm[0] = sum over voxels of mask * class_weight[label]   (the reference's num_total_samples)
__global__ void __launch_bounds__(256)
ce_norm_kernel(const uint8_t* __restrict__ labels, const uint8_t* __restrict__ mask, const float* __restrict__ cw,
               int ncls, int ignore, long nvox, float* __restrict__ norm) {
  float a = 0.f;
  for (long v = (long)blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += (long)gridDim.x * blockDim.x) {
    const int l = labels[v];
    if (l == ignore || l >= ncls) continue;
    if (mask != nullptr && mask[v] == 0) continue;
    a += cw != nullptr ? cw[l] : 1.f;
  }
  a = warp_sum(a);
  __shared__ float ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += ws[i];
    atomicAdd(norm, t);
  }
}

// logits (B, Dx, Dy, Dz, ncls) fp32 (the predictor's output layout); labels / mask (B, Dx, Dy, Dz);
// dlogits bf16 NHWC rows [(b*Dy + y)*Dx + x][ld], channel z*ncls + k (the layout the last Linear's
// backward GEMMs read); loss[0] += sum of weighted voxel losses * loss_weight / norm.
// one voxel's class logits into registers (-inf beyond ncls): 8-byte loads when the row is 8-byte aligned (even ncls)
// NC = compile-time bound of the class loops (18 for the DHD heads: no work on the 14 padding slots of the generic
// 32-wide form; ncu showed both loss kernels issue-bound), ncls <= NC the run-time class count
template <int NC>
__device__ __forceinline__ void load_logits(const float* __restrict__ lg, int ncls, float (&e)[NC]) {
  if ((ncls & 1) == 0 && (NC & 1) == 0) {
#pragma unroll
    for (int k = 0; k < NC / 2; ++k) {
      if (2 * k < ncls) {
        const float2 t = __ldg(reinterpret_cast<const float2*>(lg) + k);
        e[2 * k] = t.x;
        e[2 * k + 1] = t.y;
      } else {
        e[2 * k] = e[2 * k + 1] = -INFINITY;
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < NC; ++k) e[k] = k < ncls ? __ldg(lg + k) : -INFINITY;
  }
}

template <int NC>
__global__ void __launch_bounds__(256)
ce_loss_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ labels, const uint8_t* __restrict__ mask,
               const float* __restrict__ cw, int ncls, int ignore, int B, int Dx, int Dy, int Dz, float loss_weight,
               const float* __restrict__ norm, float* __restrict__ loss, __nv_bfloat16* __restrict__ dlogits, int ld,
               const float* __restrict__ scal_gt, const float* __restrict__ scal_gn) {
  // scal_gt / scal_gn (optional, [ncls]): d(sem_scal + geo_scal)/d p_k of a masked voxel for k == label / k != label
  // (occ_scal_coeffs_kernel); chained through the softmax here so the logits are read once for all three terms
  __shared__ float s_gt[32], s_gn[32];
  if (threadIdx.x < 32) {
    s_gt[threadIdx.x] = scal_gt != nullptr && threadIdx.x < ncls ? scal_gt[threadIdx.x] : 0.f;
    s_gn[threadIdx.x] = scal_gn != nullptr && threadIdx.x < ncls ? scal_gn[threadIdx.x] : 0.f;
  }
  __syncthreads();
  const bool scal = scal_gt != nullptr;
  const long nvox = (long)B * Dx * Dy * Dz;
  const float inv = loss_weight / (norm[0] + 1.1920929e-07f);
  float acc = 0.f;
  // 32-bit index arithmetic (host check: nvox * ncls < 2^31 is NOT required, only nvox < 2^31): the four 64-bit
  // divisions per voxel were ~400 of this kernel's ~1150 instructions per warp trip
  for (unsigned vv = blockIdx.x * blockDim.x + threadIdx.x; vv < (unsigned)nvox; vv += gridDim.x * blockDim.x) {
    const long v = (long)vv;
    const unsigned t1 = vv / (unsigned)Dz;
    const int z = (int)(vv - t1 * (unsigned)Dz);
    const unsigned t2 = t1 / (unsigned)Dy;
    const int y = (int)(t1 - t2 * (unsigned)Dy);
    const int b = (int)(t2 / (unsigned)Dx);
    const int x = (int)(t2 - (unsigned)b * (unsigned)Dx);
    const float* lg = logits + v * ncls;
    const int l = labels[v];
    float w = 0.f;
    const bool in_mask = l != ignore && l < ncls && (mask == nullptr || mask[v] != 0);
    if (in_mask) w = cw != nullptr ? cw[l] : 1.f;
    __nv_bfloat16* g = dlogits + (((size_t)b * Dy + y) * Dx + x) * ld + z * ncls;
    const bool word_ok = ((z * ncls) & 1) == 0 && (ld & 1) == 0;        // 4-byte aligned pair stores
    if (!in_mask) {
      if (word_ok) {
        uint32_t* g2 = reinterpret_cast<uint32_t*>(g);
        for (int k = 0; k < ncls / 2; ++k) g2[k] = 0u;
        if (ncls & 1) g[ncls - 1] = __float2bfloat16(0.f);
      } else {
        for (int k = 0; k < ncls; ++k) g[k] = __float2bfloat16(0.f);
      }
      continue;
    }
    float e[NC];                     // ncls <= NC (host dispatch); logits stay in registers
    float mx = -INFINITY;
    load_logits(lg, ncls, e);
#pragma unroll
    for (int k = 0; k < NC; ++k) mx = fmaxf(mx, e[k]);
    float s = 0.f, ll = 0.f;
#pragma unroll
    for (int k = 0; k < NC; ++k) {
      if (k == l) ll = e[k];
      e[k] = __expf(e[k] - mx);      // exp(-inf) = 0 beyond ncls
      s += e[k];
    }
    acc += w * (mx + __logf(s) - ll);
    const float wi = w * inv / s, is = 1.f / s;
    float dot = 0.f;
    if (scal) {
#pragma unroll
      for (int k = 0; k < NC; ++k)
        if (k < ncls) dot += (k == l ? s_gt[k] : s_gn[k]) * e[k] * is;
    }
#pragma unroll
    for (int k = 0; k < NC; k += 2) {
      if (k < ncls) {
        float a0 = wi * e[k] - (k == l ? w * inv : 0.f);
        float a1 = k + 1 < NC ? wi * e[k + 1 < NC ? k + 1 : k] - (k + 1 == l ? w * inv : 0.f) : 0.f;
        if (scal) {
          a0 += e[k] * is * ((k == l ? s_gt[k] : s_gn[k]) - dot);
          if (k + 1 < ncls) a1 += e[k + 1 < NC ? k + 1 : k] * is * ((k + 1 == l ? s_gt[k + 1] : s_gn[k + 1]) - dot);
        }
        if (word_ok && k + 1 < ncls) {
          *reinterpret_cast<__nv_bfloat162*>(g + k) = __floats2bfloat162_rn(a0, a1);
        } else {
          g[k] = __float2bfloat16(a0);
          if (k + 1 < ncls) g[k + 1] = __float2bfloat16(a1);
        }
      }
    }
  }
  acc = warp_sum(acc);
  __shared__ float ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 8; ++i) t += ws[i];
    atomicAdd(loss, t * inv);
  }
}

// ---- MGHS.depth_net backward head ---------------------------------------------------------
// depth = softmax(logits) over D (NCHW, (BN, D, HW)); g = dL/d depth from the pool backward;
// feat_grad (BN*HW, C) = dL/d context.  Writes the gradient w.r.t. the 1x1 convolution's output as
// one bf16 NHWC row per pixel: [0, D) = depth * (g - <g, depth>), [D, D + C) = feat_grad, rest 0.
__global__ void __launch_bounds__(128)
depth_head_bwd_kernel(const float* __restrict__ depth, const float* __restrict__ g, const float* __restrict__ fg,
                      int BN, int D, int HW, int C, __nv_bfloat16* __restrict__ out, int ld) {
  const long pix = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (long)BN * HW) return;
  const int bn = (int)(pix / HW), hw = (int)(pix % HW);
  const float* dp = depth + (size_t)bn * D * HW + hw;
  const float* gp = g + (size_t)bn * D * HW + hw;
  float dot = 0.f;
  for (int d = 0; d < D; ++d) dot += dp[(size_t)d * HW] * gp[(size_t)d * HW];
  __nv_bfloat16* o = out + (size_t)pix * ld;
  for (int d = 0; d < D; ++d) o[d] = __float2bfloat16(dp[(size_t)d * HW] * (gp[(size_t)d * HW] - dot));
  const float* f = fg + (size_t)pix * C;
  for (int c = 0; c < C; ++c) o[D + c] = __float2bfloat16(f[c]);
  for (int c = D + C; c < ld; ++c) o[c] = __float2bfloat16(0.f);
}


// ---- SFA gates backward (mix.py:37-59) -----------------------------------------------------
// forward:  u = a1*bev + (1-a1)*vox ;  fuse = a2*a1*bev + (1-a2)*(1-a1)*vox
//           (a1 [N][C] per-image channel gate, a2 [pix][C] spatial gate, x = [bev | vox] bf16 NHWC)
// mode 0 (fuse):  g = d fuse
//    dpre2        = g * (a1*bev - (1-a1)*vox) * a2*(1-a2)          (bf16, gradient at the sigmoid input)
//    dx[bev]      = g * a2*a1          dx[vox] = g * (1-a2)*(1-a1)  (fp32, written)
//    s[n][c]     += g * (a2*bev - (1-a2)*vox)                       (d a1, per-image partial sums)
// mode 1 (u):     g = d u
//    dx[bev]     += g * a1             dx[vox] += g * (1-a1)        (fp32, accumulated)
//    s[n][c]     += g * (bev - vox)
// Blocks cover `pb` pixels of one image; their partial sums go to partial[(n*gridDim.x + bx)][C].
// DXB16 (the bf16 form used by the training step: 2C fp32 channels per pixel written, re-read by the shortcut's
// data-gradient epilogue, read-modified-written by mode 1 and read once more by the final row-vector add were ~2 GB of
// traffic per step): mode 0 writes dx as a bf16 activation (dxh: [pix][dxh_ld], bev at dxh_coff, vox at dxh_coff + C),
// mode 1 only reduces s -- its dx terms a1*g / (1-a1)*g need no x and are added by sfa_dx_combine_kernel.
template <bool DXB16>
__global__ void __launch_bounds__(256)
sfa_gate_bwd_kernel(int mode, const __nv_bfloat16* __restrict__ g, int g_ld, int g_coff, const __nv_bfloat16* __restrict__ x,
                    int x_ld, int x_coff, int C, int HW, int pb, const float* __restrict__ a1,
                    const float* __restrict__ a2, __nv_bfloat16* __restrict__ dpre2, int d_ld, int d_coff,
                    float* __restrict__ dx, float* __restrict__ partial, __nv_bfloat16* __restrict__ dxh, int dxh_ld,
                    int dxh_coff) {
  __shared__ float red[256][9];
  const int cg = C / 8, rows = 256 / cg;
  const int gi = threadIdx.x % cg, ri = threadIdx.x / cg;
  const int n = blockIdx.y, p0 = blockIdx.x * pb, p1 = min(HW, p0 + pb);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (ri < rows) {
    const float4* g1p = reinterpret_cast<const float4*>(a1 + (size_t)n * C + gi * 8);
    const float4 ga = __ldg(g1p), gb = __ldg(g1p + 1);
    const float g1[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
    for (int p = p0 + ri; p < p1; p += rows) {
      const size_t r = (size_t)n * HW + p;
      float gv[8], bev[8], vox[8];
      load8(g + r * g_ld + g_coff + gi * 8, gv);
      load8(x + r * x_ld + x_coff + gi * 8, bev);
      load8(x + r * x_ld + x_coff + C + gi * 8, vox);
      float* db = DXB16 ? nullptr : dx + r * (size_t)(2 * C) + gi * 8;
      float* dv = DXB16 ? nullptr : db + C;
      float ob[8], ov[8];
      if (mode == 0) {
        const float4* g2p = reinterpret_cast<const float4*>(a2 + r * C + gi * 8);
        const float4 qa = __ldg(g2p), qb = __ldg(g2p + 1);
        const float g2[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};
        float dp[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float b1 = g1[j] * bev[j], v1 = (1.f - g1[j]) * vox[j];
          dp[j] = gv[j] * (b1 - v1) * g2[j] * (1.f - g2[j]);
          ob[j] = gv[j] * g2[j] * g1[j];
          ov[j] = gv[j] * (1.f - g2[j]) * (1.f - g1[j]);
          acc[j] += gv[j] * (g2[j] * bev[j] - (1.f - g2[j]) * vox[j]);
        }
        store8(dpre2 + r * d_ld + d_coff + gi * 8, dp);
      } else if (DXB16) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += gv[j] * (bev[j] - vox[j]);
        continue;
      } else {
        const float4 b0 = *reinterpret_cast<const float4*>(db), b1 = *reinterpret_cast<const float4*>(db + 4);
        const float4 v0 = *reinterpret_cast<const float4*>(dv), v1 = *reinterpret_cast<const float4*>(dv + 4);
        const float pbv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        const float pvv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          ob[j] = pbv[j] + gv[j] * g1[j];
          ov[j] = pvv[j] + gv[j] * (1.f - g1[j]);
          acc[j] += gv[j] * (bev[j] - vox[j]);
        }
      }
      if (DXB16) {
        store8(dxh + r * dxh_ld + dxh_coff + gi * 8, ob);
        store8(dxh + r * dxh_ld + dxh_coff + C + gi * 8, ov);
        continue;
      }
      *reinterpret_cast<float4*>(db) = make_float4(ob[0], ob[1], ob[2], ob[3]);
      *reinterpret_cast<float4*>(db + 4) = make_float4(ob[4], ob[5], ob[6], ob[7]);
      *reinterpret_cast<float4*>(dv) = make_float4(ov[0], ov[1], ov[2], ov[3]);
      *reinterpret_cast<float4*>(dv + 4) = make_float4(ov[4], ov[5], ov[6], ov[7]);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.x][j] = acc[j];
  __syncthreads();
  if (ri == 0) {
    for (int k = 1; k < rows; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += red[threadIdx.x + k * cg][j];
    float* pp = partial + ((size_t)n * gridDim.x + blockIdx.x) * C + gi * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) pp[j] = acc[j];
  }
}

// dx = dxb + [a1 * du | (1 - a1) * du] + ds[n]  (the channel-gate blend's data gradient and the squeeze path's per-image
// row vector joined with what the fuse blend and the shortcut left in dxb); 8 channels of both halves per thread
__global__ void __launch_bounds__(256)
sfa_dx_combine_kernel(const __nv_bfloat16* __restrict__ dxb, int b_ld, int b_coff, const __nv_bfloat16* __restrict__ du,
                      int u_ld, int u_coff, const float* __restrict__ a1, const float* __restrict__ ds, int C, long npix,
                      int HW, __nv_bfloat16* __restrict__ out, int o_ld, int o_coff) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int cg = C / 8;
  if (i >= npix * cg) return;
  int c;
  const long pix = fast_div(i, cg, &c);
  c *= 8;
  const int n = (int)fast_div(pix, HW);
  float b[8], v[8], g[8];
  unpack8(ld_nc_u4(dxb + pix * b_ld + b_coff + c), b);
  unpack8(ld_nc_u4(dxb + pix * b_ld + b_coff + C + c), v);
  unpack8(ld_nc_u4(du + pix * u_ld + u_coff + c), g);
  const float4* gp = reinterpret_cast<const float4*>(a1 + (size_t)n * C + c);
  const float4 ga = __ldg(gp), gb = __ldg(gp + 1);
  const float g1[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    b[j] += g[j] * g1[j];
    v[j] += g[j] * (1.f - g1[j]);
  }
  if (ds != nullptr) {
    const float* dp = ds + (size_t)n * 2 * C + c;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      b[j] += __ldg(dp + j);
      v[j] += __ldg(dp + C + j);
    }
  }
  store8(out + pix * o_ld + o_coff + c, b);
  store8(out + pix * o_ld + o_coff + C + c, v);
}

// sums[n][c] (+)= sum over the image's blocks (ascending)
__global__ void __launch_bounds__(256)
image_sum_reduce_kernel(const float* __restrict__ partial, int nb, int N, int C, float* __restrict__ sums, int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * C) return;
  const int n = i / C, c = i % C;
  float a = 0.f;
  for (int b = 0; b < nb; ++b) a += partial[((size_t)n * nb + b) * C + c];
  sums[i] = accumulate ? sums[i] + a : a;
}

// out[pix][c] (bf16) = in[pix][c] (fp32) + v[n][c]
__global__ void __launch_bounds__(256)
add_rowvec_kernel(const float* __restrict__ in, const float* __restrict__ v, long npix, int HW, int C,
                  __nv_bfloat16* __restrict__ out, int o_ld, int o_coff) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int cg = C / 8;
  if (i >= npix * cg) return;
  int c;
  const long pix = fast_div(i, cg, &c);
  c *= 8;
  const int n = (int)fast_div(pix, HW);
  const float4 a = *reinterpret_cast<const float4*>(in + pix * C + c), b = *reinterpret_cast<const float4*>(in + pix * C + c + 4);
  float r[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  if (v != nullptr) {
    const float4 p = __ldg(reinterpret_cast<const float4*>(v + (size_t)n * C + c)), q = __ldg(reinterpret_cast<const float4*>(v + (size_t)n * C + c) + 1);
    r[0] += p.x; r[1] += p.y; r[2] += p.z; r[3] += p.w; r[4] += q.x; r[5] += q.y; r[6] += q.z; r[7] += q.w;
  }
  store8(out + pix * o_ld + o_coff + c, r);
}


// ---- SE gate backward (depthnet.py:150-169, 624-629): h = relu(bn(conv x)) * gate[n][c] ----------
// dpre = dh * gate * (h > 0) (bf16);  partial per-image sums of dh * h / gate = d gate.
__global__ void __launch_bounds__(256)
se_gate_bwd_kernel(const __nv_bfloat16* __restrict__ dh, int g_ld, int g_coff, const __nv_bfloat16* __restrict__ h,
                   int h_ld, int h_coff, int C, int HW, int pb, const float* __restrict__ gate,
                   __nv_bfloat16* __restrict__ dpre, int d_ld, int d_coff, float* __restrict__ partial) {
  __shared__ float red[256][9];
  const int cg = C / 8, rows = 256 / cg;
  const int gi = threadIdx.x % cg, ri = threadIdx.x / cg;
  const int n = blockIdx.y, p0 = blockIdx.x * pb, p1 = min(HW, p0 + pb);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (ri < rows) {
    const float4* gp = reinterpret_cast<const float4*>(gate + (size_t)n * C + gi * 8);
    const float4 ga = __ldg(gp), gb = __ldg(gp + 1);
    const float g1[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
    for (int p = p0 + ri; p < p1; p += rows) {
      const size_t r = (size_t)n * HW + p;
      float gv[8], hv[8], o[8];
      load8(dh + r * g_ld + g_coff + gi * 8, gv);
      load8(h + r * h_ld + h_coff + gi * 8, hv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o[j] = hv[j] > 0.f ? gv[j] * g1[j] : 0.f;
        acc[j] += g1[j] != 0.f ? gv[j] * hv[j] / g1[j] : 0.f;
      }
      store8(dpre + r * d_ld + d_coff + gi * 8, o);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.x][j] = acc[j];
  __syncthreads();
  if (ri == 0) {
    for (int k = 1; k < rows; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += red[threadIdx.x + k * cg][j];
    float* pp = partial + ((size_t)n * gridDim.x + blockIdx.x) * C + gi * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) pp[j] = acc[j];
  }
}

// ---- height loss (lss_heightmap.py:595-622) ------------------------------------------------------
// height = softmax over H bins (NCHW (BN, H, HW)); label[pix] = GT bin (or -1: all-zero one-hot row);
// fg[pix] = pixel has a valid GT depth.  loss = weight * sum_fg BCE(height, onehot) / max(1, #fg);
// dz (bf16 NHWC, ld channels) = d loss / d logits through the softmax; zero rows for background pixels.
__global__ void __launch_bounds__(128)
height_loss_kernel(const float* __restrict__ height, const int* __restrict__ label, const uint8_t* __restrict__ fg,
                   int BN, int H, int HW, float weight, const float* __restrict__ nfg, float* __restrict__ loss,
                   __nv_bfloat16* __restrict__ dz, int ld) {
  const long pix = (long)blockIdx.x * blockDim.x + threadIdx.x;
  float acc = 0.f;
  if (pix < (long)BN * HW) {
    const int bn = (int)(pix / HW), hw = (int)(pix % HW);
    __nv_bfloat16* o = dz + (size_t)pix * ld;
    if (fg[pix] == 0) {
      for (int k = 0; k < ld; ++k) o[k] = __float2bfloat16(0.f);
    } else {
      const float scale = weight / fmaxf(1.f, nfg[0]);
      const float* hp = height + (size_t)bn * H * HW + hw;
      const int l = label[pix];
      float dot = 0.f;
      for (int k = 0; k < H; ++k) {
        const float p = hp[(size_t)k * HW], t = k == l ? 1.f : 0.f;
        acc -= t * fmaxf(__logf(p), -100.f) + (1.f - t) * fmaxf(__logf(1.f - p), -100.f);
        const float g = (p - t) / fmaxf(p * (1.f - p), 1e-12f);
        dot += g * p;
      }
      for (int k = 0; k < H; ++k) {
        const float p = hp[(size_t)k * HW], t = k == l ? 1.f : 0.f;
        const float g = (p - t) / fmaxf(p * (1.f - p), 1e-12f);
        o[k] = __float2bfloat16(scale * p * (g - dot));
      }
      for (int k = H; k < ld; ++k) o[k] = __float2bfloat16(0.f);
      acc *= scale;
    }
  }
  acc = warp_sum(acc);
  __shared__ float ws[4];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) atomicAdd(loss, ws[0] + ws[1] + ws[2] + ws[3]);
}

// ---- deformable convolution backward, sampling part (mmcv DeformConv2dPack / modulated-free DCNv1) ----
// forward: col[pix][g][tap][c] = bilinear(x, (y + tap_y + off_y, x + tap_x + off_x)) (dcn_im2col in layout.cu)
// backward, one warp per (pixel, tap):  dx[corner][c] += w_corner * dcol[c]  (fp32 atomics),
//   doff[pix][2 tap + 0/1] = sum_c dcol[c] * d sample / d (sy, sx).
__global__ void __launch_bounds__(256)
dcn_col2im_bwd_kernel(const __nv_bfloat16* __restrict__ dcol, int c_ld, const __nv_bfloat16* __restrict__ x, int x_ld,
                      int x_coff, int C, int N, int H, int W, const float* __restrict__ offset, int off_ld, int ksize,
                      int pad, int dil, int groups, float* __restrict__ dx, float* __restrict__ doff) {
  const int lane = threadIdx.x & 31;
  const int gw = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);     // N*H*W*taps < 2^31 (host check):
  const int taps = ksize * ksize;                                                // 32-bit index arithmetic throughout
  if (gw >= N * H * W * taps) return;
  const int t = gw % taps;
  const int pix = gw / taps;
  const int wx = pix % W, hy = (pix / W) % H, n = pix / (W * H);
  const float oy = __ldg(offset + (size_t)pix * off_ld + 2 * t), ox = __ldg(offset + (size_t)pix * off_ld + 2 * t + 1);
  const float sy = (float)(hy - pad + (t / ksize) * dil) + oy;
  const float sx = (float)(wx - pad + (t % ksize) * dil) + ox;
  const int cg = C / groups;
  const bool inside = sy > -1.f && sx > -1.f && sy < (float)H && sx < (float)W;
  float gy = 0.f, gx = 0.f;
  if (inside) {
    const int y0 = (int)floorf(sy), x0 = (int)floorf(sx), y1 = y0 + 1, x1 = x0 + 1;
    const float ly = sy - (float)y0, lx = sx - (float)x0, hy_ = 1.f - ly, hx_ = 1.f - lx;
    const bool v00 = y0 >= 0 && x0 >= 0, v01 = y0 >= 0 && x1 <= W - 1, v10 = y1 <= H - 1 && x0 >= 0,
               v11 = y1 <= H - 1 && x1 <= W - 1;
    const size_t img = (size_t)n * H * W;
    for (int c = 8 * lane; c < C; c += 256) {
      const int g = c / cg, cl = c % cg;
      float d[8], q00[8], q01[8], q10[8], q11[8];
      load8(dcol + (size_t)pix * c_ld + (size_t)g * taps * cg + (size_t)t * cg + cl, d);
#pragma unroll
      for (int j = 0; j < 8; ++j) q00[j] = q01[j] = q10[j] = q11[j] = 0.f;
      const __nv_bfloat16* xb = x + img * x_ld + x_coff + c;
      float* db = dx + img * C + c;
      if (v00) {
        load8(xb + ((size_t)y0 * W + x0) * x_ld, q00);
        {
          float4* q4 = reinterpret_cast<float4*>(db + ((size_t)y0 * W + x0) * C);
          const float ww = hy_ * hx_;
          atomicAdd(q4, make_float4(ww * d[0], ww * d[1], ww * d[2], ww * d[3]));
          atomicAdd(q4 + 1, make_float4(ww * d[4], ww * d[5], ww * d[6], ww * d[7]));
        }
      }
      if (v01) {
        load8(xb + ((size_t)y0 * W + x1) * x_ld, q01);
        {
          float4* q4 = reinterpret_cast<float4*>(db + ((size_t)y0 * W + x1) * C);
          const float ww = hy_ * lx;
          atomicAdd(q4, make_float4(ww * d[0], ww * d[1], ww * d[2], ww * d[3]));
          atomicAdd(q4 + 1, make_float4(ww * d[4], ww * d[5], ww * d[6], ww * d[7]));
        }
      }
      if (v10) {
        load8(xb + ((size_t)y1 * W + x0) * x_ld, q10);
        {
          float4* q4 = reinterpret_cast<float4*>(db + ((size_t)y1 * W + x0) * C);
          const float ww = ly * hx_;
          atomicAdd(q4, make_float4(ww * d[0], ww * d[1], ww * d[2], ww * d[3]));
          atomicAdd(q4 + 1, make_float4(ww * d[4], ww * d[5], ww * d[6], ww * d[7]));
        }
      }
      if (v11) {
        load8(xb + ((size_t)y1 * W + x1) * x_ld, q11);
        {
          float4* q4 = reinterpret_cast<float4*>(db + ((size_t)y1 * W + x1) * C);
          const float ww = ly * lx;
          atomicAdd(q4, make_float4(ww * d[0], ww * d[1], ww * d[2], ww * d[3]));
          atomicAdd(q4 + 1, make_float4(ww * d[4], ww * d[5], ww * d[6], ww * d[7]));
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        gy += d[j] * ((q10[j] - q00[j]) * hx_ + (q11[j] - q01[j]) * lx);
        gx += d[j] * ((q01[j] - q00[j]) * hy_ + (q11[j] - q10[j]) * ly);
      }
    }
  }
  gy = warp_sum(gy);
  gx = warp_sum(gx);
  if (lane == 0) {
    doff[(size_t)pix * off_ld + 2 * t] = gy;
    doff[(size_t)pix * off_ld + 2 * t + 1] = gx;
  }
}


// ---- weight re-pack after an optimizer step ------------------------------------------------------
// fp32 master weight w[co][ci_total][tap] (nn.Conv2d / nn.Linear layout) ->
//   fwd  bf16 [Cout][taps][cin_pad]            : w[co][col_lo + ci][t]                 (forward GEMM operand)
//   bwd  bf16 [Cin][taps][cout_pad]  (mode 0)  : scale[co] * w[co][col_lo + ci][taps-1-t]  (data-gradient operand)
//        bf16 [taps][Cin][cout_pad]  (mode 1)  : scale[co] * w[co][col_lo + ci][t]      (grouped 1x1 view, K = (tap, ci))
// zero padded; one launch per layer instead of a dozen small tensor ops.
__global__ void __launch_bounds__(256)
pack_conv_weights_kernel(const float* __restrict__ w, int Cout, int cin_total, int taps, int col_lo, int Cin,
                         const float* __restrict__ scale, __nv_bfloat16* __restrict__ fwd, int cin_pad,
                         __nv_bfloat16* __restrict__ bwd, int cout_pad, int bwd_mode) {
  const long nf = fwd != nullptr ? (long)Cout * taps * cin_pad : 0;
  const long nb = bwd != nullptr ? (long)Cin * taps * cout_pad : 0;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < nf + nb; i += (long)gridDim.x * blockDim.x) {
    if (i < nf) {
      const int ci = (int)(i % cin_pad);
      const int t = (int)((i / cin_pad) % taps);
      const int co = (int)(i / ((long)cin_pad * taps));
      const float v = ci < Cin ? w[((size_t)co * cin_total + col_lo + ci) * taps + t] : 0.f;
      fwd[i] = __float2bfloat16(v);
    } else {
      const long j = i - nf;
      const int co = (int)(j % cout_pad);
      int ci, t;
      if (bwd_mode == 0) {
        t = (int)((j / cout_pad) % taps);
        ci = (int)(j / ((long)cout_pad * taps));
      } else {
        ci = (int)((j / cout_pad) % Cin);
        t = (int)(j / ((long)cout_pad * Cin));
      }
      float v = 0.f;
      if (co < Cout) {
        const int ts = bwd_mode == 0 ? taps - 1 - t : t;
        v = w[((size_t)co * cin_total + col_lo + ci) * taps + ts];
        if (scale != nullptr) v *= scale[co];
      }
      bwd[j] = __float2bfloat16(v);
    }
  }
}


struct PackBatch {
  dhd_pack_desc d[DHD_PACK_MAX_BATCH];
};
// blockIdx.y = layer, blockIdx.x = a (32 output channels x 32 input channels x all taps) tile of it.  The master
// weight w[co][ci][t] is read in contiguous runs of 32 * taps floats per output channel (coalesced), transposed through
// shared memory, and both operands leave in 64-byte rows: fwd[co][t][ci0..ci0+31] and bwd[ci][t'][co0..co0+31].  (One
// thread per OUTPUT element read the bwd operand's sources at a stride of cin_total * taps floats -- 8x the sectors:
// 807 us for the 122 M parameters of the widened step, ~160 us of bytes.)
constexpr int kPackTile = 32;
__global__ void __launch_bounds__(256)
pack_conv_weights_batch_kernel(const __grid_constant__ PackBatch B) {
  extern __shared__ float tile[];                    // [32 co][32 * taps + 1]
  const dhd_pack_desc& L = B.d[blockIdx.y];
  const int Cout = L.Cout, cin_total = L.cin_total, taps = L.taps, col_lo = L.col_lo, Cin = L.Cin, cin_pad = L.cin_pad,
            cout_pad = L.cout_pad, bwd_mode = L.bwd_mode;
  const int co_tiles = (max(Cout, L.bwd != nullptr ? cout_pad : Cout) + kPackTile - 1) / kPackTile;
  const int ci_tiles = (max(Cin, L.fwd != nullptr ? cin_pad : Cin) + kPackTile - 1) / kPackTile;
  if ((int)blockIdx.x >= co_tiles * ci_tiles) return;
  const int co0 = ((int)blockIdx.x / ci_tiles) * kPackTile, ci0 = ((int)blockIdx.x % ci_tiles) * kPackTile;
  const float* __restrict__ w = L.w;
  const int run = kPackTile * taps, pitch = run + 1;
  const int ci_n = max(0, min(kPackTile, Cin - ci0));            // input channels of this tile that exist
  for (int co_l = threadIdx.x >> 5; co_l < kPackTile; co_l += 8) {
    const int co = co0 + co_l;
    const float* src = w + ((size_t)co * cin_total + col_lo + ci0) * taps;
    for (int e = threadIdx.x & 31; e < run; e += 32)
      tile[co_l * pitch + e] = (co < Cout && e < ci_n * taps) ? __ldg(src + e) : 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (L.fwd != nullptr && ci0 + lane < cin_pad) {                  // fwd[co][t][ci]: lane = ci
    __nv_bfloat16* __restrict__ fwd = (__nv_bfloat16*)L.fwd;
    for (int r = wid; r < kPackTile * taps; r += 8) {
      const int co_l = r / taps, t = r - co_l * taps, co = co0 + co_l;
      if (co < Cout) fwd[((size_t)co * taps + t) * cin_pad + ci0 + lane] = __float2bfloat16(tile[co_l * pitch + lane * taps + t]);
    }
  }
  if (L.bwd != nullptr && co0 + lane < cout_pad) {                 // bwd[ci][t'][co] / [t][ci][co]: lane = co
    __nv_bfloat16* __restrict__ bwd = (__nv_bfloat16*)L.bwd;
    const int co = co0 + lane;
    const float sc = (L.scale != nullptr && co < Cout) ? __ldg(L.scale + co) : 1.f;
    for (int r = wid; r < ci_n * taps; r += 8) {
      const int ci_l = r / taps, t = r - ci_l * taps, ci = ci0 + ci_l;
      const float v = tile[lane * pitch + ci_l * taps + t] * sc;  // zero for co >= Cout (tile rows beyond Cout are zero)
      const size_t o = bwd_mode == 0 ? ((size_t)ci * taps + (taps - 1 - t)) * cout_pad + co
                                     : ((size_t)t * Cin + ci) * cout_pad + co;
      bwd[o] = __float2bfloat16(v);
    }
  }
}

// torch.optim.AdamW, single tensor form (torch/optim/adamw.py _single_tensor_adamw): 4 values per thread
__global__ void __launch_bounds__(256)
adamw_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                  long n, float lr, float beta1, float beta2, float eps, float wd, float bc1, float bc2_sqrt,
                  const float* __restrict__ grad_scale) {
  const float gs = grad_scale != nullptr ? __ldg(grad_scale) : 1.f;
  const float step_size = lr / bc1;
  const long i4 = ((long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i4 >= n) return;
  float pv[4], gv[4], mv[4], vv[4];
  const bool full = i4 + 4 <= n && ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0);
  if (full) {
    *reinterpret_cast<float4*>(pv) = *reinterpret_cast<const float4*>(p + i4);
    *reinterpret_cast<float4*>(gv) = *reinterpret_cast<const float4*>(g + i4);
    *reinterpret_cast<float4*>(mv) = *reinterpret_cast<const float4*>(m + i4);
    *reinterpret_cast<float4*>(vv) = *reinterpret_cast<const float4*>(v + i4);
  } else {
    for (int j = 0; j < 4; ++j)
      if (i4 + j < n) { pv[j] = p[i4 + j]; gv[j] = g[i4 + j]; mv[j] = m[i4 + j]; vv[j] = v[i4 + j]; }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float gr = gv[j] * gs;
    float pp = pv[j] * (1.f - lr * wd);
    mv[j] = mv[j] + (1.f - beta1) * (gr - mv[j]);             // lerp(m, g, 1 - beta1)
    vv[j] = beta2 * vv[j] + (1.f - beta2) * gr * gr;
    const float denom = sqrtf(vv[j]) / bc2_sqrt + eps;
    pp -= step_size * (mv[j] / denom);
    pv[j] = pp;
  }
  if (full) {
    *reinterpret_cast<float4*>(p + i4) = *reinterpret_cast<const float4*>(pv);
    *reinterpret_cast<float4*>(m + i4) = *reinterpret_cast<const float4*>(mv);
    *reinterpret_cast<float4*>(v + i4) = *reinterpret_cast<const float4*>(vv);
  } else {
    for (int j = 0; j < 4; ++j)
      if (i4 + j < n) { p[i4 + j] = pv[j]; m[i4 + j] = mv[j]; v[i4 + j] = vv[j]; }
  }
}

// ---- sem_scal / geo_scal statistics (semkitti_loss.py:136-225) ---------------------------------------------
// Over the masked voxels (label != ignore, mask set): per class i  Sp_i = sum p_i, Nom_i = sum p_i [t == i],
// Cnt_i = sum [t == i]; M = number of masked voxels.  Deterministic: warp shuffles in a fixed tree, per-block
// partials [nblocks][3*32 + 1] reduced in order by occ_scal_coeffs_kernel.
constexpr int kScalRow = 3 * 32 + 1;
template <int NC>
__global__ void __launch_bounds__(256)
occ_scal_stats_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ labels, const uint8_t* __restrict__ mask,
                      int ncls, int ignore, long nvox, float* __restrict__ partial) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  extern __shared__ float2 cls_acc[];                // [256 threads][33]: (Nom, Cnt) of class k for this thread's voxels
  for (int i = threadIdx.x; i < 256 * 33; i += 256) cls_acc[i] = make_float2(0.f, 0.f);
  __syncthreads();
  float sp[NC];
#pragma unroll
  for (int k = 0; k < NC; ++k) sp[k] = 0.f;
  float nom = 0.f, cnt = 0.f, m = 0.f;               // lane i of a warp carries class i's Nom / Cnt
  const long stride = (long)gridDim.x * blockDim.x;
  for (long v0 = (long)blockIdx.x * blockDim.x + wid * 32; v0 < nvox; v0 += stride) {
    const long v = v0 + lane;
    int l = -1;
    float pl = 0.f;
    if (v < nvox) {
      const int t = labels[v];
      if (t != ignore && t < ncls && (mask == nullptr || mask[v] != 0)) {
        l = t;
        const float* lg = logits + v * ncls;
        float e[NC], mx = -INFINITY, s = 0.f;
        load_logits(lg, ncls, e);
#pragma unroll
        for (int k = 0; k < NC; ++k) mx = fmaxf(mx, e[k]);
#pragma unroll
        for (int k = 0; k < NC; ++k) {
          e[k] = __expf(e[k] - mx);
          s += e[k];
        }
        const float is = 1.f / s;
#pragma unroll
        for (int k = 0; k < NC; ++k) {
          sp[k] += e[k] * is;
          if (k == l) pl = e[k] * is;
        }
        m += 1.f;
      }
    }
    // class l's Nom / Cnt: every thread keeps its own per-class pair in shared memory (no atomics: deterministic).
    // The first version reduced every class over the warp in every trip: 18 x 2 five-step shuffle trees = ~1000 of the
    // ~1150 instructions per trip, and ncu showed the kernel issue-bound (profiles/r02_loss_kernels.txt).
    if (l >= 0) {
      float2* mine = cls_acc + threadIdx.x * 33 + l;
      float2 t = *mine;
      t.x += pl;
      t.y += 1.f;
      *mine = t;
    }
  }
  __syncwarp();
  for (int i = 0; i < ncls; ++i) {                   // once per kernel: lane i <- class i over the warp's 32 threads, in order
    float a = 0.f, c = 0.f;
    if (lane == i) {
      for (int t = 0; t < 32; ++t) {
        const float2 q = cls_acc[(wid * 32 + t) * 33 + i];
        a += q.x;
        c += q.y;
      }
      nom = a;
      cnt = c;
    }
  }
  __shared__ float red[8][kScalRow];
#pragma unroll
  for (int k = 0; k < NC; ++k) {
    const float t = warp_sum(sp[k]);
    if (lane == 0) red[wid][k] = t;
  }
  if (lane == 0)
    for (int k = NC; k < 32; ++k) red[wid][k] = 0.f;
  red[wid][32 + lane] = nom;
  red[wid][64 + lane] = cnt;
  m = warp_sum(m);
  if (lane == 0) red[wid][96] = m;
  __syncthreads();
  for (int i = threadIdx.x; i < kScalRow; i += blockDim.x) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w][i];
    partial[(size_t)blockIdx.x * kScalRow + i] = t;
  }
}

__device__ __forceinline__ float scal_step(float x) {      // inverse_sigmoid's stepping (semkitti_loss.py:8-16)
  if (x >= 1.f - 1e-5f) x -= 1e-5f;
  if (x < 1e-5f) x += 1e-5f;
  return x;
}

// One warp: lane i = class i.  out[0] = weight_sem * sem_scal, out[1] = weight_geo * geo_scal;
// gt[i] / gn[i] = d(both terms)/d p_i of a masked voxel whose label is / is not i.
__global__ void __launch_bounds__(1024) occ_scal_coeffs_kernel(const float* __restrict__ partial, int nblocks, int ncls, int non_empty,
                                       float w_sem, float w_geo, float* __restrict__ out, float* __restrict__ gt,
                                       float* __restrict__ gn) {
  // every warp sums a strided share of the per-block partials (lane i = class i), warp 0 adds the warps' sums in
  // warp order: a fixed summation tree whatever the timing (deterministic), ~nblocks / 32 dependent steps instead of nblocks
  __shared__ float red[32][4][32];
  const int i = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float sp = 0.f, nom = 0.f, cnt = 0.f, M = 0.f;
  for (int b = wid; b < nblocks; b += nw) {
    const float* p = partial + (size_t)b * kScalRow;
    sp += p[i];
    nom += p[32 + i];
    cnt += p[64 + i];
    M += p[96];
  }
  red[wid][0][i] = sp;
  red[wid][1][i] = nom;
  red[wid][2][i] = cnt;
  red[wid][3][i] = M;
  __syncthreads();
  if (wid != 0) return;
  sp = nom = cnt = M = 0.f;
  for (int w = 0; w < nw; ++w) {
    sp += red[w][0][i];
    nom += red[w][1][i];
    cnt += red[w][2][i];
    M += red[w][3][i];
  }
  const float eps = 1e-5f;
  float loss = 0.f, a = 0.f, bb = 0.f, sc = 0.f;          // a: d/dNom, bb: d/dSp, sc: coefficient of (1 - c)
  const bool present = i < ncls - 1 && cnt > 0.f;
  if (present) {
    if (sp > 0.f) {
      const float pr = scal_step(nom / (sp + eps));
      loss -= __logf(pr);
      a -= 1.f / (pr * (sp + eps));
      bb += nom / (pr * (sp + eps) * (sp + eps));
    }
    const float rc = scal_step(nom / (cnt + eps));
    loss -= __logf(rc);
    a -= 1.f / (rc * (cnt + eps));
    const float y = M - cnt;
    if (y > 0.f) {
      const float spc = scal_step((y - (sp - nom)) / (y + eps));
      loss -= __logf(spc);
      sc += 1.f / (spc * (y + eps));                        // d/dX = -1/(spc (y+eps)), dX/dp = -(1 - c)
    }
  }
  const float npresent = warp_sum(present ? 1.f : 0.f);
  const float lsem = npresent > 0.f ? warp_sum(loss) / npresent : 0.f;
  const float ks = npresent > 0.f ? w_sem / npresent : 0.f;
  float g_t = present ? ks * (a + bb) : 0.f;                // label == i: c = 1
  float g_n = present ? ks * (bb + sc) : 0.f;               // label != i: c = 0
  float lgeo = 0.f;
  if (i == non_empty) {
    // geo_scal on p_e = p[non_empty]: nonempty prob 1 - p_e, nonempty target t != non_empty
    const float inter = (M - cnt) - (sp - nom), np = M - sp, nt = M - cnt;
    const float pr = scal_step(inter / (np + eps)), rc = scal_step(inter / (nt + eps)), spc = scal_step(nom / (cnt + eps));
    lgeo = -__logf(pr) - __logf(rc) - __logf(spc);
    // label == non_empty (c = 1): d inter = 0, d np = -1, d(spec numerator) = +1
    g_t += w_geo * (-(1.f / pr) * (inter / ((np + eps) * (np + eps))) - (1.f / spc) / (cnt + eps));
    // label != non_empty (c = 0): d inter = -1, d np = -1
    g_n += w_geo * (-(1.f / pr) * (-1.f / (np + eps) + inter / ((np + eps) * (np + eps))) + (1.f / rc) / (nt + eps));
  }
  lgeo = warp_sum(lgeo);
  if (i < ncls) {
    gt[i] = g_t;
    gn[i] = g_n;
  }
  if (i == 0) {
    out[0] = w_sem * lsem;
    out[1] = w_geo * lgeo;
  }
}


// ---- batch-statistics BatchNorm (training mode of every conv -> BN [-> ReLU] pair of the path) -------------------
// forward:  out = act(scale[c] * raw + shift[c] [+ residual]) [* gate[n][c]], raw = the convolution's bf16 output,
//           scale = gamma / sqrt(var + eps), shift = beta - mean * scale from the batch statistics of raw
//           (sums by dhd_act_bwd with act = none: [sum raw, sum raw^2]);
// backward: d raw = k1[c] * dz + k2[c] * raw + k3[c]  with  k1 = gamma/sigma, k2 = -k1 * S2 / (M sigma^2),
//           k3 = -k1 * (S1 / M - mean * S2 / (M sigma^2)),  S1 = sum dz, S2 = sum dz * (raw - mean)
//           (the standard BatchNorm backward written as a per-channel affine combination of dz and raw).
__global__ void __launch_bounds__(256)
bn_apply_kernel(const __nv_bfloat16* __restrict__ raw, int r_ld, int r_coff, long rows, int C,
                const float* __restrict__ scale, const float* __restrict__ shift, int act,
                const float* __restrict__ residual, long res_ld, const float* __restrict__ gate, int rows_per_img,
                __nv_bfloat16* __restrict__ ob, int o_ld, int o_coff, float* __restrict__ of, long f_ld,
                const __nv_bfloat16* __restrict__ res16, int res16_ld, int res16_coff) {
  const int cg = C / 8;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cg) return;
  int c;
  const long r = fast_div(i, cg, &c);
  c *= 8;
  float v[8];
  load8(raw + r * r_ld + r_coff + c, v);
  const float4 s0 = __ldg(reinterpret_cast<const float4*>(scale + c)), s1 = __ldg(reinterpret_cast<const float4*>(scale + c) + 1);
  const float4 t0 = __ldg(reinterpret_cast<const float4*>(shift + c)), t1 = __ldg(reinterpret_cast<const float4*>(shift + c) + 1);
  const float sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
  const float sh[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], sc[j], sh[j]);
  if (residual != nullptr) {
    const float4 a = *reinterpret_cast<const float4*>(residual + r * res_ld + c);
    const float4 b = *reinterpret_cast<const float4*>(residual + r * res_ld + c + 4);
    v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
  }
  if (res16 != nullptr) {                          // bf16 identity path (Bottleneck blocks of the image backbone)
    float q[8];
    load8(res16 + r * res16_ld + res16_coff + c, q);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += q[j];
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (act == 1) v[j] = fmaxf(v[j], 0.f);
    else if (act == 2) v[j] = __fdividef(1.f, 1.f + __expf(-v[j]));
  }
  if (gate != nullptr) {
    const float* g = gate + fast_div(r, rows_per_img) * C + c;
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] *= __ldg(g + j);
  }
  if (ob != nullptr) store8(ob + r * o_ld + o_coff + c, v);
  if (of != nullptr) {
    *reinterpret_cast<float4*>(of + r * f_ld + c) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(of + r * f_ld + c + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
}

// 256 threads = (256 / cg) rows x cg channel groups of 8; the three coefficient vectors of a thread's 8 channels stay
// in registers while it walks the rows (grid-stride), so a row costs two 16-byte loads and one 16-byte store
__global__ void __launch_bounds__(256)
affine_combine_kernel(const __nv_bfloat16* __restrict__ a, int a_ld, int a_coff, const __nv_bfloat16* __restrict__ b, int b_ld,
                      int b_coff, long rows, int C, const float* __restrict__ k1, const float* __restrict__ k2,
                      const float* __restrict__ k3, __nv_bfloat16* __restrict__ out, int o_ld, int o_coff) {
  const int cg = C / 8, rpb = blockDim.x / cg;
  const int gi = threadIdx.x % cg, ri = threadIdx.x / cg;
  if (ri >= rpb) return;
  const int c = gi * 8;
  float q1[8], q2[8], q3[8];
  {
    const float4 u0 = __ldg(reinterpret_cast<const float4*>(k1 + c)), u1 = __ldg(reinterpret_cast<const float4*>(k1 + c) + 1);
    const float4 v0 = __ldg(reinterpret_cast<const float4*>(k2 + c)), v1 = __ldg(reinterpret_cast<const float4*>(k2 + c) + 1);
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(k3 + c)), w1 = __ldg(reinterpret_cast<const float4*>(k3 + c) + 1);
    q1[0] = u0.x; q1[1] = u0.y; q1[2] = u0.z; q1[3] = u0.w; q1[4] = u1.x; q1[5] = u1.y; q1[6] = u1.z; q1[7] = u1.w;
    q2[0] = v0.x; q2[1] = v0.y; q2[2] = v0.z; q2[3] = v0.w; q2[4] = v1.x; q2[5] = v1.y; q2[6] = v1.z; q2[7] = v1.w;
    q3[0] = w0.x; q3[1] = w0.y; q3[2] = w0.z; q3[3] = w0.w; q3[4] = w1.x; q3[5] = w1.y; q3[6] = w1.z; q3[7] = w1.w;
  }
  const long step = (long)gridDim.x * rpb;
  for (long r = (long)blockIdx.x * rpb + ri; r < rows; r += step) {
    float x[8], y[8];
    load8(a + r * a_ld + a_coff + c, x);
    load8(b + r * b_ld + b_coff + c, y);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = fmaf(q1[j], x[j], fmaf(q2[j], y[j], q3[j]));
    store8(out + r * o_ld + o_coff + c, x);
  }
}


// per-channel coefficients of the two BatchNorm passes in ONE small launch each (instead of a dozen tensor ops)
__global__ void __launch_bounds__(256)
bn_fwd_coeffs_kernel(const float* __restrict__ sums, int C, float M, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float eps, float momentum, float* __restrict__ running_mean,
                     float* __restrict__ running_var, float* __restrict__ scale, float* __restrict__ shift,
                     float* __restrict__ mean_out, float* __restrict__ invstd_out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float mean = sums[c] / M;
  const float var = fmaxf(sums[C + c] / M - mean * mean, 0.f);
  const float invstd = rsqrtf(var + eps);
  const float sc = gamma[c] * invstd;
  scale[c] = sc;
  shift[c] = beta[c] - mean * sc;
  mean_out[c] = mean;
  invstd_out[c] = invstd;
  if (running_mean != nullptr) {
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * var * (M / fmaxf(M - 1.f, 1.f));
  }
}

__global__ void __launch_bounds__(256)
bn_bwd_coeffs_kernel(const float* __restrict__ sums, int C, int sums_stride, float M, const float* __restrict__ mean,
                     const float* __restrict__ invstd, const float* __restrict__ gamma, float* __restrict__ k1,
                     float* __restrict__ k2, float* __restrict__ k3, float* __restrict__ dgamma,
                     float* __restrict__ dbeta) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float s1 = sums[c];
  const float s2 = sums[sums_stride + c] - mean[c] * s1;        // sum dz * (raw - mean)
  const float is = invstd[c];
  const float a = gamma[c] * is;
  const float b = -a * s2 * is * is / M;
  k1[c] = a;
  k2[c] = b;
  k3[c] = -a * (s1 / M) - b * mean[c];
  if (dgamma != nullptr) dgamma[c] += s2 * is;
  if (dbeta != nullptr) dbeta[c] += s1;
}


// The fixed-order finish of per-block partial sums [rows][2C] fused with the per-channel coefficient formulas above
// (one launch instead of finish + coefficients): a block owns 32 channels, lane = channel, its 32 warps stride over the
// rows for column c and column C + c, the warp partials meet in shared memory, warp 0 adds them in order.
// MODE 1: BatchNorm forward (sums = [sum raw, sum raw^2]); MODE 2: backward (sums = [sum dz, sum dz*raw]).
struct BnFinishArgs {
  float M, eps, momentum;
  const float *gamma, *beta, *mean_in, *invstd_in;
  float *running_mean, *running_var, *o0, *o1, *o2, *o3, *dgamma, *dbeta;
};
template <int MODE>
__global__ void __launch_bounds__(1024)
bn_finish_kernel(const float* __restrict__ partial, int rows, int C, const BnFinishArgs A) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
  if (c < C) {
    // every load of this loop is a DRAM / L2 round trip on a one-wave grid of 8 blocks: eight independent loads per trip
    const size_t n = 2 * (size_t)C;
    int r = w;
    for (; r + 96 < rows; r += 128) {
      const float* p0 = partial + (size_t)r * n + c;
      const float* p1 = p0 + 32 * n;
      const float* p2 = p0 + 64 * n;
      const float* p3 = p0 + 96 * n;
      a0 += __ldg(p0);
      b0 += __ldg(p0 + C);
      a1 += __ldg(p1);
      b1 += __ldg(p1 + C);
      a2 += __ldg(p2);
      b2 += __ldg(p2 + C);
      a3 += __ldg(p3);
      b3 += __ldg(p3 + C);
    }
    for (; r < rows; r += 32) {
      a0 += __ldg(partial + (size_t)r * n + c);
      b0 += __ldg(partial + (size_t)r * n + C + c);
    }
  }
  a0 += a2;
  a1 += a3;
  b0 += b2;
  b1 += b3;
  __shared__ float sa[32][33], sb[32][33];
  sa[w][lane] = a0 + a1;
  sb[w][lane] = b0 + b1;
  __syncthreads();
  if (w != 0 || c >= C) return;
  float s1 = sa[0][lane], s2 = sb[0][lane];
#pragma unroll
  for (int k = 1; k < 32; ++k) {
    s1 += sa[k][lane];
    s2 += sb[k][lane];
  }
  if (MODE == 1) {
    const float mean = s1 / A.M;
    const float var = fmaxf(s2 / A.M - mean * mean, 0.f);
    const float invstd = rsqrtf(var + A.eps);
    const float sc = A.gamma[c] * invstd;
    A.o0[c] = sc;                               // scale
    A.o1[c] = A.beta[c] - mean * sc;            // shift
    A.o2[c] = mean;
    A.o3[c] = invstd;
    if (A.running_mean != nullptr) {
      A.running_mean[c] = (1.f - A.momentum) * A.running_mean[c] + A.momentum * mean;
      A.running_var[c] = (1.f - A.momentum) * A.running_var[c] + A.momentum * var * (A.M / fmaxf(A.M - 1.f, 1.f));
    }
  } else {
    const float mean = A.mean_in[c], is = A.invstd_in[c];
    const float t2 = s2 - mean * s1;            // sum dz * (raw - mean)
    const float a = A.gamma[c] * is;
    const float b = -a * t2 * is * is / A.M;
    A.o0[c] = a;                                // k1
    A.o1[c] = b;                                // k2
    A.o2[c] = -a * (s1 / A.M) - b * mean;       // k3
    if (A.dgamma != nullptr) A.dgamma[c] += t2 * is;
    if (A.dbeta != nullptr) A.dbeta[c] += s1;
  }
}

// Dropout (nn.Dropout(0.5) behind the ASPP, depthnet.py:81, 106) as a counter-based mask: Philox-4x32-10 keyed by the
// seed, counter = (vector index, step, salt), one call per 8 channels, 16 random bits per element.  The mask is a pure
// function of (seed, step, salt, element), so the backward multiplies the gradient by the very same mask without it
// ever being stored; `rng` lives in device memory so a captured CUDA graph sees a new step on every replay.
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}

__global__ void __launch_bounds__(256)
dropout_kernel(__nv_bfloat16* __restrict__ x, int ld, int coff, long rows, int C, uint32_t thresh16, float scale,
               const long long* __restrict__ rng, uint32_t salt) {
  const int vec_per_row = C >> 3;
  const long v = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= rows * vec_per_row) return;
  int c;
  const long r = fast_div(v, vec_per_row, &c);
  c <<= 3;
  const unsigned long long seed = (unsigned long long)rng[0], step = (unsigned long long)rng[1];
  const uint4 rnd = philox4x32_10(make_uint4((uint32_t)v, (uint32_t)((unsigned long long)v >> 32), (uint32_t)step, salt),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
  float f[8];
  __nv_bfloat16* p = x + r * ld + coff + c;
  load8(p, f);
  const uint32_t w[4] = {rnd.x, rnd.y, rnd.z, rnd.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const uint32_t r16 = (w[i >> 1] >> ((i & 1) * 16)) & 0xffffu;
    f[i] = r16 >= thresh16 ? f[i] * scale : 0.f;           // keep with probability 1 - p
  }
  store8(p, f);
}


// MGHS.get_downsampled_gt_depth / get_downsampled_gt_height (lss_heightmap.py:625-701): ds x ds min-pool of a sparse
// LiDAR map with zeros ignored (the reference's 1e5 sentinel), then t = (min - lo) / interval as two separately
// rounded fp32 operations, bins outside [0, nbins + 1) -> 0, .long() truncation; the reference's one-hot row
// one_hot(t, nbins + 1)[1:] is returned as its index: label = t - 1 (-1 = all-zero row), valid = label >= 0.
// One warp per output pixel.
__global__ void __launch_bounds__(256)
gt_downsample_kernel(const float* __restrict__ gt, long nout, int H, int W, int ds, float lo, float interval, int nbins,
                     int32_t* __restrict__ label, uint8_t* __restrict__ valid) {
  const long o = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (o >= nout) return;
  const int oW = W / ds, oH = H / ds;
  const int ox = (int)(o % oW);
  const long t0 = o / oW;
  const int oy = (int)(t0 % oH);
  const long img = t0 / oH;
  const float* base = gt + (img * H + (long)oy * ds) * W + (long)ox * ds;
  float m = 1e5f;
  for (int e = lane; e < ds * ds; e += 32) {
    const float v = __ldg(base + (long)(e / ds) * W + (e % ds));
    m = fminf(m, v == 0.f ? 1e5f : v);
  }
#pragma unroll
  for (int k = 16; k > 0; k >>= 1) m = fminf(m, __shfl_xor_sync(kFull, m, k));
  if (lane == 0) {
    float t = __fdiv_rn(__fsub_rn(m, lo), interval);
    if (!(t < (float)(nbins + 1) && t >= 0.f)) t = 0.f;
    const int bin = (int)t - 1;
    label[o] = bin;
    if (valid != nullptr) valid[o] = bin >= 0 ? 1 : 0;
  }
}

}  // namespace dhd

using namespace dhd;

static inline bool ok8(int C, int ld, int coff, const void* p) {
  return C % 8 == 0 && ld % 8 == 0 && coff % 8 == 0 && ((uintptr_t)p & 15) == 0;
}

extern "C" size_t dhd_act_bwd_workspace_bytes(int C) { return (size_t)kMaxSumBlocks * 2 * C * sizeof(float); }

extern "C" int dhd_act_bwd(const void* dy, int dy_ld, int dy_coff, const void* y, int y_ld, int y_coff, long rows,
                           int C, int act, void* out, int out_ld, int out_coff, float* colsum, float* workspace,
                           const void* add, int add_ld, int add_coff, void* stream) {
  DHD_REQUIRE(dy != nullptr && rows > 0 && C > 0, "bad arguments");
  DHD_REQUIRE(act >= 0 && act <= 3, "act must be none / relu / sigmoid / softplus");
  DHD_REQUIRE(act == 0 || y != nullptr, "the activation derivative needs the saved output");
  DHD_REQUIRE(out != nullptr || colsum != nullptr, "nothing to compute");
  DHD_REQUIRE(colsum == nullptr || (workspace != nullptr && y != nullptr), "column sums need the workspace and y");
  DHD_REQUIRE(C <= 2048 && ok8(C, dy_ld, dy_coff, dy), "dy: C % 8, 16-byte aligned rows");
  if (y != nullptr) DHD_REQUIRE(ok8(C, y_ld, y_coff, y), "y: 16-byte aligned rows");
  if (out != nullptr) DHD_REQUIRE(ok8(C, out_ld, out_coff, out), "out: 16-byte aligned rows");
  if (add != nullptr) DHD_REQUIRE(ok8(C, add_ld, add_coff, add), "add: 16-byte aligned rows");
  cudaStream_t st = (cudaStream_t)stream;
  const int rpb = 256 / (C / 8);
  long rows_per_block = (rows + kMaxSumBlocks - 1) / kMaxSumBlocks;
  if (rows_per_block < 4L * rpb) rows_per_block = 4L * rpb;
  const int nblocks = (int)((rows + rows_per_block - 1) / rows_per_block);
  act_bwd_kernel<<<nblocks, 256, 0, st>>>((const __nv_bfloat16*)dy, dy_ld, dy_coff, (const __nv_bfloat16*)y, y_ld,
                                          y_coff, rows, C, act, (__nv_bfloat16*)out, out_ld, out_coff,
                                          colsum != nullptr ? workspace : nullptr, rows_per_block,
                                          (const __nv_bfloat16*)add, add_ld, add_coff);
  DHD_CUDA_LAUNCH_CHECK("act_bwd");
  if (colsum != nullptr) {
    colsum_finish_kernel<<<(2 * C + 31) / 32, 1024, 0, st>>>(workspace, nblocks, 2 * C, colsum);
    DHD_CUDA_LAUNCH_CHECK("colsum_reduce");
  }
  return DHD_OK;
}

extern "C" int dhd_colsum_finish(const float* partial, int rows, int n, float* sums, void* stream) {
  DHD_REQUIRE(partial && sums && rows > 0 && n > 0, "bad arguments");
  colsum_finish_kernel<<<(n + 31) / 32, 1024, 0, (cudaStream_t)stream>>>(partial, rows, n, sums);
  DHD_CUDA_LAUNCH_CHECK("colsum_finish");
  return DHD_OK;
}

extern "C" size_t dhd_occ_loss_workspace_bytes(void) {
  return ((size_t)148 * 16 * kScalRow + 64) * sizeof(float);
}

extern "C" int dhd_occ_ce_loss(const float* logits, const uint8_t* labels, const uint8_t* mask,
                               const float* class_weight, int ncls, int ignore_index, int B, int Dx, int Dy, int Dz,
                               float loss_weight, float weight_sem, float weight_geo, int non_empty_idx,
                               float* losses, void* dlogits, int dl_ld, float* workspace, void* stream) {
  DHD_REQUIRE(logits && labels && losses && dlogits, "null pointer");
  DHD_REQUIRE(B > 0 && Dx > 0 && Dy > 0 && Dz > 0 && ncls > 0 && ncls <= 32, "bad shape (ncls <= 32)");
  DHD_REQUIRE(dl_ld >= Dz * ncls, "dlogits rows are too short");
  DHD_REQUIRE((long)B * Dx * Dy * Dz < (1L << 31), "too many voxels for 32-bit indexing");
  const bool scal = weight_sem != 0.f || weight_geo != 0.f;
  DHD_REQUIRE(!scal || (workspace != nullptr && non_empty_idx >= 0 && non_empty_idx < ncls), "scal terms need the workspace");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(losses, 0, 4 * sizeof(float), st);
  if (e != cudaSuccess) return fail((int)e, "%s: %ld", "memset(loss)", (long)e);
  const long nvox = (long)B * Dx * Dy * Dz;
  const int blocks = (int)min((nvox + 255) / 256, (long)sm_count() * 16);
  float *gt = nullptr, *gn = nullptr;
  if (scal) {
    const int sblocks = (int)min((nvox + 255) / 256, (long)148 * 16);
    gt = workspace + (size_t)148 * 16 * kScalRow;
    gn = gt + 32;
    static bool stats_attr = false;
    const size_t stats_smem = (size_t)256 * 33 * sizeof(float2);
    if (!stats_attr) {
      cudaFuncSetAttribute(occ_scal_stats_kernel<18>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stats_smem);
      cudaFuncSetAttribute(occ_scal_stats_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stats_smem);
      stats_attr = true;
    }
    if (ncls <= 18)
      occ_scal_stats_kernel<18><<<sblocks, 256, stats_smem, st>>>(logits, labels, mask, ncls, ignore_index, nvox, workspace);
    else
      occ_scal_stats_kernel<32><<<sblocks, 256, stats_smem, st>>>(logits, labels, mask, ncls, ignore_index, nvox, workspace);
    DHD_CUDA_LAUNCH_CHECK("occ_scal_stats");
    occ_scal_coeffs_kernel<<<1, 1024, 0, st>>>(workspace, sblocks, ncls, non_empty_idx, weight_sem, weight_geo, losses + 2,
                                             gt, gn);
    DHD_CUDA_LAUNCH_CHECK("occ_scal_coeffs");
  }
  ce_norm_kernel<<<blocks, 256, 0, st>>>(labels, mask, class_weight, ncls, ignore_index, nvox, losses + 1);
  DHD_CUDA_LAUNCH_CHECK("ce_norm");
  if (ncls <= 18)
    ce_loss_kernel<18><<<blocks, 256, 0, st>>>(logits, labels, mask, class_weight, ncls, ignore_index, B, Dx, Dy, Dz,
                                         loss_weight, losses + 1, losses, (__nv_bfloat16*)dlogits, dl_ld, gt, gn);
  else
    ce_loss_kernel<32><<<blocks, 256, 0, st>>>(logits, labels, mask, class_weight, ncls, ignore_index, B, Dx, Dy, Dz,
                                         loss_weight, losses + 1, losses, (__nv_bfloat16*)dlogits, dl_ld, gt, gn);
  DHD_CUDA_LAUNCH_CHECK("ce_loss");
  return DHD_OK;
}

extern "C" int dhd_depth_head_bwd(const float* depth, const float* depth_grad, const float* feat_grad, int BN, int D,
                                  int HW, int C, void* out, int out_ld, void* stream) {
  DHD_REQUIRE(depth && depth_grad && feat_grad && out, "null pointer");
  DHD_REQUIRE(BN > 0 && D > 0 && HW > 0 && C > 0 && out_ld >= D + C, "bad shape");
  const long npix = (long)BN * HW;
  depth_head_bwd_kernel<<<(int)((npix + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      depth, depth_grad, feat_grad, BN, D, HW, C, (__nv_bfloat16*)out, out_ld);
  DHD_CUDA_LAUNCH_CHECK("depth_head_bwd");
  return DHD_OK;
}

static int sfa_pb(int N, int HW) {
  const int want = max(1, (sm_count() * 8) / max(1, N));
  return max(16, (HW + want - 1) / want);
}

extern "C" size_t dhd_sfa_gate_bwd_workspace_bytes(int N, int HW, int C) {
  const int pb = sfa_pb(N, HW);
  return (size_t)N * ((HW + pb - 1) / pb) * C * sizeof(float);
}

extern "C" int dhd_sfa_gate_bwd(int mode, const void* g, int g_ld, int g_coff, const void* x, int x_ld, int x_coff,
                                int C, int N, int HW, const float* a1, const float* a2, void* dpre2, int d_ld,
                                int d_coff, float* dx, float* a1_sums, int accumulate_sums, float* workspace,
                                void* stream) {
  DHD_REQUIRE(g && x && a1 && dx && a1_sums && workspace, "null pointer");
  DHD_REQUIRE(mode == 0 || mode == 1, "mode must be 0 (fuse) or 1 (u)");
  DHD_REQUIRE(mode == 1 || (a2 != nullptr && dpre2 != nullptr), "mode 0 needs a2 and dpre2");
  DHD_REQUIRE(N > 0 && HW > 0 && C > 0 && C <= 2048, "bad shape");
  DHD_REQUIRE(ok8(C, g_ld, g_coff, g) && ok8(C, x_ld, x_coff, x) && ((uintptr_t)dx & 15) == 0 &&
                  ((uintptr_t)a1 & 15) == 0, "needs C % 8 == 0 and 16-byte aligned rows");
  if (mode == 0) DHD_REQUIRE(ok8(C, d_ld, d_coff, dpre2) && ((uintptr_t)a2 & 15) == 0, "dpre2 / a2 alignment");
  cudaStream_t st = (cudaStream_t)stream;
  const int pb = sfa_pb(N, HW), nb = (HW + pb - 1) / pb;
  sfa_gate_bwd_kernel<false><<<dim3(nb, N), 256, 0, st>>>(mode, (const __nv_bfloat16*)g, g_ld, g_coff,
                                                          (const __nv_bfloat16*)x, x_ld, x_coff, C, HW, pb, a1, a2,
                                                          (__nv_bfloat16*)dpre2, d_ld, d_coff, dx, workspace, nullptr, 0, 0);
  DHD_CUDA_LAUNCH_CHECK("sfa_gate_bwd");
  image_sum_reduce_kernel<<<(N * C + 255) / 256, 256, 0, st>>>(workspace, nb, N, C, a1_sums, accumulate_sums);
  DHD_CUDA_LAUNCH_CHECK("image_sum_reduce");
  return DHD_OK;
}

extern "C" int dhd_sfa_gate_bwd_b16(int mode, const void* g, int g_ld, int g_coff, const void* x, int x_ld, int x_coff,
                                    int C, int N, int HW, const float* a1, const float* a2, void* dpre2, int d_ld,
                                    int d_coff, void* dx, int dx_ld, int dx_coff, float* a1_sums, int accumulate_sums,
                                    float* workspace, void* stream) {
  DHD_REQUIRE(g && x && a1 && a1_sums && workspace, "null pointer");
  DHD_REQUIRE(mode == 0 || mode == 1, "mode must be 0 (fuse) or 1 (u)");
  DHD_REQUIRE(mode == 1 || (a2 != nullptr && dpre2 != nullptr && dx != nullptr), "mode 0 needs a2, dpre2 and dx");
  DHD_REQUIRE(N > 0 && HW > 0 && C > 0 && C <= 2048, "bad shape");
  DHD_REQUIRE(ok8(C, g_ld, g_coff, g) && ok8(C, x_ld, x_coff, x) && ((uintptr_t)a1 & 15) == 0,
              "needs C % 8 == 0 and 16-byte aligned rows");
  if (mode == 0)
    DHD_REQUIRE(ok8(C, d_ld, d_coff, dpre2) && ((uintptr_t)a2 & 15) == 0 && ok8(C, dx_ld, dx_coff, dx) && dx_coff + 2 * C <= dx_ld,
                "dpre2 / a2 / dx alignment");
  cudaStream_t st = (cudaStream_t)stream;
  const int pb = sfa_pb(N, HW), nb = (HW + pb - 1) / pb;
  sfa_gate_bwd_kernel<true><<<dim3(nb, N), 256, 0, st>>>(mode, (const __nv_bfloat16*)g, g_ld, g_coff, (const __nv_bfloat16*)x,
                                                         x_ld, x_coff, C, HW, pb, a1, a2, (__nv_bfloat16*)dpre2, d_ld, d_coff,
                                                         nullptr, workspace, (__nv_bfloat16*)dx, dx_ld, dx_coff);
  DHD_CUDA_LAUNCH_CHECK("sfa_gate_bwd_b16");
  image_sum_reduce_kernel<<<(N * C + 255) / 256, 256, 0, st>>>(workspace, nb, N, C, a1_sums, accumulate_sums);
  DHD_CUDA_LAUNCH_CHECK("image_sum_reduce");
  return DHD_OK;
}

extern "C" int dhd_sfa_dx_combine(const void* dxb, int b_ld, int b_coff, const void* du, int u_ld, int u_coff,
                                  const float* a1, const float* ds, int C, int N, int HW, void* out, int o_ld, int o_coff,
                                  void* stream) {
  DHD_REQUIRE(dxb && du && a1 && out && N > 0 && HW > 0 && C > 0, "bad arguments");
  DHD_REQUIRE(ok8(C, b_ld, b_coff, dxb) && ok8(C, u_ld, u_coff, du) && ok8(C, o_ld, o_coff, out) &&
                  b_coff + 2 * C <= b_ld && o_coff + 2 * C <= o_ld && ((uintptr_t)a1 & 15) == 0,
              "needs C % 8 == 0, 16-byte aligned rows holding 2C channels");
  const long total = (long)N * HW * (C / 8);
  sfa_dx_combine_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)dxb, b_ld, b_coff, (const __nv_bfloat16*)du, u_ld, u_coff, a1, ds, C, (long)N * HW, HW,
      (__nv_bfloat16*)out, o_ld, o_coff);
  DHD_CUDA_LAUNCH_CHECK("sfa_dx_combine");
  return DHD_OK;
}

extern "C" int dhd_add_rowvec(const float* in, const float* v, int N, int HW, int C, void* out, int out_ld,
                              int out_coff, void* stream) {
  DHD_REQUIRE(in && out, "null pointer");
  DHD_REQUIRE(N > 0 && HW > 0 && ok8(C, out_ld, out_coff, out) && ((uintptr_t)in & 15) == 0, "bad shape / alignment");
  const long total = (long)N * HW * (C / 8);
  add_rowvec_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, v, (long)N * HW, HW, C,
                                                                                 (__nv_bfloat16*)out, out_ld, out_coff);
  DHD_CUDA_LAUNCH_CHECK("add_rowvec");
  return DHD_OK;
}

extern "C" int dhd_se_gate_bwd(const void* dh, int g_ld, int g_coff, const void* h, int h_ld, int h_coff, int C, int N,
                               int HW, const float* gate, void* dpre, int d_ld, int d_coff, float* gate_sums,
                               float* workspace, void* stream) {
  DHD_REQUIRE(dh && h && gate && dpre && gate_sums && workspace, "null pointer");
  DHD_REQUIRE(N > 0 && HW > 0 && C > 0 && C <= 2048, "bad shape");
  DHD_REQUIRE(ok8(C, g_ld, g_coff, dh) && ok8(C, h_ld, h_coff, h) && ok8(C, d_ld, d_coff, dpre) &&
                  ((uintptr_t)gate & 15) == 0, "needs C % 8 == 0 and 16-byte aligned rows");
  cudaStream_t st = (cudaStream_t)stream;
  const int pb = sfa_pb(N, HW), nb = (HW + pb - 1) / pb;
  se_gate_bwd_kernel<<<dim3(nb, N), 256, 0, st>>>((const __nv_bfloat16*)dh, g_ld, g_coff, (const __nv_bfloat16*)h, h_ld,
                                                  h_coff, C, HW, pb, gate, (__nv_bfloat16*)dpre, d_ld, d_coff, workspace);
  DHD_CUDA_LAUNCH_CHECK("se_gate_bwd");
  image_sum_reduce_kernel<<<(N * C + 255) / 256, 256, 0, st>>>(workspace, nb, N, C, gate_sums, 0);
  DHD_CUDA_LAUNCH_CHECK("image_sum_reduce");
  return DHD_OK;
}

extern "C" int dhd_height_loss(const float* height, const int32_t* label, const uint8_t* fg, int BN, int H, int HW,
                               float weight, const float* n_fg, float* loss, void* dz, int dz_ld, void* stream) {
  DHD_REQUIRE(height && label && fg && n_fg && loss && dz, "null pointer");
  DHD_REQUIRE(BN > 0 && H > 0 && HW > 0 && dz_ld >= H, "bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(loss, 0, sizeof(float), st);
  if (e != cudaSuccess) return fail((int)e, "%s: %ld", "memset(loss)", (long)e);
  const long npix = (long)BN * HW;
  height_loss_kernel<<<(int)((npix + 127) / 128), 128, 0, st>>>(height, label, fg, BN, H, HW, weight, n_fg, loss,
                                                               (__nv_bfloat16*)dz, dz_ld);
  DHD_CUDA_LAUNCH_CHECK("height_loss");
  return DHD_OK;
}

extern "C" int dhd_dcn_col2im_bwd(const void* dcol, int col_ld, const void* x, int x_ld, int x_coff, int C, int N, int H,
                                  int W, const float* offset, int off_ld, int ksize, int pad, int dilation, int groups,
                                  float* dx, float* doff, void* stream) {
  DHD_REQUIRE(dcol && x && offset && dx && doff, "null pointer");
  DHD_REQUIRE(C > 0 && groups > 0 && C % groups == 0 && (C / groups) % 8 == 0, "bad channel grouping");
  DHD_REQUIRE(N > 0 && H > 0 && W > 0 && ksize >= 1 && ksize <= 3, "bad shape");
  DHD_REQUIRE(col_ld % 8 == 0 && x_ld % 8 == 0 && x_coff % 8 == 0 && ((uintptr_t)dcol & 15) == 0 &&
                  ((uintptr_t)x & 15) == 0, "16-byte aligned rows");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(dx, 0, (size_t)N * H * W * C * sizeof(float), st);
  if (e != cudaSuccess) return fail((int)e, "%s: %ld", "memset(dx)", (long)e);
  const long warps = (long)N * H * W * ksize * ksize;
  DHD_REQUIRE(warps < (1L << 26), "too many sampling points for 32-bit indexing");
  dcn_col2im_bwd_kernel<<<(int)((warps * 32 + 255) / 256), 256, 0, st>>>(
      (const __nv_bfloat16*)dcol, col_ld, (const __nv_bfloat16*)x, x_ld, x_coff, C, N, H, W, offset, off_ld, ksize, pad,
      dilation, groups, dx, doff);
  DHD_CUDA_LAUNCH_CHECK("dcn_col2im_bwd");
  return DHD_OK;
}

extern "C" int dhd_pack_conv_weights(const float* w, int Cout, int cin_total, int taps, int col_lo, int Cin,
                                     const float* scale, void* fwd, int cin_pad, void* bwd, int cout_pad, int bwd_mode,
                                     void* stream) {
  DHD_REQUIRE(w != nullptr && (fwd != nullptr || bwd != nullptr), "null pointer");
  DHD_REQUIRE(Cout > 0 && Cin > 0 && taps >= 1 && col_lo >= 0 && col_lo + Cin <= cin_total, "bad shape");
  DHD_REQUIRE(cin_pad >= Cin && cout_pad >= Cout && (bwd_mode == 0 || bwd_mode == 1), "bad padding / mode");
  const long n = (fwd ? (long)Cout * taps * cin_pad : 0) + (bwd ? (long)Cin * taps * cout_pad : 0);
  const int blocks = (int)min((n + 255) / 256, (long)sm_count() * 8);
  pack_conv_weights_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(w, Cout, cin_total, taps, col_lo, Cin, scale,
                                                                    (__nv_bfloat16*)fwd, cin_pad, (__nv_bfloat16*)bwd,
                                                                    cout_pad, bwd_mode);
  DHD_CUDA_LAUNCH_CHECK("pack_conv_weights");
  return DHD_OK;
}

extern "C" int dhd_pack_conv_weights_batch(const dhd_pack_desc* descs, int n, void* stream) {
  DHD_REQUIRE(descs != nullptr && n >= 1 && n <= DHD_PACK_MAX_BATCH, "1..DHD_PACK_MAX_BATCH layers");
  PackBatch B;
  for (int i = 0; i < n; ++i) {
    const dhd_pack_desc& L = descs[i];
    DHD_REQUIRE(L.w && (L.fwd || L.bwd), "null pointer");
    DHD_REQUIRE(L.Cout > 0 && L.Cin > 0 && L.taps >= 1 && L.col_lo >= 0 && L.col_lo + L.Cin <= L.cin_total, "bad shape");
    DHD_REQUIRE(L.cin_pad >= L.Cin && L.cout_pad >= L.Cout && (L.bwd_mode == 0 || L.bwd_mode == 1), "bad padding / mode");
    B.d[i] = L;
    const long e = (L.fwd ? (long)L.Cout * L.taps * L.cin_pad : 0) + (L.bwd ? (long)L.Cin * L.taps * L.cout_pad : 0);
    DHD_REQUIRE(e < (1L << 31), "layer too large for 32-bit indexing");
  }
  for (int i = n; i < DHD_PACK_MAX_BATCH; ++i) B.d[i] = descs[0];
  int tiles_max = 0, taps_max = 1;
  for (int i = 0; i < n; ++i) {
    const dhd_pack_desc& L = descs[i];
    const int co_t = ((L.bwd ? L.cout_pad : L.Cout) + kPackTile - 1) / kPackTile;
    const int ci_t = ((L.fwd ? L.cin_pad : L.Cin) + kPackTile - 1) / kPackTile;
    tiles_max = co_t * ci_t > tiles_max ? co_t * ci_t : tiles_max;
    taps_max = L.taps > taps_max ? L.taps : taps_max;
  }
  DHD_REQUIRE(taps_max <= DHD_CONV_MAX_TAPS, "pack batch: at most DHD_CONV_MAX_TAPS taps per layer");
  const size_t smem = (size_t)kPackTile * (kPackTile * taps_max + 1) * sizeof(float);
  pack_conv_weights_batch_kernel<<<dim3(tiles_max, n), 256, smem, (cudaStream_t)stream>>>(B);
  DHD_CUDA_LAUNCH_CHECK("pack_conv_weights_batch");
  return DHD_OK;
}

extern "C" int dhd_adamw_flat(float* p, const float* g, float* m, float* v, long n, float lr, float beta1, float beta2,
                              float eps, float weight_decay, float bc1, float bc2, const float* grad_scale, void* stream) {
  DHD_REQUIRE(p && g && m && v && n > 0, "null pointer / empty");
  DHD_REQUIRE(bc1 > 0.f && bc2 > 0.f && lr >= 0.f, "bad step constants");
  const long threads = (n + 3) / 4;
  adamw_flat_kernel<<<(int)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, lr, beta1, beta2, eps,
                                                                                  weight_decay, bc1, sqrtf(bc2), grad_scale);
  DHD_CUDA_LAUNCH_CHECK("adamw_flat");
  return DHD_OK;
}

static int bn_apply_launch(const void* raw, int raw_ld, int raw_coff, long rows, int C, const float* scale,
                           const float* shift, int act, const float* residual, long res_ld, const float* gate,
                           int rows_per_img, void* out_b16, int o_ld, int o_coff, float* out_f32, long f_ld,
                           const void* res16, int res16_ld, int res16_coff, void* stream) {
  DHD_REQUIRE(raw && scale && shift && (out_b16 || out_f32) && rows > 0 && C > 0, "bad arguments");
  DHD_REQUIRE(act >= 0 && act <= 2, "act must be none / relu / sigmoid");
  DHD_REQUIRE(ok8(C, raw_ld, raw_coff, raw) && ((uintptr_t)scale & 15) == 0 && ((uintptr_t)shift & 15) == 0,
              "raw: C % 8, 16-byte aligned rows");
  if (out_b16 != nullptr) DHD_REQUIRE(ok8(C, o_ld, o_coff, out_b16), "out_b16: 16-byte aligned rows");
  if (out_f32 != nullptr) DHD_REQUIRE(f_ld % 4 == 0 && ((uintptr_t)out_f32 & 15) == 0, "out_f32: 16-byte aligned rows");
  if (residual != nullptr) DHD_REQUIRE(res_ld % 4 == 0 && ((uintptr_t)residual & 15) == 0, "residual: 16-byte aligned rows");
  if (res16 != nullptr) DHD_REQUIRE(ok8(C, res16_ld, res16_coff, res16), "bf16 residual: 16-byte aligned rows");
  DHD_REQUIRE(gate == nullptr || rows_per_img > 0, "gate needs rows_per_img");
  const long total = rows * (C / 8);
  bn_apply_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)raw, raw_ld, raw_coff, rows, C, scale, shift, act, residual, res_ld, gate, rows_per_img,
      (__nv_bfloat16*)out_b16, o_ld, o_coff, out_f32, f_ld, (const __nv_bfloat16*)res16, res16_ld, res16_coff);
  DHD_CUDA_LAUNCH_CHECK("bn_apply");
  return DHD_OK;
}

extern "C" int dhd_bn_apply(const void* raw, int raw_ld, int raw_coff, long rows, int C, const float* scale,
                            const float* shift, int act, const float* residual, long res_ld, const float* gate,
                            int rows_per_img, void* out_b16, int o_ld, int o_coff, float* out_f32, long f_ld,
                            void* stream) {
  return bn_apply_launch(raw, raw_ld, raw_coff, rows, C, scale, shift, act, residual, res_ld, gate, rows_per_img, out_b16,
                         o_ld, o_coff, out_f32, f_ld, nullptr, 0, 0, stream);
}

extern "C" int dhd_bn_apply_res16(const void* raw, int raw_ld, int raw_coff, long rows, int C, const float* scale,
                                  const float* shift, int act, const void* res16, int res16_ld, int res16_coff,
                                  void* out_b16, int o_ld, int o_coff, void* stream) {
  return bn_apply_launch(raw, raw_ld, raw_coff, rows, C, scale, shift, act, nullptr, 0, nullptr, 0, out_b16, o_ld, o_coff,
                         nullptr, 0, res16, res16_ld, res16_coff, stream);
}

extern "C" int dhd_affine_combine(const void* a, int a_ld, int a_coff, const void* b, int b_ld, int b_coff, long rows, int C,
                                  const float* k1, const float* k2, const float* k3, void* out, int o_ld, int o_coff,
                                  void* stream) {
  DHD_REQUIRE(a && b && k1 && k2 && k3 && out && rows > 0 && C > 0, "bad arguments");
  DHD_REQUIRE(ok8(C, a_ld, a_coff, a) && ok8(C, b_ld, b_coff, b) && ok8(C, o_ld, o_coff, out), "C % 8, 16-byte aligned rows");
  DHD_REQUIRE(C / 8 <= 256 && ((uintptr_t)k1 & 15) == 0 && ((uintptr_t)k2 & 15) == 0 && ((uintptr_t)k3 & 15) == 0,
              "C <= 2048 and 16-byte aligned coefficient vectors");
  const int rpb = 256 / (C / 8);
  const long want = (rows + rpb - 1) / rpb;
  const long cap = (long)sm_count() * 16;
  affine_combine_kernel<<<(int)(want < cap ? want : cap), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)a, a_ld, a_coff, (const __nv_bfloat16*)b, b_ld, b_coff, rows, C, k1, k2, k3,
      (__nv_bfloat16*)out, o_ld, o_coff);
  DHD_CUDA_LAUNCH_CHECK("affine_combine");
  return DHD_OK;
}

extern "C" int dhd_bn_fwd_coeffs(const float* sums, int C, float M, const float* gamma, const float* beta, float eps,
                                 float momentum, float* running_mean, float* running_var, float* scale, float* shift,
                                 float* mean, float* invstd, void* stream) {
  DHD_REQUIRE(sums && gamma && beta && scale && shift && mean && invstd && C > 0 && M > 0.f, "bad arguments");
  DHD_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "running statistics come as a pair");
  bn_fwd_coeffs_kernel<<<(C + 255) / 256, 256, 0, (cudaStream_t)stream>>>(sums, C, M, gamma, beta, eps, momentum,
                                                                         running_mean, running_var, scale, shift, mean,
                                                                         invstd);
  DHD_CUDA_LAUNCH_CHECK("bn_fwd_coeffs");
  return DHD_OK;
}

extern "C" int dhd_bn_fwd_coeffs_partial(const float* partial, int rows, int C, float M, const float* gamma,
                                         const float* beta, float eps, float momentum, float* running_mean,
                                         float* running_var, float* scale, float* shift, float* mean, float* invstd,
                                         void* stream) {
  DHD_REQUIRE(partial && rows > 0 && gamma && beta && scale && shift && mean && invstd && C > 0 && M > 0.f, "bad arguments");
  DHD_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "running statistics come as a pair");
  BnFinishArgs A = {};
  A.M = M; A.eps = eps; A.momentum = momentum; A.gamma = gamma; A.beta = beta;
  A.running_mean = running_mean; A.running_var = running_var;
  A.o0 = scale; A.o1 = shift; A.o2 = mean; A.o3 = invstd;
  bn_finish_kernel<1><<<(C + 31) / 32, 1024, 0, (cudaStream_t)stream>>>(partial, rows, C, A);
  DHD_CUDA_LAUNCH_CHECK("bn_finish_fwd");
  return DHD_OK;
}

extern "C" int dhd_bn_bwd_sums_coeffs(const void* dy, int dy_ld, int dy_coff, const void* raw, int raw_ld, int raw_coff,
                                      long rows, int C, float* workspace, float M, const float* mean, const float* invstd,
                                      const float* gamma, float* k1, float* k2, float* k3, float* dgamma, float* dbeta,
                                      void* stream) {
  DHD_REQUIRE(dy && raw && workspace && mean && invstd && gamma && k1 && k2 && k3 && rows > 0 && C > 0 && M > 0.f,
              "bad arguments");
  DHD_REQUIRE(C <= 2048 && ok8(C, dy_ld, dy_coff, dy) && ok8(C, raw_ld, raw_coff, raw), "C % 8, 16-byte aligned rows");
  cudaStream_t st = (cudaStream_t)stream;
  const int rpb = 256 / (C / 8);
  long rows_per_block = (rows + kMaxSumBlocks - 1) / kMaxSumBlocks;
  if (rows_per_block < 4L * rpb) rows_per_block = 4L * rpb;
  const int nblocks = (int)((rows + rows_per_block - 1) / rows_per_block);
  act_bwd_kernel<<<nblocks, 256, 0, st>>>((const __nv_bfloat16*)dy, dy_ld, dy_coff, (const __nv_bfloat16*)raw, raw_ld,
                                          raw_coff, rows, C, 0, nullptr, 0, 0, workspace, rows_per_block, nullptr, 0, 0);
  DHD_CUDA_LAUNCH_CHECK("act_bwd(bn sums)");
  BnFinishArgs A = {};
  A.M = M; A.gamma = gamma; A.mean_in = mean; A.invstd_in = invstd;
  A.o0 = k1; A.o1 = k2; A.o2 = k3; A.dgamma = dgamma; A.dbeta = dbeta;
  bn_finish_kernel<2><<<(C + 31) / 32, 1024, 0, st>>>(workspace, nblocks, C, A);
  DHD_CUDA_LAUNCH_CHECK("bn_finish_bwd");
  return DHD_OK;
}

extern "C" int dhd_bn_bwd_coeffs(const float* sums, int C, int sums_stride, float M, const float* mean,
                                 const float* invstd, const float* gamma, float* k1, float* k2, float* k3,
                                 float* dgamma, float* dbeta, void* stream) {
  DHD_REQUIRE(sums && mean && invstd && gamma && k1 && k2 && k3 && C > 0 && sums_stride >= C && M > 0.f, "bad arguments");
  bn_bwd_coeffs_kernel<<<(C + 255) / 256, 256, 0, (cudaStream_t)stream>>>(sums, C, sums_stride, M, mean, invstd, gamma,
                                                                         k1, k2, k3, dgamma, dbeta);
  DHD_CUDA_LAUNCH_CHECK("bn_bwd_coeffs");
  return DHD_OK;
}

extern "C" int dhd_dropout(void* x, int ld, int coff, long rows, int C, float p, const long long* rng, unsigned salt,
                           void* stream) {
  DHD_REQUIRE(x && rng && rows > 0 && C > 0, "bad arguments");
  DHD_REQUIRE(C % 8 == 0 && ld % 8 == 0 && coff % 8 == 0 && ((uintptr_t)x & 15) == 0, "channels must come in groups of 8");
  DHD_REQUIRE(p >= 0.f && p < 1.f, "drop probability must be in [0, 1)");
  const long vecs = rows * (C / 8);
  const uint32_t thresh = (uint32_t)(p * 65536.f + 0.5f);
  dropout_kernel<<<(unsigned)((vecs + 255) / 256), 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)x, ld, coff, rows, C, thresh,
                                                                                 1.f / (1.f - p), rng, salt);
  DHD_CUDA_LAUNCH_CHECK("dropout");
  return DHD_OK;
}

extern "C" int dhd_gt_downsample(const float* gt, int BN, int H, int W, int ds, float lo, float interval, int nbins,
                                 int32_t* label, uint8_t* valid, void* stream) {
  DHD_REQUIRE(gt && label && BN > 0 && H > 0 && W > 0 && nbins > 0, "bad arguments");
  DHD_REQUIRE(ds > 0 && H % ds == 0 && W % ds == 0, "the map must be a whole number of ds x ds blocks");
  const long nout = (long)BN * (H / ds) * (W / ds);
  gt_downsample_kernel<<<(unsigned)((nout * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(gt, nout, H, W, ds, lo, interval,
                                                                                         nbins, label, valid);
  DHD_CUDA_LAUNCH_CHECK("gt_downsample");
  return DHD_OK;
}
