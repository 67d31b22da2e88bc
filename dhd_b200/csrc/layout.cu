// HBM-bound layout / elementwise kernels around the tensor-core convolutions:
//   pack   : fp32 NCHW (what the reference modules exchange) -> NHWC split-bf16 activation
//   mean   : per-image channel means (ASPP global pool depthnet.py:77-82, SFA squeeze mix.py:41)
//   linear : tiny fp32 row-wise Linear (+act) for the camera-aware MLP / SE / SFA fc chains
//            (depthnet.py:119-169, mix.py:20-25), M = B*N or B rows -- not tensor-core work
//   sfa_mix: the two gated blends of channel_spatial_stage.forward (mix.py:44-57)
//   dcn_im2col: bilinear deformable sampling of mmcv DeformConv2dPack (depthnet.py:466-477)
//            into a [pixel][group][tap][64] bf16 matrix consumed by the GEMM kernel
// All are pure streaming kernels: every byte is read once / written once, coalesced.
#include <cuda_bf16.h>

#include <stdlib.h>

#include "common.cuh"

namespace dhd {

// split v into `parts` bf16 values, store part p at dst[p * part_stride]
__device__ __forceinline__ void store_parts(__nv_bfloat16* dst, float v, int parts, int part_stride) {
  for (int p = 0; p < parts; ++p) {
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    dst[(size_t)p * part_stride] = h;
    v -= __bfloat162float(h);
  }
}
__device__ __forceinline__ void store_parts2(__nv_bfloat16* dst, float a, float b, int parts,
                                             int part_stride) {
  for (int p = 0; p < parts; ++p) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    *reinterpret_cast<__nv_bfloat162*>(dst + (size_t)p * part_stride) = h;
    a -= __low2float(h);
    b -= __high2float(h);
  }
}
__device__ __forceinline__ float load_parts(const __nv_bfloat16* src, int parts, int part_stride) {
  float v = 0.f;
  for (int p = parts - 1; p >= 0; --p) v += __bfloat162float(src[(size_t)p * part_stride]);
  return v;
}

// 8-channel (16-byte) variants: v[8] <-> `parts` uint4 of bf16
__device__ __forceinline__ void load_parts8(const __nv_bfloat16* src, int parts, int part_stride, float (&v)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = 0.f;
  for (int p = parts - 1; p >= 0; --p) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(src + (size_t)p * part_stride));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[2 * j] += __low2float(h[j]);
      v[2 * j + 1] += __high2float(h[j]);
    }
  }
}
__device__ __forceinline__ void store_parts8(__nv_bfloat16* dst, float (&v)[8], int parts, int part_stride) {
  for (int p = 0; p < parts; ++p) {
    uint4 q;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
      v[2 * j] -= __low2float(h[j]);
      v[2 * j + 1] -= __high2float(h[j]);
    }
    *reinterpret_cast<uint4*>(dst + (size_t)p * part_stride) = q;
  }
}

// ---- pack: one block = 64 channels x 32 pixels of one image -------------------------------
__global__ void __launch_bounds__(256)
pack_nchw_to_nhwc_kernel(const float* __restrict__ in, int C, int HW, __nv_bfloat16* __restrict__ out,
                         int ld, int coff, int part_stride, int parts) {
  __shared__ float tile[64][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 64, n = blockIdx.z;
  for (int c = ty; c < 64; c += 8) {
    float v = 0.f;
    if (c0 + c < C && p0 + tx < HW) v = in[((size_t)n * C + c0 + c) * HW + p0 + tx];
    tile[c][tx] = v;
  }
  __syncthreads();
  for (int p = ty; p < 32; p += 8) {
    if (p0 + p >= HW) continue;
    const int c = 2 * tx;
    if (c0 + c >= C) continue;   // C is padded to a multiple of 64 by the caller's zero fill
    __nv_bfloat16* dst = out + ((size_t)n * HW + p0 + p) * ld + coff + c0 + c;
    if (c0 + c + 1 < C) store_parts2(dst, tile[c][p], tile[c + 1][p], parts, part_stride);
    else store_parts(dst, tile[c][p], parts, part_stride);
  }
}

// ---- split: fp32 rows [R][C] (an NHWC tensor) -> split-bf16 rows, channel pairs per thread ----
__global__ void __launch_bounds__(256)
split_nhwc_kernel(const float* __restrict__ in, long R, int C, __nv_bfloat16* __restrict__ out, int ld,
                  int coff, int part_stride, int parts) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int cp = C / 2;
  if (i >= R * cp) return;
  const long r = i / cp;
  const int c = (int)(i % cp) * 2;
  const float2 v = *reinterpret_cast<const float2*>(in + r * C + c);
  store_parts2(out + r * ld + coff + c, v.x, v.y, parts, part_stride);
}

// 8 channels per thread; block = (C/8 channel groups) x (256/(C/8) pixel rows), `pb` pixels of ONE image
// per block; optionally accumulates the per-image channel mean of the rows it touches (SFA squeeze,
// mix.py:41) so the tensor is read once.
__global__ void __launch_bounds__(256)
split_nhwc8_kernel(const float* __restrict__ in, int HW, int C, int pb, __nv_bfloat16* __restrict__ out, int ld,
                   int coff, int part_stride, int parts, float* __restrict__ mean_out, float inv) {
  __shared__ float red[256][9];
  const int cg = C / 8, rows = 256 / cg;
  const int gi = threadIdx.x % cg, ri = threadIdx.x / cg;
  const int n = blockIdx.y, p0 = blockIdx.x * pb, p1 = min(HW, p0 + pb);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (ri < rows) {
    for (int p = p0 + ri; p < p1; p += rows) {
      const size_t r = (size_t)n * HW + p;
      const float4* src = reinterpret_cast<const float4*>(in + r * C + gi * 8);
      const float4 a = __ldg(src), b = __ldg(src + 1);
      float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
      store_parts8(out + r * ld + coff + gi * 8, v, parts, part_stride);
    }
  }
  if (mean_out == nullptr) return;
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.x][j] = acc[j];
  __syncthreads();
  if (ri == 0) {
    for (int k = 1; k < rows; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += red[threadIdx.x + k * cg][j];
    // block partial, summed over the blocks in fixed order by mean_finish_kernel (deterministic: no atomics)
    float* dst = mean_out + ((size_t)n * gridDim.x + blockIdx.x) * C + gi * 8;
    reinterpret_cast<float4*>(dst)[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    reinterpret_cast<float4*>(dst)[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}

// per-image channel mean of an Act, 8 channels per thread: every block writes its partial sums [n][block][C]
// (four independent 16-byte loads in flight per thread), mean_finish_kernel adds them in block order -- bit-identical
// from run to run (the first version combined the blocks with fp32 atomics: 2.0 TB/s and run-to-run ulp noise that
// reached the height head's argmax through the ASPP's global branch)
__global__ void __launch_bounds__(256)
mean_hw8_kernel(const __nv_bfloat16* __restrict__ in, int ld, int coff, int part_stride, int parts, int C,
                int HW, int pb, float* __restrict__ partial) {
  __shared__ float red[256][9];
  const int cg = C / 8, rows = 256 / cg;
  const int gi = threadIdx.x % cg, ri = threadIdx.x / cg;
  const int n = blockIdx.y, p0 = blockIdx.x * pb, p1 = min(HW, p0 + pb);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (ri < rows) {
    const __nv_bfloat16* base = in + (size_t)n * HW * ld + coff + gi * 8;
    int p = p0 + ri;
    if (parts == 1) {
      for (; p + 3 * rows < p1; p += 4 * rows) {
        uint4 q[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) q[u] = __ldg(reinterpret_cast<const uint4*>(base + (size_t)(p + u * rows) * ld));
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&q[u]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            acc[2 * j] += __low2float(h[j]);
            acc[2 * j + 1] += __high2float(h[j]);
          }
        }
      }
    }
    for (; p < p1; p += rows) {
      float v[8];
      load_parts8(base + (size_t)p * ld, parts, part_stride, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) red[threadIdx.x][j] = acc[j];
  __syncthreads();
  if (ri == 0) {
    for (int k = 1; k < rows; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += red[threadIdx.x + k * cg][j];
    float* dst = partial + ((size_t)n * gridDim.x + blockIdx.x) * C + gi * 8;
    reinterpret_cast<float4*>(dst)[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    reinterpret_cast<float4*>(dst)[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  }
}

// out[n][c] = inv * sum_b partial[n][b][c] in a FIXED order: 8 slices per (n, c) (slice s adds the blocks b = s, s + 8,
// ... ascending), then the 8 slice sums ascending -- one block = 32 channels x 8 slices, coalesced over the channels
__global__ void __launch_bounds__(256)
mean_finish_kernel(const float* __restrict__ partial, int N, int nblk, int C, float* __restrict__ out, float inv) {
  __shared__ float red[8][33];
  const int cl = threadIdx.x & 31, sl = threadIdx.x >> 5;
  const int cblocks = (C + 31) / 32;
  const int n = blockIdx.x / cblocks, c = (blockIdx.x % cblocks) * 32 + cl;
  float s = 0.f;
  if (c < C) {
    const float* src = partial + (size_t)n * nblk * C + c;
    for (int b = sl; b < nblk; b += 8) s += __ldg(src + (size_t)b * C);
  }
  red[sl][cl] = s;
  __syncthreads();
  if (sl == 0 && c < C) {
    float t = red[0][cl];
#pragma unroll
    for (int k = 1; k < 8; ++k) t += red[k][cl];
    out[(size_t)n * C + c] = t * inv;
  }
}

// out = x * gate[n][c] (SELayer's multiply, depthnet.py:169, when one feature map feeds two
// differently gated branches as in DepthNet.forward :388-396): fp32 NHWC in, split-bf16 (+ fp32) out
__global__ void __launch_bounds__(256)
gate_channels_kernel(const float* __restrict__ x, int C, long npix_total, int HW, const float* __restrict__ gate,
                     __nv_bfloat16* __restrict__ out, int o_ld, int o_coff, int o_ps, int o_parts,
                     float* __restrict__ out32) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int cg = C / 8;
  if (i >= npix_total * cg) return;
  int c;
  const long pix = fast_div(i, cg, &c);
  c *= 8;
  const int n = (int)fast_div(pix, HW);
  const float4* xp = reinterpret_cast<const float4*>(x + (size_t)pix * C + c);
  const float4* gp = reinterpret_cast<const float4*>(gate + (size_t)n * C + c);
  const float4 a = __ldg(xp), b = __ldg(xp + 1), ga = __ldg(gp), gb = __ldg(gp + 1);
  float v[8] = {a.x * ga.x, a.y * ga.y, a.z * ga.z, a.w * ga.w, b.x * gb.x, b.y * gb.y, b.z * gb.z, b.w * gb.w};
  if (out32 != nullptr) {
    float4* o = reinterpret_cast<float4*>(out32 + (size_t)pix * C + c);
    o[0] = make_float4(v[0], v[1], v[2], v[3]);
    o[1] = make_float4(v[4], v[5], v[6], v[7]);
  }
  store_parts8(out + (size_t)pix * o_ld + o_coff + c, v, o_parts, o_ps);
}

// SFA blends, 8 channels per thread (same arithmetic as sfa_mix_kernel)
__global__ void __launch_bounds__(256)
sfa_mix8_kernel(const __nv_bfloat16* __restrict__ x, int x_ld, int x_coff, int x_ps, int x_parts, int C,
                long npix_total, int HW, const float* __restrict__ a1, const float* __restrict__ a2,
                __nv_bfloat16* __restrict__ out, int o_ld, int o_coff, int o_ps, int o_parts) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int cg = C / 8;
  if (i >= npix_total * cg) return;
  int c;
  const long pix = fast_div(i, cg, &c);
  c *= 8;
  const int n = (int)fast_div(pix, HW);
  float bev[8], vox[8], r[8];
  load_parts8(x + (size_t)pix * x_ld + x_coff + c, x_parts, x_ps, bev);
  load_parts8(x + (size_t)pix * x_ld + x_coff + C + c, x_parts, x_ps, vox);
  const float4* g1p = reinterpret_cast<const float4*>(a1 + (size_t)n * C + c);
  const float4 ga = __ldg(g1p), gb = __ldg(g1p + 1);
  const float g1[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
  float g2[8];
  if (a2 != nullptr) {
    const float4* g2p = reinterpret_cast<const float4*>(a2 + (size_t)pix * C + c);
    const float4 qa = __ldg(g2p), qb = __ldg(g2p + 1);
    g2[0] = qa.x; g2[1] = qa.y; g2[2] = qa.z; g2[3] = qa.w; g2[4] = qb.x; g2[5] = qb.y; g2[6] = qb.z; g2[7] = qb.w;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float b1 = __fmul_rn(g1[j], bev[j]);
    const float v1 = __fmul_rn(__fsub_rn(1.f, g1[j]), vox[j]);
    r[j] = a2 == nullptr ? __fadd_rn(b1, v1)
                         : __fadd_rn(__fmul_rn(g2[j], b1), __fmul_rn(__fsub_rn(1.f, g2[j]), v1));
  }
  store_parts8(out + (size_t)pix * o_ld + o_coff + c, r, o_parts, o_ps);
}

// bf16 speed mode of the second SFA blend: the spatial gate a2 arrives as a bf16 activation (written by the gate
// convolution's 128-byte-row TMA stores) instead of an fp32 tensor -- 16 B of gate per 8 channels instead of 32
__global__ void __launch_bounds__(256)
sfa_blend_b16_kernel(const __nv_bfloat16* __restrict__ x, int x_ld, int x_coff, int C, long npix_total, int HW,
                     const float* __restrict__ a1, const __nv_bfloat16* __restrict__ a2, int a2_ld, int a2_coff,
                     __nv_bfloat16* __restrict__ out, int o_ld, int o_coff) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int cg = C / 8;
  if (i >= npix_total * cg) return;
  int c;
  const long pix = fast_div(i, cg, &c);
  c *= 8;
  const int n = (int)fast_div(pix, HW);
  const uint4 qb = ld_nc_u4(x + (size_t)pix * x_ld + x_coff + c);
  const uint4 qv = ld_nc_u4(x + (size_t)pix * x_ld + x_coff + C + c);
  const uint4 qg = ld_nc_u4(a2 + (size_t)pix * a2_ld + a2_coff + c);
  const float4* g1p = reinterpret_cast<const float4*>(a1 + (size_t)n * C + c);
  const float4 ga = __ldg(g1p), gb = __ldg(g1p + 1);
  const float g1[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
  const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&qb);
  const __nv_bfloat162* hv = reinterpret_cast<const __nv_bfloat162*>(&qv);
  const __nv_bfloat162* hg = reinterpret_cast<const __nv_bfloat162*>(&qg);
  uint4 qo;
  __nv_bfloat162* ho = reinterpret_cast<__nv_bfloat162*>(&qo);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float r[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float bev = e == 0 ? __low2float(hb[j]) : __high2float(hb[j]);
      const float vox = e == 0 ? __low2float(hv[j]) : __high2float(hv[j]);
      const float g2 = e == 0 ? __low2float(hg[j]) : __high2float(hg[j]);
      const float b1 = __fmul_rn(g1[2 * j + e], bev);
      const float v1 = __fmul_rn(__fsub_rn(1.f, g1[2 * j + e]), vox);
      r[e] = __fadd_rn(__fmul_rn(g2, b1), __fmul_rn(__fsub_rn(1.f, g2), v1));
    }
    ho[j] = __floats2bfloat162_rn(r[0], r[1]);
  }
  *reinterpret_cast<uint4*>(out + (size_t)pix * o_ld + o_coff + c) = qo;
}

// SFA's channel gate folded into the weights of the 1x1 convolution that consumes the gated blend (mix.py:41-50):
// wq[n][co][c] = w[co][c] * a1[n][c], wq[n][co][C + c] = w[co][c] * (1 - a1[n][c]), bf16
__global__ void __launch_bounds__(256)
sfa_fold_gate_kernel(const float* __restrict__ w, const float* __restrict__ a1, int N, int Cout, int C,
                     __nv_bfloat16* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * Cout * C) return;
  const int c = i % C, co = (i / C) % Cout, n = i / (C * Cout);
  const float g = __ldg(a1 + (size_t)n * C + c), wv = __ldg(w + (size_t)co * C + c);
  __nv_bfloat16* o = out + ((size_t)n * Cout + co) * 2 * C;
  o[c] = __float2bfloat16_rn(__fmul_rn(wv, g));
  o[C + c] = __float2bfloat16_rn(__fmul_rn(wv, __fsub_rn(1.f, g)));
}

// deformable im2col, one warp per (pixel, tap), 8 channels per lane (C multiple of 256 per pass)
__global__ void __launch_bounds__(256)
dcn_im2col8_kernel(const __nv_bfloat16* __restrict__ x, int x_ld, int x_coff, int x_ps, int x_parts, int C,
                   int N, int H, int W, const float* __restrict__ offset, int off_ld, int ksize, int pad,
                   int dil, int groups, __nv_bfloat16* __restrict__ out, int o_ld, int o_ps, int o_parts) {
  const int lane = threadIdx.x & 31;
  const int gw = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);     // N*H*W*taps < 2^31 (host check):
  const int taps = ksize * ksize;                                                // 32-bit index arithmetic throughout
  if (gw >= N * H * W * taps) return;
  const int t = gw % taps;
  const int pix = gw / taps;
  const int wx = pix % W, hy = (pix / W) % H, n = pix / (W * H);
  const float dy = __ldg(offset + (size_t)pix * off_ld + 2 * t), dx = __ldg(offset + (size_t)pix * off_ld + 2 * t + 1);
  const float sy = (float)(hy - pad + (t / ksize) * dil) + dy;
  const float sx = (float)(wx - pad + (t % ksize) * dil) + dx;
  const int cg = C / groups;
  float w00 = 0.f, w01 = 0.f, w10 = 0.f, w11 = 0.f;
  int y0 = 0, x0 = 0, y1 = 0, x1 = 0;
  const bool inside = sy > -1.f && sx > -1.f && sy < (float)H && sx < (float)W;
  if (inside) {
    y0 = (int)floorf(sy);
    x0 = (int)floorf(sx);
    y1 = y0 + 1;
    x1 = x0 + 1;
    const float ly = sy - (float)y0, lx = sx - (float)x0, hy_ = 1.f - ly, hx_ = 1.f - lx;
    w00 = hy_ * hx_; w01 = hy_ * lx; w10 = ly * hx_; w11 = ly * lx;
  }
  for (int c = 8 * lane; c < C; c += 256) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = 0.f;
    if (inside) {
      const __nv_bfloat16* base = x + (size_t)n * H * W * x_ld + x_coff + c;
      float q[8];
      if (y0 >= 0 && x0 >= 0) {
        load_parts8(base + ((size_t)y0 * W + x0) * x_ld, x_parts, x_ps, q);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = w00 * q[j];
      }
      if (y0 >= 0 && x1 <= W - 1) {
        load_parts8(base + ((size_t)y0 * W + x1) * x_ld, x_parts, x_ps, q);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += w01 * q[j];
      }
      if (y1 <= H - 1 && x0 >= 0) {
        load_parts8(base + ((size_t)y1 * W + x0) * x_ld, x_parts, x_ps, q);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += w10 * q[j];
      }
      if (y1 <= H - 1 && x1 <= W - 1) {
        load_parts8(base + ((size_t)y1 * W + x1) * x_ld, x_parts, x_ps, q);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += w11 * q[j];
      }
    }
    const int g = c / cg, cl = c % cg;
    store_parts8(out + (size_t)pix * o_ld + (size_t)g * taps * cg + (size_t)t * cg + cl, v, o_parts, o_ps);
  }
}

// Same result, one warp per PIXEL: lane t < taps computes tap t's sampling position, bilinear weights and corner
// offsets once; the warp then walks the taps with the parameters handed round by shuffle and the lanes over the
// channels.  The per-(pixel, tap) kernel above spends ~400 instructions of index arithmetic per 16 bytes of payload;
// here that arithmetic is paid once per pixel.
__global__ void __launch_bounds__(256)
dcn_im2col_px_kernel(const __nv_bfloat16* __restrict__ x, int x_ld, int x_coff, int x_ps, int x_parts, int C,
                     int N, int H, int W, const float* __restrict__ offset, int off_ld, int ksize, int pad,
                     int dil, int groups, __nv_bfloat16* __restrict__ out, int o_ld, int o_ps, int o_parts) {
  const int lane = threadIdx.x & 31;
  const int pix = (int)(((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (pix >= N * H * W) return;
  const int taps = ksize * ksize;
  const int wx = pix % W, hy = (pix / W) % H, n = pix / (W * H);
  float w00 = 0.f, w01 = 0.f, w10 = 0.f, w11 = 0.f;
  int o00 = -1, o01 = -1, o10 = -1, o11 = -1;           // pixel index of each corner inside the image, -1 = outside
  if (lane < taps) {
    const int t = lane;
    const float dy = __ldg(offset + (size_t)pix * off_ld + 2 * t), dx = __ldg(offset + (size_t)pix * off_ld + 2 * t + 1);
    const float sy = (float)(hy - pad + (t / ksize) * dil) + dy;
    const float sx = (float)(wx - pad + (t % ksize) * dil) + dx;
    if (sy > -1.f && sx > -1.f && sy < (float)H && sx < (float)W) {
      const int y0 = (int)floorf(sy), x0 = (int)floorf(sx), y1 = y0 + 1, x1 = x0 + 1;
      const float ly = sy - (float)y0, lx = sx - (float)x0, hy_ = 1.f - ly, hx_ = 1.f - lx;
      w00 = hy_ * hx_; w01 = hy_ * lx; w10 = ly * hx_; w11 = ly * lx;
      if (y0 >= 0 && x0 >= 0) o00 = y0 * W + x0;
      if (y0 >= 0 && x1 <= W - 1) o01 = y0 * W + x1;
      if (y1 <= H - 1 && x0 >= 0) o10 = y1 * W + x0;
      if (y1 <= H - 1 && x1 <= W - 1) o11 = y1 * W + x1;
    }
  }
  const int cg = C / groups;
  const __nv_bfloat16* img = x + (size_t)n * H * W * x_ld + x_coff;
  __nv_bfloat16* orow = out + (size_t)pix * o_ld;
  for (int t = 0; t < taps; ++t) {
    const float a00 = __shfl_sync(kFull, w00, t), a01 = __shfl_sync(kFull, w01, t);
    const float a10 = __shfl_sync(kFull, w10, t), a11 = __shfl_sync(kFull, w11, t);
    const int p00 = __shfl_sync(kFull, o00, t), p01 = __shfl_sync(kFull, o01, t);
    const int p10 = __shfl_sync(kFull, o10, t), p11 = __shfl_sync(kFull, o11, t);
    for (int c = 8 * lane; c < C; c += 256) {
      float v[8], q[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;
      if (p00 >= 0) {
        load_parts8(img + (size_t)p00 * x_ld + c, x_parts, x_ps, q);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = a00 * q[j];
      }
      if (p01 >= 0) {
        load_parts8(img + (size_t)p01 * x_ld + c, x_parts, x_ps, q);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += a01 * q[j];
      }
      if (p10 >= 0) {
        load_parts8(img + (size_t)p10 * x_ld + c, x_parts, x_ps, q);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += a10 * q[j];
      }
      if (p11 >= 0) {
        load_parts8(img + (size_t)p11 * x_ld + c, x_parts, x_ps, q);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += a11 * q[j];
      }
      const int g = c / cg, cl = c - g * cg;
      store_parts8(orow + (size_t)g * taps * cg + (size_t)t * cg + cl, v, o_parts, o_ps);
    }
  }
}

// ---- encoder helpers (backbones/unet.py:65-74 MaxPool2d(2); necks/lss_fpn.py:27-28, 41-42 bilinear
// Upsample(align_corners=True)): bf16 NHWC in, bf16 NHWC out (possibly a channel slice of a wider buffer), 8 channels
// per thread.
__global__ void __launch_bounds__(256)
maxpool2_kernel(const __nv_bfloat16* __restrict__ in, int in_ld, int in_coff, int in_ps, int N, int H, int W, int C,
                __nv_bfloat16* __restrict__ out, int o_ld, int o_coff, int o_ps, int parts) {
  const int oH = H / 2, oW = W / 2, cg = C / 8;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)N * oH * oW * cg) return;
  int c, ox, oy;
  long p = fast_div(i, cg, &c);
  c *= 8;
  p = fast_div(p, oW, &ox);
  const int n = (int)fast_div(p, oH, &oy);
  float m[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      float v[8];
      load_parts8(in + (((size_t)n * H + 2 * oy + dy) * W + 2 * ox + dx) * in_ld + in_coff + c, parts, in_ps, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], v[j]);
    }
  store_parts8(out + (((size_t)n * oH + oy) * oW + ox) * o_ld + o_coff + c, m, parts, o_ps);
}

// ---- image backbone helpers (mmdet ResNet stem / CustomFPN top-down path) ---------------------------------------
// im2col of the ResNet stem (conv1: 7x7, stride 2, pad 3 on 3-channel images, torchvision / mmdet `ResNet.conv1`): the
// K axis is (ky, kx, c) = ksize*ksize*Cin values zero-padded to Kpad (a multiple of 64), so the convolution becomes a
// 1x1 tcgen05 GEMM over [N*oH*oW][Kpad].  img fp32 NCHW; 8 consecutive K values per thread.
// The (ky, kx, c) decomposition of every K index comes from a table in kernel-parameter space (one entry per K value:
// dy | dx << 8 | c << 16) instead of two divisions per element; the zero padding of the row is written here too (no
// separate memset of the 415 MB buffer).  First version: 601 us at 24 x 256x704 images, of which ~500 were the divisions.
struct StemTable {
  uint32_t e[512];
};
// A block owns kStemPx consecutive output pixels of one output row: the image patch they read (ksize rows x
// (kStemPx-1)*stride + ksize columns x Cin channels, zero outside the image) is staged in shared memory with coalesced
// row reads, then thread = (pixel, group of 8 K values) assembles its 16 bytes from shared memory and the block's
// output rows leave as contiguous 16-byte stores.  (One thread gathering its 8 values from global memory: 385 us; the
// write of the 415 MB alone is ~70 us.)
constexpr int kStemPx = 64;
__global__ void __launch_bounds__(256)
stem_im2col_kernel(const __grid_constant__ StemTable T, const float* __restrict__ img, int N, int Cin, int H, int W,
                   int ksize, int stride, int pad, int oH, int oW, int K, int groups, __nv_bfloat16* __restrict__ out,
                   int o_ld, int o_ps, int parts) {
  extern __shared__ float patch[];                   // [Cin][ksize][pw]
  __shared__ uint32_t tab[512];
  for (int k = threadIdx.x; k < K; k += blockDim.x) tab[k] = T.e[k];
  const int tiles_x = (oW + kStemPx - 1) / kStemPx;
  int bx = blockIdx.x;
  const int tx = bx % tiles_x;
  bx /= tiles_x;
  const int oy = bx % oH, n = bx / oH;
  const int ox0 = tx * kStemPx;
  const int pw = (kStemPx - 1) * stride + ksize;
  const int x0 = ox0 * stride - pad, y0 = oy * stride - pad;
  const float* base = img + (size_t)n * Cin * H * W;
  // staging: every thread issues all of its (<= 12) loads before the first shared-memory store -- one row per trip
  // with a dependent load -> store made the block's life a chain of 21 global-memory latencies (ncu: 415 us)
  {
    const unsigned total = (unsigned)(Cin * ksize * pw);
    float stg[12];
#pragma unroll
    for (int m = 0; m < 12; ++m) {
      const unsigned i = threadIdx.x + 256u * m;
      float val = 0.f;
      if (i < total) {
        const unsigned r = i / (unsigned)pw, px = i - r * (unsigned)pw;
        const unsigned c = r / (unsigned)ksize, ky = r - c * (unsigned)ksize;
        const int iy = y0 + (int)ky, ix = x0 + (int)px;
        if ((unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W) val = __ldg(base + ((size_t)c * H + iy) * W + ix);
      }
      stg[m] = val;
    }
#pragma unroll
    for (int m = 0; m < 12; ++m) {
      const unsigned i = threadIdx.x + 256u * m;
      if (i < total) patch[i] = stg[m];
    }
    for (unsigned i = threadIdx.x + 256u * 12; i < total; i += 256u) {      // larger patches than the 7x7x3 stem's
      const unsigned r = i / (unsigned)pw, px = i - r * (unsigned)pw;
      const unsigned c = r / (unsigned)ksize, ky = r - c * (unsigned)ksize;
      const int iy = y0 + (int)ky, ix = x0 + (int)px;
      patch[i] = ((unsigned)iy < (unsigned)H && (unsigned)ix < (unsigned)W) ? __ldg(base + ((size_t)c * H + iy) * W + ix) : 0.f;
    }
  }
  __syncthreads();
  const int npx = min(kStemPx, oW - ox0);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int g = lane; g < groups; g += 32)            // lane = group of 8 K values (24 groups for the 7x7x3 stem)
  for (int lp = wid; lp < npx; lp += nw) {
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = g * 8 + j;
      float val = 0.f;
      if (k < K) {
        const uint32_t e = tab[k];
        val = patch[((int)(e >> 16) * ksize + (int)(e & 0xffu)) * pw + lp * stride + (int)((e >> 8) & 0xffu)];
      }
      v[j] = val;
    }
    store_parts8(out + (((size_t)n * oH + oy) * oW + ox0 + lp) * o_ld + g * 8, v, parts, o_ps);
  }
}

// MaxPool2d(kernel 3, stride 2, padding 1) (ResNet.maxpool): bf16 NHWC (split parts summed before the max)
__global__ void __launch_bounds__(256)
maxpool3s2_kernel(const __nv_bfloat16* __restrict__ in, int in_ld, int in_coff, int in_ps, int N, int H, int W, int C,
                  int oH, int oW, __nv_bfloat16* __restrict__ out, int o_ld, int o_coff, int o_ps, int parts) {
  const int cg = C / 8;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)N * oH * oW * cg) return;
  int c, ox, oy;
  long p = fast_div(i, cg, &c);
  c *= 8;
  p = fast_div(p, oW, &ox);
  const int n = (int)fast_div(p, oH, &oy);
  float m[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      const int iy = 2 * oy + dy, ix = 2 * ox + dx;
      if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
      float v[8];
      load_parts8(in + (((size_t)n * H + iy) * W + ix) * in_ld + in_coff + c, parts, in_ps, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) m[j] = fmaxf(m[j], v[j]);
    }
  store_parts8(out + (((size_t)n * oH + oy) * oW + ox) * o_ld + o_coff + c, m, parts, o_ps);
}

// MaxPool2d(kernel 3, stride 2, padding 1) backward (ResNet.maxpool under training): one thread per INPUT pixel and
// 8 channels gathers from the <= 4 windows that cover it -- window o spans input rows 2o-1 .. 2o+1, so an even row
// belongs to one window row and an odd row to two.  A window's gradient goes to its FIRST maximum in (ky, kx) scan
// order (torch's rule: `val > maxval` replaces), recomputed from x; no atomics, no index tensor, deterministic.
// relu_mask: x is the ReLU output that fed the pool -- the gradient is also multiplied by [x > 0] (the stem's ReLU
// backward fused into this pass).
__global__ void __launch_bounds__(256)
maxpool3s2_bwd_ref_kernel(const __nv_bfloat16* __restrict__ x, int x_ld, int x_coff, const __nv_bfloat16* __restrict__ dy,
                      int dy_ld, int dy_coff, int N, int H, int W, int C, int oH, int oW, __nv_bfloat16* __restrict__ dx,
                      int dx_ld, int dx_coff, int relu_mask) {
  const int cg = C / 8;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)N * H * W * cg) return;
  int c, xx, yy;
  long p = fast_div(i, cg, &c);
  c *= 8;
  p = fast_div(p, W, &xx);
  const int n = (int)fast_div(p, H, &yy);
  float r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = 0.f;
  const int oy1 = min((yy + 1) >> 1, oH - 1), ox1 = min((xx + 1) >> 1, oW - 1);
  for (int oy = yy >> 1; oy <= oy1; ++oy)
    for (int ox = xx >> 1; ox <= ox1; ++ox) {
      const int me = (yy - (2 * oy - 1)) * 3 + (xx - (2 * ox - 1));      // this pixel's place in the window's scan
      float m[8];
      int best[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        m[j] = -INFINITY;
        best[j] = -1;
      }
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const int iy = 2 * oy - 1 + k / 3, ix = 2 * ox - 1 + k % 3;
        if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
        float v[8];
        load_parts8(x + (((size_t)n * H + iy) * W + ix) * x_ld + x_coff + c, 1, 0, v);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (v[j] > m[j] || best[j] < 0) {
            m[j] = v[j];
            best[j] = k;
          }
      }
      float g[8];
      load_parts8(dy + (((size_t)n * oH + oy) * oW + ox) * dy_ld + dy_coff + c, 1, 0, g);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (best[j] == me) r[j] += g[j];
    }
  if (relu_mask) {
    float v[8];
    load_parts8(x + (((size_t)n * H + yy) * W + xx) * x_ld + x_coff + c, 1, 0, v);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (!(v[j] > 0.f)) r[j] = 0.f;
  }
  store_parts8(dx + (((size_t)n * H + yy) * W + xx) * dx_ld + dx_coff + c, r, 1, 0);
}

// The same result with packed bf16x2 compares and no arg-max bookkeeping (ncu of the kernel above, profiles/
// r02_maxpool3s2_bwd_full.txt: instruction-bound, 559 M warp instructions, the bf16 -> fp32 unpacking of 36 taps per
// thread on top): this pixel (tap k of window o) is the window's FIRST maximum iff it is > every earlier tap and >= every
// later one, so each of the 8 other taps costs one 16-byte load and four `set.{gt,ge}.bf16x2` + AND per 8 channels.
__global__ void __launch_bounds__(256)
maxpool3s2_bwd_kernel(const __nv_bfloat16* __restrict__ x, int x_ld, int x_coff, const __nv_bfloat16* __restrict__ dy,
                      int dy_ld, int dy_coff, int N, int H, int W, int C, int oH, int oW, __nv_bfloat16* __restrict__ dx,
                      int dx_ld, int dx_coff, int relu_mask) {
  const int cg = C / 8;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)N * H * W * cg) return;
  int c, xx, yy;
  long p = fast_div(i, cg, &c);
  c *= 8;
  p = fast_div(p, W, &xx);
  const int n = (int)fast_div(p, H, &yy);
  const __nv_bfloat16* img = x + (size_t)n * H * W * x_ld + x_coff + c;
  const uint4 me4 = __ldg(reinterpret_cast<const uint4*>(img + ((size_t)yy * W + xx) * x_ld));
  const __nv_bfloat162* me = reinterpret_cast<const __nv_bfloat162*>(&me4);
  float r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = 0.f;
  const int oy1 = min((yy + 1) >> 1, oH - 1), ox1 = min((xx + 1) >> 1, oW - 1);
  for (int oy = yy >> 1; oy <= oy1; ++oy)
    for (int ox = xx >> 1; ox <= ox1; ++ox) {
      const int k = (yy - (2 * oy - 1)) * 3 + (xx - (2 * ox - 1));       // this pixel's place in the window's scan
      unsigned ok[4] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu};
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        const int iy = 2 * oy - 1 + j / 3, ix = 2 * ox - 1 + j % 3;
        if (j == k || iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
        const uint4 q4 = __ldg(reinterpret_cast<const uint4*>(img + ((size_t)iy * W + ix) * x_ld));
        const __nv_bfloat162* q = reinterpret_cast<const __nv_bfloat162*>(&q4);
#pragma unroll
        for (int h = 0; h < 4; ++h) ok[h] &= j < k ? __hgt2_mask(me[h], q[h]) : __hge2_mask(me[h], q[h]);
      }
      float g[8];
      load_parts8(dy + (((size_t)n * oH + oy) * oW + ox) * dy_ld + dy_coff + c, 1, 0, g);
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        if (ok[h] & 0xffffu) r[2 * h] += g[2 * h];
        if (ok[h] >> 16) r[2 * h + 1] += g[2 * h + 1];
      }
    }
  if (relu_mask) {
    const __nv_bfloat162 zero = __floats2bfloat162_rn(0.f, 0.f);
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const unsigned pos = __hgt2_mask(me[h], zero);
      if (!(pos & 0xffffu)) r[2 * h] = 0.f;
      if (!(pos >> 16)) r[2 * h + 1] = 0.f;
    }
  }
  store_parts8(dx + (((size_t)n * H + yy) * W + xx) * dx_ld + dx_coff + c, r, 1, 0);
}

// io += F.interpolate(lo, size=(H, W), mode='nearest') (CustomFPN's top-down path, necks/fpn.py:166-176):
// src index = floor(dst * in / out), torch's nearest rule
__global__ void __launch_bounds__(256)
upsample_nearest_add_kernel(const __nv_bfloat16* __restrict__ lo, int l_ld, int l_coff, int l_ps, int h, int w,
                            __nv_bfloat16* __restrict__ io, int o_ld, int o_coff, int o_ps, int N, int H, int W, int C,
                            int parts) {
  const int cg = C / 8;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)N * H * W * cg) return;
  int c, ox, oy;
  long p = fast_div(i, cg, &c);
  c *= 8;
  p = fast_div(p, W, &ox);
  const int n = (int)fast_div(p, H, &oy);
  const int sy = min((int)floorf((float)oy * ((float)h / (float)H)), h - 1);
  const int sx = min((int)floorf((float)ox * ((float)w / (float)W)), w - 1);
  float a[8], b[8];
  __nv_bfloat16* dst = io + (((size_t)n * H + oy) * W + ox) * o_ld + o_coff + c;
  load_parts8(dst, parts, o_ps, a);
  load_parts8(lo + (((size_t)n * h + sy) * w + sx) * l_ld + l_coff + c, parts, l_ps, b);
#pragma unroll
  for (int j = 0; j < 8; ++j) a[j] += b[j];
  store_parts8(dst, a, parts, o_ps);
}

__global__ void __launch_bounds__(256)
upsample_bilinear_kernel(const __nv_bfloat16* __restrict__ in, int in_ld, int in_coff, int in_ps, int N, int H, int W,
                         int C, int oH, int oW, __nv_bfloat16* __restrict__ out, int o_ld, int o_coff, int o_ps,
                         int parts) {
  const int cg = C / 8;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)N * oH * oW * cg) return;
  int c, ox, oy;
  long p = fast_div(i, cg, &c);
  c *= 8;
  p = fast_div(p, oW, &ox);
  const int n = (int)fast_div(p, oH, &oy);
  // align_corners=True: src = dst * (in - 1) / (out - 1), torch's area_pixel_compute_source_index
  const float sy = oH > 1 ? (float)oy * ((float)(H - 1) / (float)(oH - 1)) : 0.f;
  const float sx = oW > 1 ? (float)ox * ((float)(W - 1) / (float)(oW - 1)) : 0.f;
  const int y0 = (int)sy, x0 = (int)sx;
  const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
  const float ly = sy - (float)y0, lx = sx - (float)x0;
  float a[8], b[8], cc[8], d[8], r[8];
  const __nv_bfloat16* base = in + (size_t)n * H * W * in_ld + in_coff + c;
  load_parts8(base + ((size_t)y0 * W + x0) * in_ld, parts, in_ps, a);
  load_parts8(base + ((size_t)y0 * W + x1) * in_ld, parts, in_ps, b);
  load_parts8(base + ((size_t)y1 * W + x0) * in_ld, parts, in_ps, cc);
  load_parts8(base + ((size_t)y1 * W + x1) * in_ld, parts, in_ps, d);
#pragma unroll
  for (int j = 0; j < 8; ++j)
    r[j] = (1.f - ly) * ((1.f - lx) * a[j] + lx * b[j]) + ly * ((1.f - lx) * cc[j] + lx * d[j]);
  store_parts8(out + (((size_t)n * oH + oy) * oW + ox) * o_ld + o_coff + c, r, parts, o_ps);
}

// MaxPool2d(2) backward: dx[2y+i, 2x+j] = dy[y, x] for the FIRST maximum of the window in (i, j) scan order (torch),
// 0 elsewhere -- including the last row / column of an odd-sized map, which no window covers.
__global__ void __launch_bounds__(256)
maxpool2_bwd_kernel(const __nv_bfloat16* __restrict__ x, int x_ld, int x_coff, const __nv_bfloat16* __restrict__ dy,
                    int dy_ld, int dy_coff, int N, int H, int W, int C, __nv_bfloat16* __restrict__ dx, int dx_ld,
                    int dx_coff) {
  const int oH = H / 2, oW = W / 2, cg = C / 8;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)N * H * W * cg) return;
  int c, xx, yy;
  long p = fast_div(i, cg, &c);
  c *= 8;
  p = fast_div(p, W, &xx);
  const int n = (int)fast_div(p, H, &yy);
  float r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = 0.f;
  const int oy = yy / 2, ox = xx / 2;
  if (oy < oH && ox < oW) {
    float g[8], v[4][8];
    load_parts8(dy + (((size_t)n * oH + oy) * oW + ox) * dy_ld + dy_coff + c, 1, 0, g);
#pragma unroll
    for (int k = 0; k < 4; ++k)
      load_parts8(x + (((size_t)n * H + 2 * oy + (k >> 1)) * W + 2 * ox + (k & 1)) * x_ld + x_coff + c, 1, 0, v[k]);
    const int me = (yy & 1) * 2 + (xx & 1);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int best = 0;
      float m = v[0][j];
#pragma unroll
      for (int k = 1; k < 4; ++k)
        if (v[k][j] > m) {
          m = v[k][j];
          best = k;
        }
      if (best == me) r[j] = g[j];
    }
  }
  store_parts8(dx + (((size_t)n * H + yy) * W + xx) * dx_ld + dx_coff + c, r, 1, 0);
}

// bilinear Upsample(align_corners=True) backward: every output pixel scatters its gradient to its 4 sources
__global__ void __launch_bounds__(256)
upsample_bilinear_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int dy_ld, int dy_coff, int N, int H, int W, int C,
                             int oH, int oW, float* __restrict__ dx) {
  const int cg = C / 8;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)N * oH * oW * cg) return;
  int c, ox, oy;
  long p = fast_div(i, cg, &c);
  c *= 8;
  p = fast_div(p, oW, &ox);
  const int n = (int)fast_div(p, oH, &oy);
  const float sy = oH > 1 ? (float)oy * ((float)(H - 1) / (float)(oH - 1)) : 0.f;
  const float sx = oW > 1 ? (float)ox * ((float)(W - 1) / (float)(oW - 1)) : 0.f;
  const int y0 = (int)sy, x0 = (int)sx;
  const int y1 = min(y0 + 1, H - 1), x1 = min(x0 + 1, W - 1);
  const float ly = sy - (float)y0, lx = sx - (float)x0;
  float g[8];
  load_parts8(dy + (((size_t)n * oH + oy) * oW + ox) * dy_ld + dy_coff + c, 1, 0, g);
  float* base = dx + (size_t)n * H * W * C + c;
  const float w[4] = {(1.f - ly) * (1.f - lx), (1.f - ly) * lx, ly * (1.f - lx), ly * lx};
  const size_t off[4] = {((size_t)y0 * W + x0) * C, ((size_t)y0 * W + x1) * C, ((size_t)y1 * W + x0) * C,
                         ((size_t)y1 * W + x1) * C};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (w[k] == 0.f) continue;
    float4* q = reinterpret_cast<float4*>(base + off[k]);
    atomicAdd(q, make_float4(w[k] * g[0], w[k] * g[1], w[k] * g[2], w[k] * g[3]));
    atomicAdd(q + 1, make_float4(w[k] * g[4], w[k] * g[5], w[k] * g[6], w[k] * g[7]));
  }
}

// The same backward in GATHER form: one thread per INPUT pixel and 8 channels sums w(o, i) * dy[o] over the output pixels
// whose bilinear footprint contains it.  Footprint membership and weights are recomputed with the forward kernel's own
// arithmetic (sy = oy * ((H-1)/(oH-1)), y0 = (int)sy, y1 = min(y0+1, H-1)), the candidate range is conservative and the
// sum runs in fixed (oy, ox) order: deterministic, no atomics, no zero-fill, the result written once (bf16 and / or
// fp32).  (The scatter kernel above spends 4 float4 atomics per output element: 441 us for FPN_LSS's 200x200x512 map.)
__global__ void __launch_bounds__(256)
upsample_bilinear_bwd_gather_kernel(const __nv_bfloat16* __restrict__ dy, int dy_ld, int dy_coff, int N, int H, int W,
                                    int C, int oH, int oW, __nv_bfloat16* __restrict__ dx16, int dx_ld, int dx_coff,
                                    float* __restrict__ dx32) {
  const int cg = C / 8;
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)N * H * W * cg) return;
  int c, x, y;
  long p = fast_div(i, cg, &c);
  c *= 8;
  p = fast_div(p, W, &x);
  const int n = (int)fast_div(p, H, &y);
  const float ry = oH > 1 ? (float)(H - 1) / (float)(oH - 1) : 0.f;
  const float rx = oW > 1 ? (float)(W - 1) / (float)(oW - 1) : 0.f;
  // outputs o with src = o * r in (i - 1, i + 1); r == 0 (a one-pixel axis): every output reads input 0
  int oy_lo = 0, oy_hi = oH - 1, ox_lo = 0, ox_hi = oW - 1;
  if (ry > 0.f) {
    oy_lo = max(0, (int)floorf((float)(y - 1) / ry) - 1);
    oy_hi = min(oH - 1, (int)ceilf((float)(y + 1) / ry) + 1);
  }
  if (rx > 0.f) {
    ox_lo = max(0, (int)floorf((float)(x - 1) / rx) - 1);
    ox_hi = min(oW - 1, (int)ceilf((float)(x + 1) / rx) + 1);
  }
  float r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = 0.f;
  for (int oy = oy_lo; oy <= oy_hi; ++oy) {
    const float sy = (float)oy * ry;
    const int y0 = (int)sy, y1 = min(y0 + 1, H - 1);
    const float ly = sy - (float)y0;
    const float wy = (y0 == y ? 1.f - ly : 0.f) + (y1 == y ? ly : 0.f);
    if (wy == 0.f) continue;
    const __nv_bfloat16* row = dy + ((size_t)n * oH + oy) * oW * dy_ld + dy_coff + c;
    for (int ox = ox_lo; ox <= ox_hi; ++ox) {
      const float sx = (float)ox * rx;
      const int x0 = (int)sx, x1 = min(x0 + 1, W - 1);
      const float lx = sx - (float)x0;
      const float wx = (x0 == x ? 1.f - lx : 0.f) + (x1 == x ? lx : 0.f);
      if (wx == 0.f) continue;
      float g[8];
      load_parts8(row + (size_t)ox * dy_ld, 1, 0, g);
      const float w = wy * wx;
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = fmaf(w, g[j], r[j]);
    }
  }
  const size_t pix = ((size_t)n * H + y) * W + x;
  if (dx32 != nullptr) {                       // before store_parts8, which leaves the rounding residual in r
    float4* q = reinterpret_cast<float4*>(dx32 + pix * C + c);
    q[0] = make_float4(r[0], r[1], r[2], r[3]);
    q[1] = make_float4(r[4], r[5], r[6], r[7]);
  }
  if (dx16 != nullptr) store_parts8(dx16 + pix * dx_ld + dx_coff + c, r, 1, 0);
}

// ---- unpack: NHWC split-bf16 -> fp32 NCHW (hand-off to reference-layout consumers) --------
__global__ void __launch_bounds__(256)
unpack_nhwc_to_nchw_kernel(const __nv_bfloat16* __restrict__ in, int ld, int coff, int part_stride,
                           int parts, int C, int HW, float* __restrict__ out) {
  __shared__ float tile[64][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 64, n = blockIdx.z;
  for (int p = ty; p < 32; p += 8) {
    for (int h = 0; h < 2; ++h) {
      const int c = tx + 32 * h;
      float v = 0.f;
      if (p0 + p < HW && c0 + c < C)
        v = load_parts(in + ((size_t)n * HW + p0 + p) * ld + coff + c0 + c, parts, part_stride);
      tile[c][p] = v;
    }
  }
  __syncthreads();
  for (int c = ty; c < 64; c += 8)
    if (c0 + c < C && p0 + tx < HW) out[((size_t)n * C + c0 + c) * HW + p0 + tx] = tile[c][tx];
}

// ---- per-image channel mean of an NHWC split-bf16 activation -> fp32 [N][C] -----------------
// grid (C/64, N, splits); partial sums are combined with atomicAdd into a pre-zeroed buffer
// when splits > 1 (then scaled by a second tiny launch), or written directly when splits == 1.
__global__ void __launch_bounds__(256)
mean_hw_kernel(const __nv_bfloat16* __restrict__ in, int ld, int coff, int part_stride, int parts,
               int C, int HW, float* __restrict__ out, float inv, int splits) {
  __shared__ float red[4][64];
  const int c = blockIdx.x * 64 + (threadIdx.x & 63), g = threadIdx.x >> 6, n = blockIdx.y;
  const int per = (HW + splits - 1) / splits;
  const int lo = blockIdx.z * per, hi = min(HW, lo + per);
  float s = 0.f;
  if (c < C)
    for (int p = lo + g; p < hi; p += 4)
      s += load_parts(in + ((size_t)n * HW + p) * ld + coff + c, parts, part_stride);
  red[g][threadIdx.x & 63] = s;
  __syncthreads();
  if (g == 0 && c < C) {
    s = red[0][threadIdx.x] + red[1][threadIdx.x] + red[2][threadIdx.x] + red[3][threadIdx.x];
    if (splits == 1) out[(size_t)n * C + c] = s * inv;
    else out[((size_t)n * splits + blockIdx.z) * C + c] = s;        // partial [n][split][C] for mean_finish_kernel
  }
}

// ---- y[r][o] = act(sum_k x[r][k] * w[o][k] + b[o]) (optionally x is first normalised by an
// eval-mode BatchNorm1d: (x - mean) * rstd * gamma + beta); one warp per output element -------
__global__ void __launch_bounds__(256)
linear_rows_kernel(const float* __restrict__ x, int R, int K, const float* __restrict__ w,
                   const float* __restrict__ b, int O, int act, const float* __restrict__ in_scale,
                   const float* __restrict__ in_shift, float* __restrict__ y, int one_minus) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (gw >= R * O) return;
  const int r = gw / O, o = gw % O;
  float s = 0.f;
  for (int k = lane; k < K; k += 32) {
    float v = x[(size_t)r * K + k];
    if (in_scale != nullptr) v = v * in_scale[k] + in_shift[k];
    s = fmaf(v, w[(size_t)o * K + k], s);
  }
  s = warp_sum(s);
  if (lane == 0) {
    if (b != nullptr) s += b[o];
    if (act == DHD_ACT_RELU) s = fmaxf(s, 0.f);
    else if (act == DHD_ACT_SIGMOID) s = 1.f / (1.f + expf(-s));
    if (one_minus) s = 1.f - s;
    y[(size_t)r * O + o] = s;
  }
}

// ---- SFA blends (mix.py:44-57), op order and roundings as torch evaluates them -------------
//   b1 = a1*bev ; v1 = (1-a1)*vox
//   a2 == null : out = b1 + v1                         (fea_U_1, input of spacial_leanring)
//   a2 != null : out = a2*b1 + (1-a2)*v1               (x_fuse), a2 = sigmoid already applied
// x: NHWC split-bf16 with bev channels [0,C) and voxel channels [C,2C); a1: [N][C] fp32.
__global__ void __launch_bounds__(256)
sfa_mix_kernel(const __nv_bfloat16* __restrict__ x, int x_ld, int x_coff, int x_ps, int x_parts, int C,
               long npix_total, int HW, const float* __restrict__ a1, const float* __restrict__ a2,
               __nv_bfloat16* __restrict__ out, int o_ld, int o_coff, int o_ps, int o_parts) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;   // (pixel, channel pair)
  const int cp = C / 2;
  if (i >= npix_total * cp) return;
  int c;
  const long pix = fast_div(i, cp, &c);
  c *= 2;
  const int n = (int)fast_div(pix, HW);
  const __nv_bfloat16* xb = x + (size_t)pix * x_ld + x_coff + c;
  float r[2];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const float bev = load_parts(xb + j, x_parts, x_ps);
    const float vox = load_parts(xb + C + j, x_parts, x_ps);
    const float g1 = a1[(size_t)n * C + c + j];
    const float b1 = __fmul_rn(g1, bev);
    const float v1 = __fmul_rn(__fsub_rn(1.f, g1), vox);
    if (a2 == nullptr) {
      r[j] = __fadd_rn(b1, v1);
    } else {
      const float g2 = a2[(size_t)pix * C + c + j];
      r[j] = __fadd_rn(__fmul_rn(g2, b1), __fmul_rn(__fsub_rn(1.f, g2), v1));
    }
  }
  store_parts2(out + (size_t)pix * o_ld + o_coff + c, r[0], r[1], o_parts, o_ps);
}

// ---- deformable im2col (mmcv DeformConv2dPack, deform_groups = 1, stride 1) -----------------
// x: NHWC split-bf16 (C channels); offset: fp32 NHWC [pix][2*taps] with (dy, dx) interleaved per
// tap; out: NHWC split-bf16 "image" with taps*C channels ordered [group][tap][C/groups] so that
// each conv group's K range is contiguous.  One warp per (pixel, tap); lanes own channel pairs.
__global__ void __launch_bounds__(256)
dcn_im2col_kernel(const __nv_bfloat16* __restrict__ x, int x_ld, int x_coff, int x_ps, int x_parts,
                  int C, int N, int H, int W, const float* __restrict__ offset, int off_ld, int ksize,
                  int pad, int dil, int groups, __nv_bfloat16* __restrict__ out, int o_ld, int o_ps,
                  int o_parts) {
  const int lane = threadIdx.x & 31;
  const long gw = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int taps = ksize * ksize;
  const long total = (long)N * H * W * taps;
  if (gw >= total) return;
  const int t = (int)(gw % taps);
  const long pix = gw / taps;
  const int wx = (int)(pix % W), hy = (int)((pix / W) % H), n = (int)(pix / ((long)W * H));
  const float dy = offset[(size_t)pix * off_ld + 2 * t], dx = offset[(size_t)pix * off_ld + 2 * t + 1];
  const float sy = (float)(hy - pad + (t / ksize) * dil) + dy;
  const float sx = (float)(wx - pad + (t % ksize) * dil) + dx;
  const int cg = C / groups;
  // bilinear weights; a sample is zero when it lies outside (-1, H) x (-1, W), and each corner
  // outside the image contributes zero (mmcv dmcn_im2col_bilinear / torchvision bilinear_interpolate)
  float w00 = 0.f, w01 = 0.f, w10 = 0.f, w11 = 0.f;
  int y0 = 0, x0 = 0, y1 = 0, x1 = 0;
  const bool inside = sy > -1.f && sx > -1.f && sy < (float)H && sx < (float)W;
  if (inside) {
    y0 = (int)floorf(sy);
    x0 = (int)floorf(sx);
    y1 = y0 + 1;
    x1 = x0 + 1;
    const float ly = sy - (float)y0, lx = sx - (float)x0, hy_ = 1.f - ly, hx_ = 1.f - lx;
    w00 = hy_ * hx_; w01 = hy_ * lx; w10 = ly * hx_; w11 = ly * lx;
  }
  for (int c = 2 * lane; c < C; c += 64) {
    float v[2] = {0.f, 0.f};
    if (inside) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const __nv_bfloat16* base = x + (size_t)n * H * W * x_ld + x_coff + c + j;
        float v00 = 0.f, v01 = 0.f, v10 = 0.f, v11 = 0.f;
        if (y0 >= 0 && x0 >= 0) v00 = load_parts(base + ((size_t)y0 * W + x0) * x_ld, x_parts, x_ps);
        if (y0 >= 0 && x1 <= W - 1) v01 = load_parts(base + ((size_t)y0 * W + x1) * x_ld, x_parts, x_ps);
        if (y1 <= H - 1 && x0 >= 0) v10 = load_parts(base + ((size_t)y1 * W + x0) * x_ld, x_parts, x_ps);
        if (y1 <= H - 1 && x1 <= W - 1) v11 = load_parts(base + ((size_t)y1 * W + x1) * x_ld, x_parts, x_ps);
        v[j] = w00 * v00 + w01 * v01 + w10 * v10 + w11 * v11;
      }
    }
    const int g = c / cg, cl = c % cg;
    store_parts2(out + (size_t)pix * o_ld + (size_t)g * taps * cg + (size_t)t * cg + cl, v[0], v[1],
                 o_parts, o_ps);
  }
}

// ---- occupancy class map: predictor.get_occ's softmax -> argmax -> uint8 (occ_head.py:141-153).  One thread per
// voxel, logits [voxel][ncls] fp32: plain argmax of the logits (first maximum wins, as torch.argmax), and where the
// runner-up sits within kSoftmaxTieGap of the maximum the exact restatement of torch's softmax kernel decides
// (common.cuh: rounding can merge two different logits into one probability, then the lower index wins).
template <int NCLS>
__global__ void __launch_bounds__(256)
occ_argmax_kernel(const float* __restrict__ logits, long nvox, int ncls, uint8_t* __restrict__ out) {
  const long v = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= nvox) return;
  const float* row = logits + v * ncls;
  if constexpr (NCLS > 0) {
    float x[NCLS];
#pragma unroll
    for (int k = 0; k < NCLS; ++k) x[k] = row[k];
    float best = x[0], second = -INFINITY;
    int arg = 0;
#pragma unroll
    for (int k = 1; k < NCLS; ++k) {
      second = fmaxf(second, fminf(x[k], best));
      if (x[k] > best) {
        best = x[k];
        arg = k;
      }
    }
    if (best - second <= kSoftmaxTieGap) arg = softmax_argmax_torch<NCLS>(x);
    out[v] = (uint8_t)arg;
    return;
  }
  float best = row[0];
  int arg = 0;
  for (int k = 1; k < ncls; ++k) {
    const float x = row[k];
    if (x > best) {
      best = x;
      arg = k;
    }
  }
  out[v] = (uint8_t)arg;
}

}  // namespace dhd

using namespace dhd;

extern "C" long dhd_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" int dhd_occ_argmax(const float* logits, long nvox, int ncls, uint8_t* out, void* stream) {
  DHD_REQUIRE(logits && out, "null pointer");
  DHD_REQUIRE(nvox > 0 && ncls > 0 && ncls <= 256, "bad shape");
  // 18 = Occ3D-nuScenes classes (every DHD config): exact softmax-tie handling; other class counts: argmax of the logits
  if (ncls == 18) occ_argmax_kernel<18><<<(int)((nvox + 255) / 256), 256, 0, (cudaStream_t)stream>>>(logits, nvox, ncls, out);
  else occ_argmax_kernel<0><<<(int)((nvox + 255) / 256), 256, 0, (cudaStream_t)stream>>>(logits, nvox, ncls, out);
  DHD_CUDA_LAUNCH_CHECK("occ_argmax");
  return DHD_OK;
}

extern "C" int dhd_pack_nchw_to_nhwc(const float* in, int N, int C, int H, int W, void* out, int out_ld,
                                     int out_coff, int part_stride, int parts, void* stream) {
  DHD_REQUIRE(in && out, "null pointer");
  DHD_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && parts >= 1 && parts <= 3, "bad shape");
  DHD_REQUIRE(out_ld % 2 == 0 && out_coff % 2 == 0 && part_stride % 2 == 0, "channel offsets must be even");
  dim3 grid((H * W + 31) / 32, (C + 63) / 64, N);
  pack_nchw_to_nhwc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      in, C, H * W, (__nv_bfloat16*)out, out_ld, out_coff, part_stride, parts);
  DHD_CUDA_LAUNCH_CHECK("pack_nchw_to_nhwc");
  return DHD_OK;
}

static bool vec8_ok(int C, int ld, int coff, int ps, const void* a, const void* b) {
  return C % 8 == 0 && C / 8 <= 256 && 256 % (C / 8) == 0 && ld % 8 == 0 && coff % 8 == 0 && ps % 8 == 0 &&
         ((uintptr_t)a & 15) == 0 && ((uintptr_t)b & 15) == 0;
}

// blocks per image so that the grid is a few waves of the SM count
static int pixels_per_block(int N, int HW) {
  const int want = max(1, (sm_count() * 8) / max(1, N));
  return max(16, (HW + want - 1) / want);
}

extern "C" size_t dhd_mean_workspace_bytes(int N, int C, int HW) {
  if (N <= 0 || C <= 0 || HW <= 0) return 0;
  const int pb = pixels_per_block(N, HW);
  const size_t nblk = (size_t)max((HW + pb - 1) / pb, 64);
  return (size_t)N * nblk * C * sizeof(float);
}

static int mean_finish(const float* partial, int N, int nblk, int C, float* out, float inv, cudaStream_t st) {
  mean_finish_kernel<<<N * ((C + 31) / 32), 256, 0, st>>>(partial, N, nblk, C, out, inv);
  DHD_CUDA_LAUNCH_CHECK("mean_finish");
  return DHD_OK;
}

extern "C" int dhd_split_nhwc_mean(const float* in, int N, int HW, int C, void* out, int out_ld, int out_coff,
                                   int part_stride, int parts, float* mean_out, float* workspace, void* stream) {
  DHD_REQUIRE(in && out, "null pointer");
  DHD_REQUIRE(N > 0 && HW > 0 && C > 0 && parts >= 1 && parts <= 3, "bad shape");
  DHD_REQUIRE(vec8_ok(C, out_ld, out_coff, part_stride, in, out), "split_nhwc_mean needs C % 8 == 0 and 16-byte alignment");
  DHD_REQUIRE(mean_out == nullptr || (workspace != nullptr && ((uintptr_t)workspace & 15) == 0),
              "mean_out needs a 16-byte aligned workspace of dhd_mean_workspace_bytes()");
  cudaStream_t st = (cudaStream_t)stream;
  const int pb = pixels_per_block(N, HW);
  const int nblk = (HW + pb - 1) / pb;
  split_nhwc8_kernel<<<dim3(nblk, N), 256, 0, st>>>(in, HW, C, pb, (__nv_bfloat16*)out, out_ld, out_coff, part_stride,
                                                    parts, mean_out != nullptr ? workspace : nullptr,
                                                    1.0f / (float)HW);
  DHD_CUDA_LAUNCH_CHECK("split_nhwc8");
  if (mean_out != nullptr) return mean_finish(workspace, N, nblk, C, mean_out, 1.0f / (float)HW, st);
  return DHD_OK;
}

extern "C" int dhd_split_nhwc(const float* in, long rows, int C, void* out, int out_ld, int out_coff,
                              int part_stride, int parts, void* stream) {
  DHD_REQUIRE(in && out, "null pointer");
  DHD_REQUIRE(rows > 0 && C > 0 && C % 2 == 0 && parts >= 1 && parts <= 3, "bad shape (C must be even)");
  DHD_REQUIRE(out_ld % 2 == 0 && out_coff % 2 == 0 && part_stride % 2 == 0, "channel offsets must be even");
  DHD_REQUIRE(((uintptr_t)in & 7) == 0, "input must be 8-byte aligned");
  if (vec8_ok(C, out_ld, out_coff, part_stride, in, out) && rows < (1L << 31))
    return dhd_split_nhwc_mean(in, 1, (int)rows, C, out, out_ld, out_coff, part_stride, parts, nullptr, nullptr, stream);
  const long total = rows * (C / 2);
  split_nhwc_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      in, rows, C, (__nv_bfloat16*)out, out_ld, out_coff, part_stride, parts);
  DHD_CUDA_LAUNCH_CHECK("split_nhwc");
  return DHD_OK;
}

extern "C" int dhd_unpack_nhwc_to_nchw(const void* in, int in_ld, int in_coff, int part_stride,
                                       int parts, int N, int C, int H, int W, float* out, void* stream) {
  DHD_REQUIRE(in && out, "null pointer");
  DHD_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && parts >= 1 && parts <= 3, "bad shape");
  dim3 grid((H * W + 31) / 32, (C + 63) / 64, N);
  unpack_nhwc_to_nchw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)in, in_ld, in_coff, part_stride, parts, C, H * W, out);
  DHD_CUDA_LAUNCH_CHECK("unpack_nhwc_to_nchw");
  return DHD_OK;
}

extern "C" int dhd_mean_hw(const void* in, int in_ld, int in_coff, int part_stride, int parts, int N,
                           int C, int HW, float* out, float* workspace, void* stream) {
  DHD_REQUIRE(in && out, "null pointer");
  DHD_REQUIRE(N > 0 && C > 0 && HW > 0 && parts >= 1 && parts <= 3, "bad shape");
  DHD_REQUIRE(workspace != nullptr && ((uintptr_t)workspace & 15) == 0,
              "mean_hw needs a 16-byte aligned workspace of dhd_mean_workspace_bytes()");
  cudaStream_t st = (cudaStream_t)stream;
  if (vec8_ok(C, in_ld, in_coff, part_stride, in, out)) {
    const int pb = pixels_per_block(N, HW);
    const int nblk = (HW + pb - 1) / pb;
    mean_hw8_kernel<<<dim3(nblk, N), 256, 0, st>>>((const __nv_bfloat16*)in, in_ld, in_coff, part_stride, parts, C,
                                                   HW, pb, workspace);
    DHD_CUDA_LAUNCH_CHECK("mean_hw8");
    return mean_finish(workspace, N, nblk, C, out, 1.0f / (float)HW, st);
  }
  const int cblocks = (C + 63) / 64;
  int splits = 1;
  if (HW > 4096) splits = min(64, max(1, (sm_count() * 4) / (cblocks * N)));
  mean_hw_kernel<<<dim3(cblocks, N, splits), 256, 0, st>>>((const __nv_bfloat16*)in, in_ld, in_coff,
                                                          part_stride, parts, C, HW, splits > 1 ? workspace : out,
                                                          1.0f / (float)HW, splits);
  DHD_CUDA_LAUNCH_CHECK("mean_hw");
  if (splits > 1) return mean_finish(workspace, N, splits, C, out, 1.0f / (float)HW, st);
  return DHD_OK;
}

extern "C" int dhd_linear_rows(const float* x, int R, int K, const float* w, const float* b, int O,
                               int act, const float* in_scale, const float* in_shift, int one_minus,
                               float* y, void* stream) {
  DHD_REQUIRE(x && w && y, "null pointer");
  DHD_REQUIRE(R > 0 && K > 0 && O > 0, "bad shape");
  DHD_REQUIRE((in_scale == nullptr) == (in_shift == nullptr), "in_scale / in_shift come together");
  const long warps = (long)R * O;
  linear_rows_kernel<<<(int)((warps * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      x, R, K, w, b, O, act, in_scale, in_shift, y, one_minus);
  DHD_CUDA_LAUNCH_CHECK("linear_rows");
  return DHD_OK;
}

extern "C" int dhd_gate_channels(const float* x, int N, int HW, int C, const float* gate, void* out, int o_ld,
                                 int o_coff, int o_part_stride, int o_parts, float* out32, void* stream) {
  DHD_REQUIRE(x && gate && out, "null pointer");
  DHD_REQUIRE(N > 0 && HW > 0 && C > 0 && C % 8 == 0 && o_parts >= 1 && o_parts <= 3, "bad shape (C % 8)");
  DHD_REQUIRE(o_ld % 8 == 0 && o_coff % 8 == 0 && o_part_stride % 8 == 0, "channel offsets must be multiples of 8");
  DHD_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)gate & 15) == 0 && ((uintptr_t)out & 15) == 0 &&
                  ((uintptr_t)out32 & 15) == 0, "pointers must be 16-byte aligned");
  const long total = (long)N * HW * (C / 8);
  gate_channels_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      x, C, (long)N * HW, HW, gate, (__nv_bfloat16*)out, o_ld, o_coff, o_part_stride, o_parts, out32);
  DHD_CUDA_LAUNCH_CHECK("gate_channels");
  return DHD_OK;
}

extern "C" int dhd_sfa_mix(const void* x, int x_ld, int x_coff, int x_part_stride, int x_parts, int C,
                           int N, int HW, const float* a1, const float* a2, void* out, int o_ld,
                           int o_coff, int o_part_stride, int o_parts, void* stream) {
  DHD_REQUIRE(x && a1 && out, "null pointer");
  DHD_REQUIRE(C > 0 && C % 2 == 0 && N > 0 && HW > 0, "bad shape");
  if (vec8_ok(C, x_ld, x_coff, x_part_stride, x, out) && o_ld % 8 == 0 && o_coff % 8 == 0 &&
      o_part_stride % 8 == 0 && ((uintptr_t)a1 & 15) == 0 && ((uintptr_t)a2 & 15) == 0) {
    const long total8 = (long)N * HW * (C / 8);
    sfa_mix8_kernel<<<(int)((total8 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x, x_ld, x_coff, x_part_stride, x_parts, C, (long)N * HW, HW, a1, a2,
        (__nv_bfloat16*)out, o_ld, o_coff, o_part_stride, o_parts);
    DHD_CUDA_LAUNCH_CHECK("sfa_mix8");
    return DHD_OK;
  }
  const long total = (long)N * HW * (C / 2);
  sfa_mix_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, x_ld, x_coff, x_part_stride, x_parts, C, (long)N * HW, HW, a1, a2,
      (__nv_bfloat16*)out, o_ld, o_coff, o_part_stride, o_parts);
  DHD_CUDA_LAUNCH_CHECK("sfa_mix");
  return DHD_OK;
}

extern "C" int dhd_sfa_blend_b16(const void* x, int x_ld, int x_coff, int C, int N, int HW, const float* a1,
                                 const void* a2, int a2_ld, int a2_coff, void* out, int o_ld, int o_coff,
                                 void* stream) {
  DHD_REQUIRE(x && a1 && a2 && out, "null pointer");
  DHD_REQUIRE(C > 0 && C % 8 == 0 && N > 0 && HW > 0, "C must be a positive multiple of 8");
  DHD_REQUIRE(x_ld % 8 == 0 && x_coff % 8 == 0 && a2_ld % 8 == 0 && a2_coff % 8 == 0 && o_ld % 8 == 0 && o_coff % 8 == 0 &&
                  ((uintptr_t)x & 15) == 0 && ((uintptr_t)a2 & 15) == 0 && ((uintptr_t)out & 15) == 0 &&
                  ((uintptr_t)a1 & 15) == 0, "16-byte aligned rows required");
  const long total8 = (long)N * HW * (C / 8);
  DHD_REQUIRE(total8 < (1L << 31) * 256, "tensor too large");
  sfa_blend_b16_kernel<<<(int)((total8 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, x_ld, x_coff, C, (long)N * HW, HW, a1, (const __nv_bfloat16*)a2, a2_ld, a2_coff,
      (__nv_bfloat16*)out, o_ld, o_coff);
  DHD_CUDA_LAUNCH_CHECK("sfa_blend_b16");
  return DHD_OK;
}

extern "C" int dhd_sfa_fold_gate(const float* w, const float* a1, int N, int Cout, int C, void* out, void* stream) {
  DHD_REQUIRE(w && a1 && out, "null pointer");
  DHD_REQUIRE(N > 0 && Cout > 0 && C > 0 && (long)N * Cout * C < (1L << 31), "bad shape");
  const int total = N * Cout * C;
  sfa_fold_gate_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(w, a1, N, Cout, C, (__nv_bfloat16*)out);
  DHD_CUDA_LAUNCH_CHECK("sfa_fold_gate");
  return DHD_OK;
}

extern "C" int dhd_dcn_im2col(const void* x, int x_ld, int x_coff, int x_part_stride, int x_parts, int C,
                              int N, int H, int W, const float* offset, int off_ld, int ksize, int pad,
                              int dilation, int groups, void* out, int o_ld, int o_part_stride,
                              int o_parts, void* stream) {
  DHD_REQUIRE(x && offset && out, "null pointer");
  DHD_REQUIRE(C > 0 && groups > 0 && C % groups == 0 && (C / groups) % 2 == 0, "bad channel grouping");
  DHD_REQUIRE(N > 0 && H > 0 && W > 0 && ksize >= 1 && ksize <= 3, "bad shape");
  const long warps = (long)N * H * W * ksize * ksize;
  DHD_REQUIRE(warps < (1L << 26), "too many sampling points for 32-bit indexing");
  if ((C / groups) % 8 == 0 && x_ld % 8 == 0 && x_coff % 8 == 0 && x_part_stride % 8 == 0 && o_ld % 8 == 0 &&
      o_part_stride % 8 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)out & 15) == 0) {
    static const bool per_tap = [] {                    // DHD_DCN_IM2COL=tap: the earlier warp-per-(pixel, tap) kernel (A/B)
      const char* e = getenv("DHD_DCN_IM2COL");
      return e != nullptr && e[0] == 't';
    }();
    if (per_tap) {
      dcn_im2col8_kernel<<<(int)((warps * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
          (const __nv_bfloat16*)x, x_ld, x_coff, x_part_stride, x_parts, C, N, H, W, offset, off_ld, ksize, pad,
          dilation, groups, (__nv_bfloat16*)out, o_ld, o_part_stride, o_parts);
      DHD_CUDA_LAUNCH_CHECK("dcn_im2col8");
      return DHD_OK;
    }
    const long pixels = (long)N * H * W;
    dcn_im2col_px_kernel<<<(int)((pixels * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x, x_ld, x_coff, x_part_stride, x_parts, C, N, H, W, offset, off_ld, ksize, pad,
        dilation, groups, (__nv_bfloat16*)out, o_ld, o_part_stride, o_parts);
    DHD_CUDA_LAUNCH_CHECK("dcn_im2col_px");
    return DHD_OK;
  }
  dcn_im2col_kernel<<<(int)((warps * 32 + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, x_ld, x_coff, x_part_stride, x_parts, C, N, H, W, offset, off_ld, ksize,
      pad, dilation, groups, (__nv_bfloat16*)out, o_ld, o_part_stride, o_parts);
  DHD_CUDA_LAUNCH_CHECK("dcn_im2col");
  return DHD_OK;
}

extern "C" int dhd_maxpool2(const void* in, int in_ld, int in_coff, int in_part_stride, int N, int H, int W, int C,
                            void* out, int out_ld, int out_coff, int out_part_stride, int parts, void* stream) {
  DHD_REQUIRE(in && out && N > 0 && H >= 2 && W >= 2 && C > 0 && parts >= 1 && parts <= 3, "bad arguments");
  DHD_REQUIRE(C % 8 == 0 && in_ld % 8 == 0 && in_coff % 8 == 0 && out_ld % 8 == 0 && out_coff % 8 == 0 &&
                  in_part_stride % 8 == 0 && out_part_stride % 8 == 0 &&
                  ((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0, "needs C % 8 == 0 and 16-byte aligned rows");
  const long total = (long)N * (H / 2) * (W / 2) * (C / 8);
  maxpool2_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)in, in_ld, in_coff, in_part_stride, N, H, W, C, (__nv_bfloat16*)out, out_ld, out_coff,
      out_part_stride, parts);
  DHD_CUDA_LAUNCH_CHECK("maxpool2");
  return DHD_OK;
}

extern "C" int dhd_upsample_bilinear(const void* in, int in_ld, int in_coff, int in_part_stride, int N, int H, int W,
                                     int C, int out_H, int out_W, void* out, int out_ld, int out_coff,
                                     int out_part_stride, int parts, void* stream) {
  DHD_REQUIRE(in && out && N > 0 && H > 0 && W > 0 && C > 0 && out_H > 0 && out_W > 0 && parts >= 1 && parts <= 3,
              "bad arguments");
  DHD_REQUIRE(C % 8 == 0 && in_ld % 8 == 0 && in_coff % 8 == 0 && out_ld % 8 == 0 && out_coff % 8 == 0 &&
                  in_part_stride % 8 == 0 && out_part_stride % 8 == 0 &&
                  ((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0, "needs C % 8 == 0 and 16-byte aligned rows");
  const long total = (long)N * out_H * out_W * (C / 8);
  upsample_bilinear_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)in, in_ld, in_coff, in_part_stride, N, H, W, C, out_H, out_W, (__nv_bfloat16*)out, out_ld,
      out_coff, out_part_stride, parts);
  DHD_CUDA_LAUNCH_CHECK("upsample_bilinear");
  return DHD_OK;
}

extern "C" int dhd_upsample_bilinear_bwd_gather(const void* dy, int dy_ld, int dy_coff, int N, int H, int W, int C,
                                                int out_H, int out_W, void* dx_b16, int dx_ld, int dx_coff, float* dx_f32,
                                                void* stream) {
  DHD_REQUIRE(dy && (dx_b16 || dx_f32) && N > 0 && H > 0 && W > 0 && C > 0 && out_H > 0 && out_W > 0, "bad arguments");
  DHD_REQUIRE(C % 8 == 0 && dy_ld % 8 == 0 && dy_coff % 8 == 0 && ((uintptr_t)dy & 15) == 0, "dy: C % 8, 16-byte aligned rows");
  if (dx_b16 != nullptr)
    DHD_REQUIRE(dx_ld % 8 == 0 && dx_coff % 8 == 0 && ((uintptr_t)dx_b16 & 15) == 0, "dx_b16: 16-byte aligned rows");
  if (dx_f32 != nullptr) DHD_REQUIRE(((uintptr_t)dx_f32 & 15) == 0, "dx_f32: 16-byte aligned");
  const long total = (long)N * H * W * (C / 8);
  upsample_bilinear_bwd_gather_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)dy, dy_ld, dy_coff, N, H, W, C, out_H, out_W, (__nv_bfloat16*)dx_b16, dx_ld, dx_coff, dx_f32);
  DHD_CUDA_LAUNCH_CHECK("upsample_bilinear_bwd_gather");
  return DHD_OK;
}

extern "C" int dhd_stem_im2col(const float* img, int N, int Cin, int H, int W, int ksize, int stride, int pad, void* out,
                               int out_ld, int out_part_stride, int parts, void* stream) {
  DHD_REQUIRE(img && out && N > 0 && Cin > 0 && H > 0 && W > 0 && ksize >= 1 && stride >= 1 && pad >= 0 && parts >= 1 &&
                  parts <= 3, "bad arguments");
  const int K = ksize * ksize * Cin;
  DHD_REQUIRE(out_part_stride % 8 == 0 && out_part_stride >= (K + 7) / 8 * 8 && out_ld % 8 == 0 &&
                  out_ld >= (parts - 1) * out_part_stride + (K + 7) / 8 * 8 && ((uintptr_t)out & 15) == 0,
              "output rows must hold ksize*ksize*Cin values per part, 16-byte aligned");
  const int oH = (H + 2 * pad - ksize) / stride + 1, oW = (W + 2 * pad - ksize) / stride + 1;
  DHD_REQUIRE(K <= 512 && ksize <= 255 && Cin <= 255, "stem im2col: ksize*ksize*Cin <= 512");
  const int groups = out_part_stride / 8;            // the whole row of every part is written (zeros beyond K)
  StemTable T;                                         // K index -> (dy, dx, c), K ordered (ky, kx, c)
  for (int k = 0; k < 512; ++k) {
    const int c = k % Cin, t = k / Cin;
    T.e[k] = k < K ? ((uint32_t)(t / ksize) | ((uint32_t)(t % ksize) << 8) | ((uint32_t)c << 16)) : 0u;
  }
  const int tiles_x = (oW + kStemPx - 1) / kStemPx;
  const long blocks = (long)N * oH * tiles_x;
  DHD_REQUIRE(blocks < (1L << 31), "image batch too large");
  const size_t smem = (size_t)Cin * ksize * ((kStemPx - 1) * stride + ksize) * sizeof(float);
  DHD_REQUIRE(smem <= 48 * 1024, "stem patch does not fit in shared memory");
  stem_im2col_kernel<<<(unsigned)blocks, 256, smem, (cudaStream_t)stream>>>(
      T, img, N, Cin, H, W, ksize, stride, pad, oH, oW, K, groups, (__nv_bfloat16*)out, out_ld, out_part_stride, parts);
  DHD_CUDA_LAUNCH_CHECK("stem_im2col");
  return DHD_OK;
}

extern "C" int dhd_maxpool3s2(const void* in, int in_ld, int in_coff, int in_part_stride, int N, int H, int W, int C,
                              void* out, int out_ld, int out_coff, int out_part_stride, int parts, void* stream) {
  DHD_REQUIRE(in && out && N > 0 && H >= 1 && W >= 1 && C > 0 && parts >= 1 && parts <= 3, "bad arguments");
  DHD_REQUIRE(C % 8 == 0 && in_ld % 8 == 0 && in_coff % 8 == 0 && out_ld % 8 == 0 && out_coff % 8 == 0 &&
                  in_part_stride % 8 == 0 && out_part_stride % 8 == 0 &&
                  ((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0, "needs C % 8 == 0 and 16-byte aligned rows");
  const int oH = (H - 1) / 2 + 1, oW = (W - 1) / 2 + 1;                  // floor((H + 2 - 3) / 2) + 1
  const long total = (long)N * oH * oW * (C / 8);
  maxpool3s2_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)in, in_ld, in_coff, in_part_stride, N, H, W, C, oH, oW, (__nv_bfloat16*)out, out_ld, out_coff,
      out_part_stride, parts);
  DHD_CUDA_LAUNCH_CHECK("maxpool3s2");
  return DHD_OK;
}

extern "C" int dhd_upsample_nearest_add(const void* lo, int lo_ld, int lo_coff, int lo_part_stride, int h, int w, void* io,
                                        int io_ld, int io_coff, int io_part_stride, int N, int H, int W, int C, int parts,
                                        void* stream) {
  DHD_REQUIRE(lo && io && N > 0 && h > 0 && w > 0 && H > 0 && W > 0 && C > 0 && parts >= 1 && parts <= 3, "bad arguments");
  DHD_REQUIRE(C % 8 == 0 && lo_ld % 8 == 0 && lo_coff % 8 == 0 && io_ld % 8 == 0 && io_coff % 8 == 0 &&
                  lo_part_stride % 8 == 0 && io_part_stride % 8 == 0 &&
                  ((uintptr_t)lo & 15) == 0 && ((uintptr_t)io & 15) == 0, "needs C % 8 == 0 and 16-byte aligned rows");
  const long total = (long)N * H * W * (C / 8);
  upsample_nearest_add_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)lo, lo_ld, lo_coff, lo_part_stride, h, w, (__nv_bfloat16*)io, io_ld, io_coff, io_part_stride, N,
      H, W, C, parts);
  DHD_CUDA_LAUNCH_CHECK("upsample_nearest_add");
  return DHD_OK;
}

extern "C" int dhd_maxpool3s2_bwd(const void* x, int x_ld, int x_coff, const void* dy, int dy_ld, int dy_coff, int N,
                                  int H, int W, int C, void* dx, int dx_ld, int dx_coff, int relu_mask, void* stream) {
  DHD_REQUIRE(x && dy && dx && N > 0 && H > 0 && W > 0 && C > 0, "bad arguments");
  DHD_REQUIRE(C % 8 == 0 && x_ld % 8 == 0 && x_coff % 8 == 0 && dy_ld % 8 == 0 && dy_coff % 8 == 0 && dx_ld % 8 == 0 &&
                  dx_coff % 8 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)dy & 15) == 0 && ((uintptr_t)dx & 15) == 0,
              "needs C % 8 == 0 and 16-byte aligned rows");
  const int oH = (H - 1) / 2 + 1, oW = (W - 1) / 2 + 1;
  const long total = (long)N * H * W * (C / 8);
  static const bool ref = [] {                          // DHD_MAXPOOL_BWD=ref: the first (arg-max bookkeeping) kernel (A/B)
    const char* e = getenv("DHD_MAXPOOL_BWD");
    return e != nullptr && e[0] == 'r';
  }();
  (ref ? maxpool3s2_bwd_ref_kernel : maxpool3s2_bwd_kernel)<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, x_ld, x_coff, (const __nv_bfloat16*)dy, dy_ld, dy_coff, N, H, W, C, oH, oW,
      (__nv_bfloat16*)dx, dx_ld, dx_coff, relu_mask);
  DHD_CUDA_LAUNCH_CHECK("maxpool3s2_bwd");
  return DHD_OK;
}

extern "C" int dhd_maxpool2_bwd(const void* x, int x_ld, int x_coff, const void* dy, int dy_ld, int dy_coff, int N,
                                int H, int W, int C, void* dx, int dx_ld, int dx_coff, void* stream) {
  DHD_REQUIRE(x && dy && dx && N > 0 && H >= 2 && W >= 2 && C > 0, "bad arguments");
  DHD_REQUIRE(C % 8 == 0 && x_ld % 8 == 0 && x_coff % 8 == 0 && dy_ld % 8 == 0 && dy_coff % 8 == 0 && dx_ld % 8 == 0 &&
                  dx_coff % 8 == 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)dy & 15) == 0 && ((uintptr_t)dx & 15) == 0,
              "needs C % 8 == 0 and 16-byte aligned rows");
  const long total = (long)N * H * W * (C / 8);
  maxpool2_bwd_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      (const __nv_bfloat16*)x, x_ld, x_coff, (const __nv_bfloat16*)dy, dy_ld, dy_coff, N, H, W, C, (__nv_bfloat16*)dx,
      dx_ld, dx_coff);
  DHD_CUDA_LAUNCH_CHECK("maxpool2_bwd");
  return DHD_OK;
}

extern "C" int dhd_upsample_bilinear_bwd(const void* dy, int dy_ld, int dy_coff, int N, int H, int W, int C, int out_H,
                                         int out_W, float* dx, void* stream) {
  DHD_REQUIRE(dy && dx && N > 0 && H > 0 && W > 0 && C > 0 && out_H > 0 && out_W > 0, "bad arguments");
  DHD_REQUIRE(C % 8 == 0 && dy_ld % 8 == 0 && dy_coff % 8 == 0 && ((uintptr_t)dy & 15) == 0 && ((uintptr_t)dx & 15) == 0,
              "needs C % 8 == 0 and 16-byte aligned rows");
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(dx, 0, (size_t)N * H * W * C * sizeof(float), st);
  if (e != cudaSuccess) return fail((int)e, "%s: %ld", "memset(dx)", (long)e);
  const long total = (long)N * out_H * out_W * (C / 8);
  upsample_bilinear_bwd_kernel<<<(int)((total + 255) / 256), 256, 0, st>>>((const __nv_bfloat16*)dy, dy_ld, dy_coff, N, H,
                                                                          W, C, out_H, out_W, dx);
  DHD_CUDA_LAUNCH_CHECK("upsample_bilinear_bwd");
  return DHD_OK;
}
