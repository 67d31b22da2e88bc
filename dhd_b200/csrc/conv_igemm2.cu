// Persistent implicit-GEMM convolution, second generation (same dhd_conv_desc contract as
// conv_igemm.cu, selected by dhd_conv2d_fwd unless DHD_CONV_V=1).
//
// What changed against the first kernel and why (B200 measurements in profiles/):
//  * the first kernel was bound by L2 -> shared-memory traffic (128x128x64 steps move 32 KB per
//    2.1 MFLOP = 64 FLOP/B; the chip delivers ~6.5 TB/s from L2, so ~400 TFLOP/s): the N tile is
//    now 256 wide for layers with Cout > 128 (85 FLOP/B) -- A is fetched once for 256 channels;
//  * it is persistent (one CTA per SM walks the tile list) with the accumulator double-buffered
//    in TMEM (2 x NT columns), so tile i's epilogue overlaps tile i+1's main loop and the
//    TMEM allocation / barrier set-up is paid once per SM instead of once per tile;
//  * the epilogue reads scale / bias / per-image vectors from shared memory (staged once per
//    tile) instead of issuing two global loads per element, and channel-contiguous outputs
//    (every bf16 activation and every NHWC fp32 tensor) leave through 128B-swizzled shared
//    memory and TMA box stores: full-line coalesced writes, image borders clipped by the TMA
//    unit; strided outputs (NCHW heads) keep the direct path.
#include <stdlib.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace dhd {

constexpr int kM2 = 128;          // pixels per tile (TMEM lanes)
constexpr int kK2 = 64;           // bf16 per smem row (128 B swizzle span)
// G groups of four epilogue warps (G = 2 or 4): group g owns the 32-column chunks with chunk % G == g
constexpr uint32_t kA2Bytes = kM2 * kK2 * 2;
constexpr uint32_t kStageBufBytes = 16384;   // one TMA-store staging tile: 128 rows x 128 B
constexpr int kMaxConvBatch = DHD_CONV_MAX_BATCH;

// CTAS = 2: a CTA pair (cluster of two on one TPC) computes a 256-pixel x NT tile with ONE tcgen05.mma.cta_group::2
// per k-step: each CTA stages its own 128 pixels of A and HALF of the weight tile, so a k-step moves 32 KB per SM
// instead of 48 KB.  The 3x3 layers at 200x200 ran at 0.49 us per 64-channel step against 0.27 us of tensor-core time
// (profiles/r01_conv3x3_sfa_full.txt: the pipe waits on operand tiles, L2 -> shared memory at ~14 TB/s chip-wide);
// fewer bytes per FLOP and a fourth ring stage in the freed shared memory attack exactly that.
template <int NT, int CTAS = 1>
struct Cfg2 {
  static constexpr int kStages = CTAS == 2 ? 4 : (NT == 256 ? 3 : 4);
  static constexpr uint32_t kBBytes = (NT / CTAS) * kK2 * 2;
  static constexpr uint32_t kStageBytes = kA2Bytes + kBBytes;
  static constexpr uint32_t kVecBytes = 3 * NT * 4 + 4 * 2048; // scale, bias(+img_bias), gate; statistics scratch per group
  static constexpr uint32_t kSmem = kStages * kStageBytes + 4 * kStageBufBytes + kVecBytes + 256 + 1024;
  static constexpr int kTmemCols = 2 * NT;
};

struct Conv2Params {
  dhd_conv_desc d;
  int tiles_w, tiles_h, n_tiles, total_tiles;
  int seg_b16_tma[DHD_CONV_MAX_SEGS], seg_f32_tma[DHD_CONV_MAX_SEGS];
  int seg_b16_wide[DHD_CONV_MAX_SEGS];      // bf16 output leaves in 64-channel (128-byte row) TMA boxes
};

struct Conv2Maps {
  CUtensorMap a, b;
  CUtensorMap o16[DHD_CONV_MAX_SEGS][3];
  CUtensorMap o16w[DHD_CONV_MAX_SEGS];      // 64-channel boxes, SWIZZLE_128B (single-part bf16 outputs)
  CUtensorMap o32[DHD_CONV_MAX_SEGS];
};

// One launch = NP independent convolutions ("problems") with the same N-tile class: the persistent CTAs walk the
// concatenated tile list (problem 0's tiles, then problem 1's, ...).  Layers that read the same activation and are
// each too small to fill the GPU (the four ASPP branches, the groups of the deformable convolution, depth_net next to
// HeightNet's first convolution: 132 tiles for 148 SMs, one tile per CTA, nothing overlaps) become one launch whose
// CTAs hold 3-4 tiles each, so tile i's epilogue runs under tile i+1's main loop and the prologue is paid once.
// Everything sits in kernel parameter space (<= 32 KB since CUDA 12.1), indexed by the problem id.
template <int NP>
struct ConvBatch {
  Conv2Maps M[NP];
  Conv2Params P[NP];
  int n;                 // problems in use
  int tile_end[NP];      // running total of tiles
};

__device__ __forceinline__ float act2(float v, int act) {
  switch (act) {
    case DHD_ACT_RELU: return fmaxf(v, 0.f);
    case DHD_ACT_SIGMOID: return 1.f / (1.f + expf(-v));
    case DHD_ACT_SOFTPLUS: return v > 20.f ? v : log1pf(expf(v));
    default: return v;
  }
}

__device__ __forceinline__ bool tap_dead2(const dhd_conv_desc& d, int t, int x0, int y0) {
  const int st = d.stride > 1 ? d.stride : 1;
  const int iw = d.in_W > 0 ? d.in_W : d.W, ih = d.in_H > 0 ? d.in_H : d.H;
  const int xs = x0 * st + d.tap_dx[t], ys = y0 * st + d.tap_dy[t];
  return xs >= iw || xs + d.bw * st <= 0 || ys >= ih || ys + d.bh * st <= 0;
}

template <int NT, int G, int NP, int CTAS>
__global__ void __launch_bounds__((4 * G + 2) * 32, 1)
conv_igemm2_kernel(const __grid_constant__ ConvBatch<NP> B) {
  static_assert(CTAS == 1 || (NP == 1 && NT == 256), "CTA pairs: single problem, 256-wide N tile");
  using C = Cfg2<NT, CTAS>;
  constexpr int kEpiWarps = 4 * G;
  constexpr int kBufPerGroup = 4 / G;          // staging tiles per group: 4 x 16 KB in total
  extern __shared__ uint8_t smem_raw[];
  // CTAS == 2: work items are PAIRS of M tiles (2i, 2i+1) x one N tile; cluster c walks items c, c + #clusters, ..
  const uint32_t crank = CTAS == 2 ? cluster_ctarank() : 0u;
  const bool leader = crank == 0;
  const int total_tiles = CTAS == 2 ? ((B.P[0].total_tiles / B.P[0].n_tiles + 1) / 2) * B.P[0].n_tiles
                                    : B.tile_end[B.n - 1];
  const int tile_first = CTAS == 2 ? (int)cluster_id_x() : (int)blockIdx.x;
  const int tile_step = CTAS == 2 ? (int)cluster_nid_x() : (int)gridDim.x;
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  const uint32_t stagebuf = base + C::kStages * C::kStageBytes;            // 2 groups x 2 x 16 KB, 1024-aligned
  const uint32_t vec_base = stagebuf + 4 * kStageBufBytes;
  const uint32_t bar_base = vec_base + C::kVecBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (C::kStages + s); };
  auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * C::kStages + s); };
  auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * C::kStages + 2 + s); };
  uint8_t* gen = smem_raw + (base - raw);                                   // generic view of `base`
  float* s_scale = reinterpret_cast<float*>(gen + (vec_base - base));
  float* s_bias = s_scale + NT;
  float* s_gate = s_bias + NT;
  float* s_stat = s_gate + NT;          // [group][4 warps][64 channels][sum, sum of squares]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + (bar_base - base) + 8u * (2 * C::kStages + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t kIdesc = umma_instr_desc_bf16(kM2 * CTAS, NT);

  if (warp == kEpiWarps && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&B.M[0].a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&B.M[0].b) : "memory");
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar(s), 1);
      mbar_init(tempty_bar(s), G * CTAS);   // one arrival per epilogue group (of both CTAs of a pair)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kEpiWarps + 1) {
    if (CTAS == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "n"(C::kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "n"(C::kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();          // the peer's barriers exist before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // global tile index -> problem id (and the tile index inside that problem)
  auto problem_of = [&](int& tile) {
    int p = 0;
    if (NP > 1) {
      while (tile >= B.tile_end[p]) ++p;
      if (p > 0) tile -= B.tile_end[p - 1];
    }
    return p;
  };
  // CTAS == 2: `tile` is a pair item; rank r of the pair owns M tile 2*(item / n_tiles) + r.  Returns false for the
  // missing second tile of an odd tile count (img = N: every TMA box is out of bounds -> zeros, nothing is stored)
  auto decode = [&](const Conv2Params& P, int tile, int& img, int& x0, int& y0, int& n0, uint32_t rank = 0u) -> bool {
    const int nt = tile % P.n_tiles;
    int mt = tile / P.n_tiles;
    if (CTAS == 2) mt = 2 * mt + (int)rank;
    n0 = nt * NT;
    if (CTAS == 2 && mt >= P.total_tiles / P.n_tiles) {
      img = P.d.N;
      x0 = y0 = 0;
      return false;
    }
    const int tx = mt % P.tiles_w;
    mt /= P.tiles_w;
    const int ty = mt % P.tiles_h;
    img = mt / P.tiles_h;
    x0 = tx * P.d.bw;
    y0 = ty * P.d.bh;
    return true;
  };
  // a filter tap is skipped when its shifted box lies outside the image -- for a CTA pair: outside for BOTH tiles
  auto tap_skipped = [&](const Conv2Params& P, int t, int item, int x0, int y0, bool valid) {
    bool dead = !valid || tap_dead2(P.d, t, x0, y0);
    if (CTAS == 2 && dead) {
      int img2, x2, y2, n2;
      const bool v2 = decode(P, item, img2, x2, y2, n2, crank ^ 1u);
      dead = !v2 || tap_dead2(P.d, t, x2, y2);
    }
    return dead;
  };

  if (warp == kEpiWarps) {
    // ===================================================== TMA producer
    if (lane == 0) {
      int it = 0;
      const uint32_t lead_full0 = CTAS == 2 ? mapa_u32(full_bar(0), 0u) : 0u;
      for (int gtile = tile_first; gtile < total_tiles; gtile += tile_step) {
        int tile = gtile;
        const int pid = problem_of(tile);
        const Conv2Params& P = B.P[pid];
        const Conv2Maps& M = B.M[pid];
        const dhd_conv_desc& d = P.d;
        const int kchunks = d.Cin / kK2;
        const int in_stride = d.stride > 1 ? d.stride : 1;     // the tensor map strides the box (elementStrides)
        int img, x0, y0, n0;
        const bool tvalid = decode(P, tile, img, x0, y0, n0, crank);
        for (int t = 0; t < d.taps; ++t) {
          if (tap_skipped(P, t, tile, x0, y0, tvalid)) continue;
          for (int kc = 0; kc < kchunks; ++kc) {
            for (int e = 0; e < d.n_terms; ++e, ++it) {
              const int s = it % C::kStages;
              const uint32_t ph = (it / C::kStages) & 1;
              mbar_wait(empty_bar(s), ph ^ 1);
              const uint32_t sa = base + s * C::kStageBytes, sb = sa + kA2Bytes;
              if (CTAS == 2) {
                // both CTAs' bytes are counted by the LEADER's barrier (the MMA issuer waits there)
                if (leader) mbar_expect_tx(full_bar(s), 2 * C::kStageBytes);
                const uint32_t lf = lead_full0 + 8u * s;
                tma_load_4d_2sm(sa, &M.a, lf, d.in_coff + d.term_a[e] * d.in_part_stride + kc * kK2,
                                x0 * in_stride + d.tap_dx[t], y0 * in_stride + d.tap_dy[t], img);
                tma_load_2d_2sm(sb, &M.b, lf, (t * d.w_parts + d.term_b[e]) * d.Cin + kc * kK2,
                                n0 + (int)crank * (NT / 2));
                continue;
              }
              mbar_expect_tx(full_bar(s), C::kStageBytes);
              tma_load_4d(sa, &M.a, full_bar(s), d.in_coff + d.term_a[e] * d.in_part_stride + kc * kK2,
                          x0 * in_stride + d.tap_dx[t], y0 * in_stride + d.tap_dy[t], img);
              tma_load_2d(sb, &M.b, full_bar(s), (t * d.w_parts + d.term_b[e]) * d.Cin + kc * kK2, n0 + img * d.w_image_rows);
            }
          }
        }
      }
    }
  } else if (warp == kEpiWarps + 1) {
    // ===================================================== MMA issuer
    if (lane == 0 && leader) {
      int it = 0, lt = 0;
      for (int gtile = tile_first; gtile < total_tiles; gtile += tile_step, ++lt) {
        int tile = gtile;
        const int pid = problem_of(tile);
        const Conv2Params& P = B.P[pid];
        const dhd_conv_desc& d = P.d;
        const int kchunks = d.Cin / kK2;
        int img, x0, y0, n0;
        const bool tvalid = decode(P, tile, img, x0, y0, n0, crank);
        const int as = lt & 1;
        mbar_wait(tempty_bar(as), ((lt >> 1) & 1) ^ 1);      // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(as * NT);
        int first = 1;
        for (int t = 0; t < d.taps; ++t) {
          if (tap_skipped(P, t, tile, x0, y0, tvalid)) continue;
          for (int kc = 0; kc < kchunks; ++kc) {
            for (int e = 0; e < d.n_terms; ++e, ++it) {
              const int s = it % C::kStages;
              const uint32_t ph = (it / C::kStages) & 1;
              mbar_wait(full_bar(s), ph);
              tc_fence_after();
              const uint32_t sa = base + s * C::kStageBytes, sb = sa + kA2Bytes;
              const uint64_t da = umma_desc_sw128(sa), db = umma_desc_sw128(sb);
#pragma unroll
              for (int k = 0; k < kK2 / kUmmaK; ++k) {
                if (CTAS == 2) umma_bf16_2sm(tacc, da + 2u * k, db + 2u * k, kIdesc, (first && k == 0) ? 0u : 1u);
                else umma_bf16(tacc, da + 2u * k, db + 2u * k, kIdesc, (first && k == 0) ? 0u : 1u);
              }
              first = 0;
              if (CTAS == 2) umma_commit_2sm(empty_bar(s));      // frees the stage in both CTAs
              else umma_commit(empty_bar(s));
            }
          }
        }
        if (CTAS == 2) umma_commit_2sm(tfull_bar(as));
        else umma_commit(tfull_bar(as));
      }
    }
  } else {
    // ===================================================== epilogue (warps 0..3 = TMEM lanes 32w..)
    // Kept deliberately COMPACT (32-column chunks in a rolled loop, one activation switch per
    // chunk): the first version unrolled 64 columns x every activation and its 220 KB of SASS
    // made the four epilogue warps stall on instruction fetch (ncu: stall_no_inst 34 %).
    // Two groups of four warps: warp w reads TMEM lanes 32*(w%4).. (the hardware's lane window of a
    // warp), group g = w/4 takes the 32-column chunks with (chunk & 1) == g.  1x1 layers are bound
    // by this epilogue (4 k-steps per 256 columns), so the second group doubles their throughput;
    // each group has its own pair of staging tiles and its own named barrier.
    const int grp = warp >> 2;              // 0 / 1
    const int tid = threadIdx.x & 127;      // 0..127 inside the group
    const int gbar = 1 + grp;               // named barrier of this group
    const uint32_t gstage = stagebuf + (uint32_t)grp * kBufPerGroup * kStageBufBytes;
    const int row = tid;
    int lt = 0;
    uint32_t nstore = 0;                    // TMA stores issued so far by this group (selects the staging buffer)
    const uint32_t lead_tempty0 = CTAS == 2 ? mapa_u32(tempty_bar(0), 0u) : 0u;
    for (int gtile = tile_first; gtile < total_tiles; gtile += tile_step, ++lt) {
      int tile = gtile;
      const int pid = problem_of(tile);
      const Conv2Params& P = B.P[pid];
      const Conv2Maps& M = B.M[pid];
      const dhd_conv_desc& d = P.d;
      int img, x0, y0, n0;
      const bool tvalid = decode(P, tile, img, x0, y0, n0, crank);
      const int as = lt & 1;
      if (CTAS == 2 && !tvalid) {             // the missing half of the last pair: only the accumulator hand-shake
        mbar_wait(tfull_bar(as), (lt >> 1) & 1);
        tc_fence_before();
        named_bar_sync(gbar, 128);
        if (tid == 0) mbar_arrive_cluster(lead_tempty0 + 8u * as);
        continue;
      }
      named_bar_sync(5, 128 * G);           // everyone is done with the previous tile's vectors
      for (int c = threadIdx.x; c < NT; c += 128 * G) {
        const int ch = n0 + c;
        float sc = 1.f, bi = 0.f, ga = 1.f;
        if (ch < d.Cout) {
          if (d.scale != nullptr) sc = __ldg(d.scale + ch);
          if (d.bias != nullptr) bi = __ldg(d.bias + ch);
          if (d.img_bias != nullptr) bi += __ldg(d.img_bias + (size_t)img * d.Cout + ch);
          if (d.img_gate != nullptr) ga = __ldg(d.img_gate + (size_t)img * d.Cout + ch);
          if (d.mix_x != nullptr) ga = __ldg(d.mix_a1 + (size_t)img * d.Cout + ch);     // a1 rides in the gate slot
        }
        s_scale[c] = sc;
        s_bias[c] = bi;
        s_gate[c] = ga;
      }
      named_bar_sync(5, 128 * G);
      mbar_wait(tfull_bar(as), (lt >> 1) & 1);
      tc_fence_after();
      const int px = x0 + row % d.bw, py = y0 + row / d.bw;
      const bool valid = px < d.W && py < d.H;
      const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(as * NT);
      const float* res = nullptr;
      if (d.residual != nullptr && valid)
        res = d.residual + (size_t)img * d.res_sN + (size_t)py * d.res_sY + (size_t)px * d.res_sX + n0;
      const __nv_bfloat16* resb = nullptr;
      if (d.res_b16 != nullptr && valid)
        resb = reinterpret_cast<const __nv_bfloat16*>(d.res_b16) +
               ((size_t)img * d.H * d.W + (size_t)py * d.W + px) * d.res_b16_ld + d.res_b16_coff + n0;
      const bool has_gate = d.img_gate != nullptr;

#pragma unroll 1
      for (int sgi = 0; sgi < d.n_seg; ++sgi) {
        const dhd_conv_seg& sg = d.seg[sgi];
        const int c_lo = max(sg.c_lo, n0), c_hi = min(sg.c_hi, min(d.Cout, n0 + NT));
        if (c_lo >= c_hi) continue;
        const bool tma16 = P.seg_b16_tma[sgi] != 0, tma32 = P.seg_f32_tma[sgi] != 0;
        const int act = sg.act;
        const int cb_lo = (c_lo - n0) / 32, cb_hi = (c_hi - n0 + 31) / 32;
        // y = acc*scale + bias (+ residual) for one 32-column chunk, out-of-segment columns -> -inf / 0
        auto affine = [&](int cb, float (&v)[32], float fill) {
          tmem_ld32(taddr + cb * 32, v);
          const float4* sc4 = reinterpret_cast<const float4*>(s_scale + cb * 32);
          const float4* bi4 = reinterpret_cast<const float4*>(s_bias + cb * 32);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 s = sc4[j], b = bi4[j];
            v[4 * j] = fmaf(v[4 * j], s.x, b.x);
            v[4 * j + 1] = fmaf(v[4 * j + 1], s.y, b.y);
            v[4 * j + 2] = fmaf(v[4 * j + 2], s.z, b.z);
            v[4 * j + 3] = fmaf(v[4 * j + 3], s.w, b.w);
          }
          if (res != nullptr) {
            const float* rp = res + cb * 32;
#pragma unroll
            for (int j = 0; j < 8; ++j) {     // host guarantees 16-byte alignment and Cout % 32 == 0
              const float4 q = __ldg(reinterpret_cast<const float4*>(rp) + j);
              v[4 * j] += q.x; v[4 * j + 1] += q.y; v[4 * j + 2] += q.z; v[4 * j + 3] += q.w;
            }
          }
          if (resb != nullptr) {
            const uint4* rp = reinterpret_cast<const uint4*>(resb + cb * 32);
#pragma unroll
            for (int j = 0; j < 4; ++j) {     // host guarantees 16-byte alignment and Cout % 32 == 0
              const uint4 q = __ldg(rp + j);
              const __nv_bfloat162* hq = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                v[8 * j + 2 * i] += __low2float(hq[i]);
                v[8 * j + 2 * i + 1] += __high2float(hq[i]);
              }
            }
          }
          if (n0 + cb * 32 < c_lo || n0 + cb * 32 + 32 > c_hi) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int c = n0 + cb * 32 + j;
              if (c < c_lo || c >= c_hi) v[j] = fill;
            }
          }
        };
        if (P.seg_b16_wide[sgi] != 0) {
          // ---- bf16-only output, 64 channels (one 128-byte row) per TMA store.  The 32-channel path below issues one
          // 8 KB box store (128 rows x 64 B) per chunk and waits for the store two chunks back to release its staging
          // tile: on the 1x1 layers (4 k-steps per 256 columns) that store cadence bounds the tile (ncu
          // profiles/r02_conv_sfa_full.txt: tensor pipe 14 % active, the epilogue warps parked behind
          // cp.async.bulk.wait_group.read).  Twice the bytes per store and per barrier pair halves both.
          const int wb_lo = (c_lo - n0) / 64, wb_hi = (c_hi - n0 + 63) / 64;
#pragma unroll 1
          for (int wb = wb_lo; wb < wb_hi; ++wb) {
            if (wb % G != grp) continue;
            uint4 q[8];
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(q);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const int cb = wb * 2 + half;
              float v[32];
              if (n0 + cb * 32 < c_hi) {
                affine(cb, v, 0.f);
                switch (act) {
                  case DHD_ACT_RELU:
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                    break;
                  case DHD_ACT_SIGMOID:
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __fdividef(1.f, 1.f + __expf(-v[j]));
                    break;
                  case DHD_ACT_SOFTPLUS:
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = v[j] > 20.f ? v[j] : __logf(1.f + __expf(v[j]));
                    break;
                  default: break;
                }
                if (has_gate) {
                  const float4* g4 = reinterpret_cast<const float4*>(s_gate + cb * 32);
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    const float4 g = g4[j];
                    v[4 * j] *= g.x; v[4 * j + 1] *= g.y; v[4 * j + 2] *= g.z; v[4 * j + 3] *= g.w;
                  }
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0.f;
              }
#pragma unroll
              for (int j = 0; j < 16; ++j) h[half * 16 + j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
            }
            const bool stats = d.stat_partial != nullptr;
            if (stats && !valid) {                // rows beyond the image border must not count (the TMA store clips them)
#pragma unroll
              for (int j = 0; j < 8; ++j) q[j] = make_uint4(0u, 0u, 0u, 0u);
            }
            const uint32_t buf = gstage + (nstore % kBufPerGroup) * kStageBufBytes;
            if (tid == 0) tma_store_wait_read<kBufPerGroup - 1>();
            named_bar_sync(gbar, 128);
            // 64 channels = 128-byte rows, SWIZZLE_128B: 16-byte chunk j of row r sits at j ^ (r & 7)
            uint4* dst = reinterpret_cast<uint4*>(gen + (buf - base) + row * 128);
#pragma unroll
            for (int j = 0; j < 8; ++j) dst[j ^ (row & 7)] = q[j];
            fence_proxy_async_smem();
            named_bar_sync(gbar, 128);
            if (tid == 0) {
              tma_store_4d(&M.o16w[sgi], buf, n0 + wb * 64 - sg.c_lo, x0, y0, img);
              tma_store_commit();
            }
            ++nstore;
            if (stats) {
              // BatchNorm batch statistics of the layer output, from the staged tile (the bf16 values the consumer will
              // read): warp w sums rows [32w, 32w+32) of the 64 channels, lane = channel pair (one bank per lane), the
              // four warps meet in shared memory, and partial row `tile` = [sum | sum of squares][Cout] goes to global
              // memory; dhd_colsum_finish adds the rows in fixed order
              const int cp = tid & 31, qr = tid >> 5;
              const uint8_t* tb = gen + (buf - base);
              float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll 8
              for (int rr = 0; rr < 32; ++rr) {
                const int r = qr * 32 + rr;
                const uint32_t wv = *reinterpret_cast<const uint32_t*>(tb + r * 128 + (((cp >> 2) ^ (r & 7)) << 4) + ((cp & 3) << 2));
                const float lo = __uint_as_float(wv << 16), hi = __uint_as_float(wv & 0xFFFF0000u);
                s0 += lo; q0 = fmaf(lo, lo, q0);
                s1 += hi; q1 = fmaf(hi, hi, q1);
              }
              float4* sred = reinterpret_cast<float4*>(s_stat + grp * 512);
              sred[qr * 32 + cp] = make_float4(s0, q0, s1, q1);          // channel 2cp: (sum, sq), channel 2cp+1: (sum, sq)
              named_bar_sync(gbar, 128);
              const int ch = n0 + wb * 64 + tid;
              if (tid < 64 && ch < d.Cout) {
                const float2* sr = reinterpret_cast<const float2*>(s_stat + grp * 512);
                float2 a = sr[tid];
#pragma unroll
                for (int w4 = 1; w4 < 4; ++w4) {                         // fixed order
                  const float2 b = sr[w4 * 64 + tid];
                  a.x += b.x;
                  a.y += b.y;
                }
                const int mt_idx = CTAS == 2 ? 2 * (tile / P.n_tiles) + (int)crank : tile / P.n_tiles;
                float* pr = d.stat_partial + (size_t)mt_idx * 2 * d.Cout + ch;
                pr[0] = a.x;
                pr[d.Cout] = a.y;
              }
            }
          }
          continue;
        }
        // one rolled loop nest: softmax makes three passes over the chunks (max, sum, write), every
        // other activation one; a single inlined copy of `affine` keeps the code small
        float mx = -INFINITY, sum = 0.f, inv = 1.f;
        const int npass = act == DHD_ACT_SOFTMAX ? 3 : 1;
#pragma unroll 1
        for (int pass = 0; pass < npass; ++pass) {
        if (pass == 2) inv = 1.f / sum;
#pragma unroll 1
        for (int cb = cb_lo; cb < cb_hi; ++cb) {
          if ((npass == 1 || pass == 2) && cb % G != grp) continue;       // row statistics need every chunk
          float v[32];
          affine(cb, v, act == DHD_ACT_SOFTMAX ? -INFINITY : 0.f);
          if (npass == 3 && pass == 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j) mx = fmaxf(mx, v[j]);
            continue;
          }
          if (npass == 3 && pass == 1) {
#pragma unroll
            for (int j = 0; j < 32; ++j) sum += __expf(v[j] - mx);    // exp(-inf) = 0 outside the segment
            continue;
          }
          switch (act) {
            case DHD_ACT_RELU:
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
              break;
            case DHD_ACT_SIGMOID:
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = __fdividef(1.f, 1.f + __expf(-v[j]));
              break;
            case DHD_ACT_SOFTPLUS:   // torch Softplus(beta=1, threshold=20)
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = v[j] > 20.f ? v[j] : __logf(1.f + __expf(v[j]));
              break;
            case DHD_ACT_SOFTMAX:
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = __expf(v[j] - mx) * inv;
              break;
            default: break;
          }
          if (d.mix_x != nullptr) {
            // v = spatial gate a2 -> fuse = a2 * a1*bev + (1 - a2) * (1 - a1)*vox (same operation order as dhd_sfa_mix)
            const int cfirst_m = n0 + cb * 32;
            if (valid && cfirst_m < d.Cout) {
              const __nv_bfloat16* mp = reinterpret_cast<const __nv_bfloat16*>(d.mix_x) +
                                        ((size_t)img * d.H * d.W + (size_t)py * d.W + px) * d.mix_ld + d.mix_coff + cfirst_m;
#pragma unroll
              for (int j8 = 0; j8 < 4; ++j8) {
                float bev[8], vox[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) bev[j] = vox[j] = 0.f;
                for (int p = 0; p < d.mix_parts; ++p) {
                  const uint4 qb = __ldg(reinterpret_cast<const uint4*>(mp + (size_t)p * d.mix_part_stride + j8 * 8));
                  const uint4 qv = __ldg(reinterpret_cast<const uint4*>(mp + (size_t)p * d.mix_part_stride + d.Cout + j8 * 8));
                  const __nv_bfloat162* hb = reinterpret_cast<const __nv_bfloat162*>(&qb);
                  const __nv_bfloat162* hv = reinterpret_cast<const __nv_bfloat162*>(&qv);
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    bev[2 * j] += __low2float(hb[j]); bev[2 * j + 1] += __high2float(hb[j]);
                    vox[2 * j] += __low2float(hv[j]); vox[2 * j + 1] += __high2float(hv[j]);
                  }
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  const float a1 = s_gate[cb * 32 + j8 * 8 + j], g2 = v[j8 * 8 + j];
                  const float b1 = __fmul_rn(a1, bev[j]), v1 = __fmul_rn(__fsub_rn(1.f, a1), vox[j]);
                  v[j8 * 8 + j] = __fadd_rn(__fmul_rn(g2, b1), __fmul_rn(__fsub_rn(1.f, g2), v1));
                }
              }
            }
          } else if (has_gate) {
            const float4* g4 = reinterpret_cast<const float4*>(s_gate + cb * 32);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 g = g4[j];
              v[4 * j] *= g.x; v[4 * j + 1] *= g.y; v[4 * j + 2] *= g.z; v[4 * j + 3] *= g.w;
            }
          }
          const int cfirst = n0 + cb * 32;
          // ---------------- fp32 output
          if (sg.out_f32 != nullptr) {
            if (tma32) {
              const uint32_t buf = gstage + (nstore % kBufPerGroup) * kStageBufBytes;
              if (tid == 0) tma_store_wait_read<kBufPerGroup - 1>();   // the store that last used this buffer has read it
              named_bar_sync(gbar, 128);
              float4* dst = reinterpret_cast<float4*>(gen + (buf - base) + row * 128);
#pragma unroll
              for (int j = 0; j < 8; ++j)
                dst[j ^ (row & 7)] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
              fence_proxy_async_smem();
              named_bar_sync(gbar, 128);
              if (tid == 0) {
                tma_store_4d(&M.o32[sgi], buf, cfirst - sg.c_lo, x0, y0, img);
                tma_store_commit();
              }
              ++nstore;
            } else if (valid) {
              float* o = sg.out_f32 + (size_t)img * sg.f32_sN + (size_t)py * sg.f32_sY + (size_t)px * sg.f32_sX;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int c = cfirst + j;
                if (c >= c_lo && c < c_hi) o[(size_t)(c - sg.c_lo) * sg.f32_sC] = v[j];
              }
            }
          }
          // ---------------- bf16 (split) output: part p = bf16 of what parts < p left over
          if (sg.out_b16 != nullptr) {
#pragma unroll 1
            for (int p = 0; p < sg.b16_parts; ++p) {
              uint4 q[4];
              __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(q);
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
                v[2 * j] -= __low2float(h[j]);
                v[2 * j + 1] -= __high2float(h[j]);
              }
              if (tma16) {
                // 32 channels = 64-byte rows, SWIZZLE_64B: 16-byte chunk j of row r sits at j ^ ((r >> 1) & 3)
                const uint32_t buf = gstage + (nstore % kBufPerGroup) * kStageBufBytes;
                if (tid == 0) tma_store_wait_read<kBufPerGroup - 1>();
                named_bar_sync(gbar, 128);
                uint4* dst = reinterpret_cast<uint4*>(gen + (buf - base) + row * 64);
#pragma unroll
                for (int j = 0; j < 4; ++j) dst[j ^ ((row >> 1) & 3)] = q[j];
                fence_proxy_async_smem();
                named_bar_sync(gbar, 128);
                if (tid == 0) {
                  tma_store_4d(&M.o16[sgi][p], buf, cfirst - sg.c_lo, x0, y0, img);
                  tma_store_commit();
                }
                ++nstore;
              } else if (valid) {
                const size_t pixoff = sg.b16_sX != 0
                                          ? (size_t)img * sg.b16_sN + (size_t)py * sg.b16_sY + (size_t)px * sg.b16_sX
                                          : ((size_t)img * d.H * d.W + (size_t)py * d.W + px) * sg.b16_ld;
                __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(sg.out_b16) + pixoff + sg.b16_coff +
                                    (size_t)p * sg.b16_part_stride;
                const __nv_bfloat16* hs = reinterpret_cast<const __nv_bfloat16*>(q);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  const int c = cfirst + j;
                  if (c >= c_lo && c < c_hi) op[c - sg.c_lo] = hs[j];
                }
              }
            }
          }
        }
        }
      }
      // accumulator drained: hand it back to the MMA warp
      tc_fence_before();
      named_bar_sync(gbar, 128);
      if (tid == 0) {
        if (CTAS == 2) mbar_arrive_cluster(lead_tempty0 + 8u * as);     // the pair's MMA issuer lives in the leader CTA
        else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty_bar(as)) : "memory");
      }
    }
    if (tid == 0) tma_store_wait<0>();      // all bulk stores complete before the CTA exits
  }
  __syncthreads();
  if (CTAS == 2) cluster_sync_all();          // the leader's MMAs read the peer's shared memory and write its TMEM
  if (warp == kEpiWarps + 1) {
    tc_fence_after();
    if (CTAS == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::kTmemCols)
                   : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::kTmemCols)
                   : "memory");
  }
}

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);

template <int NT, int G, int NP>
static int launch2(const ConvBatch<NP>& batch, cudaStream_t st) {
  using C = Cfg2<NT, 1>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv_igemm2_kernel<NT, G, NP, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)C::kSmem);
    if (e != cudaSuccess) return fail((int)e, "%s: %ld", "cudaFuncSetAttribute(conv_igemm2)", (long)e);
    attr_set = true;
  }
  const int grid = min(batch.tile_end[batch.n - 1], sm_count());
  conv_igemm2_kernel<NT, G, NP, 1><<<grid, (4 * G + 2) * 32, C::kSmem, st>>>(batch);
  DHD_CUDA_LAUNCH_CHECK("conv_igemm2");
  return DHD_OK;
}

static int epi_groups();

// CTA-pair variant (NT = 256, one problem): a cluster of two CTAs per work item.  Returns DHD_EUNSUPPORTED when the
// device cannot co-schedule a single pair (the caller then uses the one-CTA kernel).
static int max_pair_clusters() {
  using C = Cfg2<256, 2>;
  static int n = -1;
  if (n < 0) {
    auto kern = conv_igemm2_kernel<256, 2, 1, 2>;
    n = 0;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::kSmem) == cudaSuccess) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(2 * sm_count());
      cfg.blockDim = dim3(10 * 32);
      cfg.dynamicSmemBytes = C::kSmem;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = 2;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      int c = 0;
      if (cudaOccupancyMaxActiveClusters(&c, kern, &cfg) == cudaSuccess) n = c;
      else cudaGetLastError();
    }
  }
  return n;
}

static int launch2_pair(const ConvBatch<1>& batch, cudaStream_t st) {
  using C = Cfg2<256, 2>;
  const int clusters_max = max_pair_clusters();
  if (clusters_max <= 0) return DHD_EUNSUPPORTED;
  const Conv2Params& P = batch.P[0];
  const int items = ((P.total_tiles / P.n_tiles + 1) / 2) * P.n_tiles;
  const int clusters = min(items, clusters_max);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(10 * 32);
  cfg.dynamicSmemBytes = C::kSmem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, conv_igemm2_kernel<256, 2, 1, 2>, batch);
  if (e != cudaSuccess) return fail((int)e, "%s: %ld", "conv_igemm2 (CTA pairs) launch failed", (long)e);
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return DHD_OK;
}

static int epi_groups() {
  static const int groups = [] {
    const char* v = getenv("DHD_CONV_EPI_GROUPS");
    return v != nullptr && atoi(v) == 4 ? 4 : 2;
  }();
  return groups;
}

static int conv2_fill(const dhd_conv_desc* d, EncodeTiledFn2 enc, Conv2Maps& maps, Conv2Params& P, int NT, int ctas);

// CTA pairs serve the wide layers (Cout > 128) with shared weights; DHD_CONV_CTA2=0 keeps everything on one-CTA tiles
static int g_pair_mode = -1;            // -1: not read yet; 0 off; 1 where it pays; 2 wherever it is possible
int conv2_pair_mode(int set) {          // dhd_conv_pair_mode (conv_igemm.cu): A/B switch for tests and benchmarks
  if (g_pair_mode < 0) {
    const char* v = getenv("DHD_CONV_CTA2");
    g_pair_mode = v == nullptr ? 1 : max(0, min(2, atoi(v)));
  }
  const int prev = g_pair_mode;
  if (set >= 0) g_pair_mode = min(set, 2);
  return prev;
}
// Measured on B200 (profiles/r02_conv_pair_ab.txt, sustained clocks): pairs win on the tensor-bound layers with many
// tiles per SM (3x3 256->256 at 4x200x200: 172.9 -> 167.2 us) and lose on the epilogue-bound 1x1 layers (+5..13 %) and
// on single-wave layers (132 tiles: 20.4 -> 21.3 us), so mode 1 uses them for K >= 1024 and >= 4 tiles per SM only.
static bool pair_mode(const dhd_conv_desc* d) {
  const int mode = conv2_pair_mode(-1);
  if (mode == 0 || d->Cout <= 128 || d->w_image_rows != 0 || epi_groups() != 2) return false;
  if (mode == 1) {
    const long tiles = (long)((d->W + d->bw - 1) / d->bw) * ((d->H + d->bh - 1) / d->bh) * d->N;
    if ((long)d->taps * d->Cin * d->n_terms < 1024 || tiles < 4L * sm_count()) return false;
  }
  return max_pair_clusters() > 0;
}

// called by dhd_conv2d_fwd (conv_igemm.cu) after argument validation
int conv2_launch(const dhd_conv_desc* d, void* encode, void* stream) {
  const int NT = d->Cout > 128 ? 256 : 128;
  ConvBatch<1> batch;
  const bool pair = pair_mode(d);
  int rc = conv2_fill(d, (EncodeTiledFn2)encode, batch.M[0], batch.P[0], NT, pair ? 2 : 1);
  if (rc != DHD_OK) return rc;
  batch.n = 1;
  batch.tile_end[0] = batch.P[0].total_tiles;
  if (pair) return launch2_pair(batch, (cudaStream_t)stream);
  if (epi_groups() == 4) {
    if (NT == 256) return launch2<256, 4, 1>(batch, (cudaStream_t)stream);
    return launch2<128, 4, 1>(batch, (cudaStream_t)stream);
  }
  if (NT == 256) return launch2<256, 2, 1>(batch, (cudaStream_t)stream);
  return launch2<128, 2, 1>(batch, (cudaStream_t)stream);
}

// dhd_conv2d_fwd_batch: n <= kMaxConvBatch validated descriptors in one launch.  The N tile is 256 wide when any
// problem has Cout > 128 (a narrower problem then wastes tensor-core columns, not time that matters at these sizes).
int conv2_launch_batch(const dhd_conv_desc* const* descs, int n, void* encode, void* stream) {
  int NT = 128;
  for (int i = 0; i < n; ++i)
    if (descs[i]->Cout > 128) NT = 256;
  ConvBatch<kMaxConvBatch> batch;
  int total = 0;
  for (int i = 0; i < kMaxConvBatch; ++i) {
    if (i < n) {
      int rc = conv2_fill(descs[i], (EncodeTiledFn2)encode, batch.M[i], batch.P[i], NT, 1);
      if (rc != DHD_OK) return rc;
      total += batch.P[i].total_tiles;
    }
    batch.tile_end[i] = total;
  }
  batch.n = n;
  if (NT == 256) return launch2<256, 2, kMaxConvBatch>(batch, (cudaStream_t)stream);
  return launch2<128, 2, kMaxConvBatch>(batch, (cudaStream_t)stream);
}

static int conv2_fill(const dhd_conv_desc* d, EncodeTiledFn2 enc, Conv2Maps& maps, Conv2Params& P, int NT, int ctas) {
  P.d = *d;
  {
    const int st = d->stride > 1 ? d->stride : 1;
    const cuuint64_t iw = d->in_W > 0 ? d->in_W : d->W, ih = d->in_H > 0 ? d->in_H : d->H;   // input grid (>= output grid)
    cuuint64_t dims[4] = {(cuuint64_t)d->in_ld, iw, ih, (cuuint64_t)d->N};
    cuuint64_t strides[3] = {(cuuint64_t)d->in_ld * 2, iw * d->in_ld * 2, ih * iw * d->in_ld * 2};
    // stride-2 layers: the box spans 2*bw x 2*bh input pixels and the TMA unit keeps every 2nd one
    cuuint32_t box[4] = {(cuuint32_t)kK2, (cuuint32_t)(d->bw * st), (cuuint32_t)(d->bh * st), 1};
    cuuint32_t es[4] = {1, (cuuint32_t)st, (cuuint32_t)st, 1};
    CUresult r = enc(&maps.a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)d->in, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(DHD_EINVAL, "%s: %ld", "cuTensorMapEncodeTiled(A) failed", (long)r);
  }
  {
    const cuuint64_t ktot = (cuuint64_t)d->taps * d->w_parts * d->Cin;
    const cuuint64_t wrows = d->w_image_rows > 0 ? (cuuint64_t)d->w_image_rows * d->N : (cuuint64_t)d->Cout;
    cuuint64_t dims[2] = {ktot, wrows};
    cuuint64_t strides[1] = {ktot * 2};
    cuuint32_t box[2] = {(cuuint32_t)kK2, (cuuint32_t)(NT / ctas)};      // a CTA pair: each CTA stages half the rows
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&maps.b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)d->weight, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(DHD_EINVAL, "%s: %ld", "cuTensorMapEncodeTiled(B) failed", (long)r);
  }
  for (int s = 0; s < DHD_CONV_MAX_SEGS; ++s) {
    P.seg_b16_tma[s] = 0;
    P.seg_f32_tma[s] = 0;
    P.seg_b16_wide[s] = 0;
    if (s >= d->n_seg) continue;
    const dhd_conv_seg& sg = d->seg[s];
    const cuuint64_t nch = (cuuint64_t)(sg.c_hi - sg.c_lo);
    // a TMA store may hang over the far edges of the tensor but its start coordinate must not be
    // negative: 32-channel boxes start at multiples of 32, so the segment has to as well
    if (sg.c_lo % 32 != 0) continue;
    cuuint32_t es[4] = {1, 1, 1, 1};
    if (sg.out_b16 != nullptr && sg.b16_ld % 8 == 0 && sg.b16_coff % 8 == 0 && sg.b16_part_stride % 8 == 0 &&
        sg.b16_sX % 8 == 0 && sg.b16_sY % 8 == 0 && sg.b16_sN % 8 == 0 && ((uintptr_t)sg.out_b16 & 15) == 0) {
      bool ok = true;
      for (int p = 0; p < sg.b16_parts && ok; ++p) {
        cuuint64_t dims[4] = {nch, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
        cuuint64_t strides[3] = {(cuuint64_t)sg.b16_ld * 2, (cuuint64_t)d->W * sg.b16_ld * 2,
                                 (cuuint64_t)d->H * d->W * sg.b16_ld * 2};
        if (sg.b16_sX != 0) {            // strided pixel view
          strides[0] = (cuuint64_t)sg.b16_sX * 2;
          strides[1] = (cuuint64_t)sg.b16_sY * 2;
          strides[2] = (cuuint64_t)sg.b16_sN * 2;
        }
        cuuint32_t box[4] = {32, (cuuint32_t)d->bw, (cuuint32_t)d->bh, 1};   // 64-byte rows
        void* basep = (char*)sg.out_b16 + ((size_t)sg.b16_coff + (size_t)p * sg.b16_part_stride) * 2;
        ok = enc(&maps.o16[s][p], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, basep, dims, strides, box, es,
                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
      }
      P.seg_b16_tma[s] = ok ? 1 : 0;
      // single-part bf16-only segments whose channel range starts on a 64-channel boundary of the N tile: wide boxes
      static const bool wide_ok = [] { const char* v = getenv("DHD_CONV_WIDE_STORE"); return v == nullptr || atoi(v) != 0; }();
      if (ok && wide_ok && sg.b16_parts == 1 && sg.out_f32 == nullptr && sg.act != DHD_ACT_SOFTMAX && d->mix_x == nullptr &&
          sg.c_lo % 64 == 0) {
        cuuint64_t dims[4] = {nch, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
        cuuint64_t strides[3] = {(cuuint64_t)sg.b16_ld * 2, (cuuint64_t)d->W * sg.b16_ld * 2,
                                 (cuuint64_t)d->H * d->W * sg.b16_ld * 2};
        if (sg.b16_sX != 0) {
          strides[0] = (cuuint64_t)sg.b16_sX * 2;
          strides[1] = (cuuint64_t)sg.b16_sY * 2;
          strides[2] = (cuuint64_t)sg.b16_sN * 2;
        }
        cuuint32_t box[4] = {64, (cuuint32_t)d->bw, (cuuint32_t)d->bh, 1};   // 128-byte rows
        void* basep = (char*)sg.out_b16 + (size_t)sg.b16_coff * 2;
        P.seg_b16_wide[s] = enc(&maps.o16w[s], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, basep, dims, strides, box, es,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS ? 1 : 0;
      }
    }
    if (sg.out_f32 != nullptr && sg.f32_sC == 1 && sg.f32_sX % 4 == 0 && sg.f32_sY % 4 == 0 &&
        sg.f32_sN % 4 == 0 && ((uintptr_t)sg.out_f32 & 15) == 0 && sg.act != DHD_ACT_SOFTMAX) {
      cuuint64_t dims[4] = {nch, (cuuint64_t)d->W, (cuuint64_t)d->H, (cuuint64_t)d->N};
      cuuint64_t strides[3] = {(cuuint64_t)sg.f32_sX * 4, (cuuint64_t)sg.f32_sY * 4, (cuuint64_t)sg.f32_sN * 4};
      cuuint32_t box[4] = {32, (cuuint32_t)d->bw, (cuuint32_t)d->bh, 1};
      P.seg_f32_tma[s] = enc(&maps.o32[s], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, (void*)sg.out_f32, dims, strides,
                             box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                             CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
                             ? 1 : 0;
    }
  }
  if (d->stat_partial != nullptr && P.seg_b16_wide[0] == 0)
    return fail(DHD_EUNSUPPORTED, "%s", "fused statistics need the 64-channel TMA store path (DHD_CONV_WIDE_STORE=0 or a bad view?)");
  P.tiles_w = (d->W + d->bw - 1) / d->bw;
  P.tiles_h = (d->H + d->bh - 1) / d->bh;
  P.n_tiles = (d->Cout + NT - 1) / NT;
  P.total_tiles = P.tiles_w * P.tiles_h * d->N * P.n_tiles;
  return DHD_OK;
}

}  // namespace dhd
