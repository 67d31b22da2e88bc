// Fused occupancy-head tail: Linear(256 -> 512) + Softplus + Linear(512 -> Dz*n_cls) [+ argmax over the classes]
// as ONE back-to-back tcgen05 GEMM kernel (reference: models/dense_heads/occ_head.py:63-67 `predicter`, 84-100
// forward incl. the permute(0, 3, 2, 1), 141-153 get_occ's softmax -> argmax -> uint8).
//
// Why: as three launches (GEMM + softplus epilogue, GEMM, argmax) the tail cost 379 us of a 1.93 ms step and moved
// 164 MB (hidden layer out) + 164 MB (in again) + 2 x 184 MB (logits out / in) through HBM.  Here a CTA owns a
// 128-pixel tile and never lets the hidden layer or -- at inference -- the logits leave the SM:
//
//   A tile   [128 px x 256] bf16, resident in shared memory for the whole tile (4 TMA boxes, 64 KB)
//   for each 64-column chunk c of the hidden layer (8 chunks):
//     G1(c)  acc1[c % 3] (TMEM, 64 cols)  = A x W1[c]^T            16 x tcgen05.mma 128x64x16, W1 streamed by TMA
//     epi(c) 8 warps: tcgen05.ld -> + b1 -> softplus (ex2 / lg2) -> bf16 -> 128B-swizzled smem tile h[c % 2]
//     G2(c)  acc2 (TMEM, 288 cols)       += h[c % 2] x W2[:, c]^T   4 k-steps x 2 x tcgen05.mma 128x144x16
//   final    8 more warps: tcgen05.ld acc2 -> + b2 -> fp32 logits (optional) and / or per-z argmax over the classes -> uint8
//
// Warp roles (20 warps): 0-7 chunk epilogue, 8 G1 issuer, 9 G2 issuer, 10 weight producer, 11 A producer, 12-19 logits
// epilogue.  The first version had ONE issuing thread for both GEMMs and the chunk-epilogue warps also drained the
// logits: ncu (profiles/r02_tail_full.txt) showed the tensor pipe 33 % active -- the issuing thread's ~430 scalar
// instructions per chunk and its blocking wait for h[c] before it could issue G1(c + 1) put issuer and epilogue into
// lock-step, and every tile boundary stalled the whole pipeline behind the 288-column logits drain.  Now G1 runs up to
// three chunks ahead (acc1 triple-buffered) on its own warp, G2 follows the epilogue on another, and the logits of
// tile t are drained by their own warps while tile t + 1's chunks are already in flight.
// Shared memory: A 64 KB + W1 ring 6 x 8 KB + W2 ring 2 x 36 KB + h 2 x 16 KB + biases = 220 KB; TMEM 480 of 512 cols.
// Roofline: 89 GFLOP at DHD-S B=4 (160 000 pixels) -> 64 us at the sustained bf16 peak; weights are re-streamed from
// L2 per tile (544 KB / tile), which bounds a single-CTA design at ~65 us.
#include "common.cuh"
#include "tc_ptx.cuh"

namespace dhd {

namespace pt {
constexpr int kM = 128;            // pixels per tile (TMEM lanes)
constexpr int kK1 = 256;           // predictor out_dim (input of predicter[0])
constexpr int kN1 = 512;           // hidden width (2 * out_dim)
constexpr int kN2 = 288;           // Dz * n_cls = 16 * 18
constexpr int kDz = 16, kCls = 18;
constexpr int kChunk = 64;         // hidden columns per chunk = one 128-byte bf16 row
constexpr int kNChunks = kN1 / kChunk;               // 8
constexpr int kKc1 = kK1 / 64;                       // 4 K-chunks of A / W1
constexpr int kN2Half = kN2 / 2;                     // 144: two MMAs per k-step (N <= 256)
constexpr uint32_t kABytes = kM * 128;               // one K-chunk of A: 128 rows x 128 B
constexpr uint32_t kW1Bytes = kChunk * 128;          // 64 hidden rows x 128 B
constexpr uint32_t kW2Bytes = kN2 * 128;             // 288 logit rows x 128 B
constexpr uint32_t kHBytes = kM * 128;
constexpr int kW1Stages = 6, kW2Stages = 2, kHBufs = 2, kAcc1Bufs = 3;
constexpr uint32_t kOffA = 0;
constexpr uint32_t kOffW1 = kOffA + kKc1 * kABytes;
constexpr uint32_t kOffW2 = kOffW1 + kW1Stages * kW1Bytes;
constexpr uint32_t kOffH = kOffW2 + kW2Stages * kW2Bytes;
constexpr uint32_t kOffB1 = kOffH + kHBufs * kHBytes;
constexpr uint32_t kOffB2 = kOffB1 + kN1 * 4;
constexpr uint32_t kOffBar = kOffB2 + kN2 * 4;
// barriers (8 B each)
constexpr int kBarAFull = 0;                         // [4]  TMA -> MMA, one per K-chunk of A
constexpr int kBarAEmpty = kBarAFull + kKc1;         // [1]  MMA -> A producer
constexpr int kBarW1Full = kBarAEmpty + 1;           // [6]
constexpr int kBarW1Empty = kBarW1Full + kW1Stages;  // [6]
constexpr int kBarW2Full = kBarW1Empty + kW1Stages;  // [2]
constexpr int kBarW2Empty = kBarW2Full + kW2Stages;  // [2]
constexpr int kBarAcc1Full = kBarW2Empty + kW2Stages;    // [3]  MMA -> epilogue
constexpr int kBarAcc1Empty = kBarAcc1Full + kAcc1Bufs;  // [3]  epilogue (256 arrivals) -> MMA
constexpr int kBarHFull = kBarAcc1Empty + kAcc1Bufs;     // [2]  epilogue (256 arrivals) -> MMA
constexpr int kBarHEmpty = kBarHFull + kHBufs;           // [2]  MMA -> epilogue
constexpr int kBarAcc2Full = kBarHEmpty + kHBufs;        // [1]
constexpr int kBarAcc2Empty = kBarAcc2Full + 1;          // [1]  epilogue (256 arrivals) -> MMA
constexpr int kNumBars = kBarAcc2Empty + 1;
constexpr uint32_t kSmem = kOffBar + kNumBars * 8 + 16 + 1024;     // + tmem slot + alignment slack
constexpr int kEpiThreads = 256;                     // warps 0-7: softplus epilogue of the hidden chunks
constexpr int kFinThreads = 256;                     // warps 12-19: logits epilogue (argmax / store), off the chunk pipeline
constexpr int kThreads = kEpiThreads + 4 * 32 + kFinThreads;   // + G1 issuer, G2 issuer, weight producer, A producer
constexpr uint32_t kAcc2Col = kAcc1Bufs * kChunk;    // 192
static_assert(kAcc2Col + kN2 <= 512, "TMEM columns");
static_assert(kSmem <= 232448, "shared memory");
}  // namespace pt

struct TailParams {
  int M, HW, W, H;                 // pixels, pixels per image, image width / height
  int in_coff;
  int n_tiles;
  const float* b1;
  const float* b2;
  float* logits;                   // [(n * W + x) * H + y][288] (transpose_xy) or [m][288]; may be null
  uint8_t* occ;                    // same pixel order, [16] per pixel; may be null
  int transpose_xy;
  __nv_bfloat16* hidden;           // [m][hidden_ld] Softplus output (training), channel c at hidden_coff + c; may be null
  int hidden_ld, hidden_coff;
};

struct TailMaps {
  CUtensorMap a, w1, w2;
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld2(uint32_t taddr, float& a, float& b) {
  uint32_t r0, r1;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  a = __uint_as_float(r0);
  b = __uint_as_float(r1);
}

// torch.nn.Softplus(beta=1, threshold=20): max(x, 0) + log1p(exp(-|x|)); above 20 the second term is below half an
// ulp of x, so no branch is needed.  ex2 / lg2 are the MUFU approximations (abs error ~1e-7); the result is rounded to
// bf16 right after.
__device__ __forceinline__ float softplus_fast(float x) {
  float t, l;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(-fabsf(x) * 1.4426950408889634f));
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(1.f + t));
  return fmaf(0.6931471805599453f, l, fmaxf(x, 0.f));
}

__global__ void __launch_bounds__(pt::kThreads, 1)
predictor_tail_kernel(const __grid_constant__ TailMaps M, const __grid_constant__ TailParams P) {
  using namespace pt;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - raw);
  auto bar = [&](int i) { return base + kOffBar + 8u * (uint32_t)i; };
  float* s_b1 = reinterpret_cast<float*>(gen + kOffB1);
  float* s_b2 = reinterpret_cast<float*>(gen + kOffB2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + kOffBar + kNumBars * 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kMmaWarp = kEpiThreads / 32, kG2Warp = kMmaWarp + 1, kWWarp = kMmaWarp + 2, kAWarp = kMmaWarp + 3;
  constexpr int kFinWarp0 = kMmaWarp + 4;              // 12: a multiple of 4, so warp % 4 is again the TMEM lane quarter

  if (warp == kWWarp && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&M.a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&M.w1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&M.w2) : "memory");
    for (int i = 0; i < kNumBars; ++i) {
      const bool many = (i >= kBarAcc1Empty && i < kBarAcc1Empty + kAcc1Bufs) || (i >= kBarHFull && i < kBarHFull + kHBufs);
      mbar_init(bar(i), many ? (uint32_t)kEpiThreads : i == kBarAcc2Empty ? (uint32_t)kFinThreads : 1u);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = threadIdx.x; i < kN1; i += kThreads) s_b1[i] = P.b1 != nullptr ? __ldg(P.b1 + i) : 0.f;
  for (int i = threadIdx.x; i < kN2; i += kThreads) s_b2[i] = P.b2 != nullptr ? __ldg(P.b2 + i) : 0.f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int my_tiles = (P.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == kAWarp) {
    // ===================================================== G1 operand producer: the tile's A (128 x 256 pixels, once per
    // tile) and the W1 chunks.  W2 has its own producer warp: with one producer for both rings (first version) the
    // thread sat in the W2 ring's empty-wait -- G2 runs a chunk or two behind G1 by design -- and could not issue the W1
    // loads G1 was waiting for (ncu: 750 k retries on W2-empty, G1 and the producer ping-ponging on the W1 ring).
    if (lane == 0) {
      int i1 = 0;
      for (int lt = 0; lt < my_tiles; ++lt) {
        const int tile = blockIdx.x + lt * gridDim.x;
        mbar_wait(bar(kBarAEmpty), (uint32_t)(lt & 1) ^ 1u);
        for (int kc = 0; kc < kKc1; ++kc) {
          mbar_expect_tx(bar(kBarAFull + kc), kABytes);
          tma_load_2d(base + kOffA + kc * kABytes, &M.a, bar(kBarAFull + kc), P.in_coff + kc * 64, tile * kM);
        }
        for (int c = 0; c < kNChunks; ++c) {
          for (int kc = 0; kc < kKc1; ++kc, ++i1) {
            const int s = i1 % kW1Stages;
            mbar_wait(bar(kBarW1Empty + s), (uint32_t)((i1 / kW1Stages) & 1) ^ 1u);
            mbar_expect_tx(bar(kBarW1Full + s), kW1Bytes);
            tma_load_2d(base + kOffW1 + s * kW1Bytes, &M.w1, bar(kBarW1Full + s), kc * 64, c * kChunk);
          }
        }
      }
    }
  } else if (warp == kWWarp) {
    // ===================================================== W2 producer: one 288 x 64 chunk per hidden chunk
    if (lane == 0) {
      const int total = my_tiles * kNChunks;
      for (int g = 0; g < total; ++g) {
        const int c = g % kNChunks, s = g % kW2Stages;
        mbar_wait(bar(kBarW2Empty + s), (uint32_t)((g / kW2Stages) & 1) ^ 1u);
        mbar_expect_tx(bar(kBarW2Full + s), kW2Bytes);
        const uint32_t dst = base + kOffW2 + s * kW2Bytes;
        tma_load_2d(dst, &M.w2, bar(kBarW2Full + s), c * kChunk, 0);
        tma_load_2d(dst + kN2Half * 128, &M.w2, bar(kBarW2Full + s), c * kChunk, kN2Half);
      }
    }
  } else if (warp == kMmaWarp) {
    // ===================================================== G1 issuer: hidden chunk g = A x W1[c]^T, up to 3 chunks ahead
    if (lane == 0) {
      constexpr uint32_t kIdesc1 = umma_instr_desc_bf16(kM, kChunk);
      const int total = my_tiles * kNChunks;
      int s = 0, ab = 0, c = 0;
      uint32_t ws_par = 0, ab_par = 0, a_par = 0;            // phase parities of the W1 ring, the acc1 ring, the A tile
      for (int g = 0; g < total; ++g) {
        mbar_wait(bar(kBarAcc1Empty + ab), ab_par ^ 1u);
        tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(ab * kChunk);
#pragma unroll 1
        for (int kc = 0; kc < kKc1; ++kc) {
          if (c == 0) mbar_wait(bar(kBarAFull + kc), a_par);
          mbar_wait(bar(kBarW1Full + s), ws_par);
          tc_fence_after();
          const uint64_t da = umma_desc_sw128(base + kOffA + kc * kABytes);
          const uint64_t db = umma_desc_sw128(base + kOffW1 + s * kW1Bytes);
#pragma unroll
          for (int k = 0; k < 64 / kUmmaK; ++k)
            umma_bf16(tacc, da + 2u * k, db + 2u * k, kIdesc1, (kc == 0 && k == 0) ? 0u : 1u);
          umma_commit(bar(kBarW1Empty + s));
          if (++s == kW1Stages) { s = 0; ws_par ^= 1u; }
        }
        umma_commit(bar(kBarAcc1Full + ab));
        if (++ab == kAcc1Bufs) { ab = 0; ab_par ^= 1u; }
        if (++c == kNChunks) {
          c = 0;
          a_par ^= 1u;
          umma_commit(bar(kBarAEmpty));                      // every G1 of this tile has read A
        }
      }
    }
  } else if (warp == kG2Warp) {
    // ===================================================== G2 issuer: logits += h[c] x W2[:, c]^T, behind the epilogue
    if (lane == 0) {
      constexpr uint32_t kIdesc2 = umma_instr_desc_bf16(kM, kN2Half);
      const int total = my_tiles * kNChunks;
      const uint32_t tacc = tmem_base + kAcc2Col;
      int c = 0;
      uint32_t t_par = 0;                                    // tile parity (acc2)
      for (int g = 0; g < total; ++g) {
        const int hb = g & 1;                                // kHBufs == kW2Stages == 2
        const uint32_t par = (uint32_t)(g >> 1) & 1u;
        if (c == 0) mbar_wait(bar(kBarAcc2Empty), t_par ^ 1u);          // the previous tile's logits have been drained
        mbar_wait(bar(kBarHFull + hb), par);
        mbar_wait(bar(kBarW2Full + hb), par);
        tc_fence_after();
        const uint64_t da = umma_desc_sw128(base + kOffH + hb * kHBytes);
        const uint64_t db = umma_desc_sw128(base + kOffW2 + hb * kW2Bytes);
#pragma unroll
        for (int k = 0; k < 64 / kUmmaK; ++k) {
          const uint32_t acc = (c == 0 && k == 0) ? 0u : 1u;
          umma_bf16(tacc, da + 2u * k, db + 2u * k, kIdesc2, acc);
          umma_bf16(tacc + kN2Half, da + 2u * k, db + (uint64_t)((kN2Half * 128) >> 4) + 2u * k, kIdesc2, acc);
        }
        umma_commit(bar(kBarW2Empty + hb));
        umma_commit(bar(kBarHEmpty + hb));
        if (++c == kNChunks) {
          c = 0;
          t_par ^= 1u;
          umma_commit(bar(kBarAcc2Full));
        }
      }
    }
  } else if (warp < kMmaWarp) {
    // ===================================================== chunk epilogue: 8 warps, warp w owns TMEM lanes 32 * (w % 4) ..
    const int grp = warp >> 2;                     // 0 / 1: which half of a chunk
    const int row = (warp & 3) * 32 + lane;        // pixel row of the tile = TMEM lane
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    int g = 0;
    for (int lt = 0; lt < my_tiles; ++lt) {
      const long m_row = (long)(blockIdx.x + lt * gridDim.x) * kM + row;
      for (int c = 0; c < kNChunks; ++c, ++g) {
        const int ab = g % kAcc1Bufs, hb = g % kHBufs;
        mbar_wait(bar(kBarAcc1Full + ab), (uint32_t)((g / kAcc1Bufs) & 1));
        tc_fence_after();
        float v[32];
        tmem_ld32(lane_addr + (uint32_t)(ab * kChunk + grp * 32), v);
        tc_fence_before();
        mbar_arrive(bar(kBarAcc1Empty + ab));                      // accumulator chunk is in registers
        const float* b = s_b1 + c * kChunk + grp * 32;
        uint4 q[4];
        __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(q);
#pragma unroll
        for (int j = 0; j < 16; ++j)
          h2[j] = __floats2bfloat162_rn(softplus_fast(v[2 * j] + b[2 * j]), softplus_fast(v[2 * j + 1] + b[2 * j + 1]));
        if (P.hidden != nullptr && m_row < (long)P.M) {       // training: the backward needs the hidden layer (64 B per thread)
          uint4* hp = reinterpret_cast<uint4*>(P.hidden + m_row * P.hidden_ld + P.hidden_coff + c * kChunk + grp * 32);
#pragma unroll
          for (int j = 0; j < 4; ++j) hp[j] = q[j];
        }
        mbar_wait(bar(kBarHEmpty + hb), (uint32_t)((g / kHBufs) & 1) ^ 1u);     // G2(g - 2) has read this buffer
        // K-major SWIZZLE_128B operand tile: 16-byte chunk j of row r sits at chunk j ^ (r & 7)
        uint4* dst = reinterpret_cast<uint4*>(gen + kOffH + hb * kHBytes + row * 128);
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[(grp * 4 + j) ^ (row & 7)] = q[j];
        fence_proxy_async_smem();
        mbar_arrive(bar(kBarHFull + hb));
      }
    }
  } else {
    // ===================================================== logits epilogue: 8 warps, group g owns z planes
    // [8 * grp, 8 * grp + 8) = columns [144 * grp, +144) of the rows of its TMEM lane quarter
    const int fw = warp - kFinWarp0;
    const int grp = fw >> 2;
    const int row = (fw & 3) * 32 + lane;
    const uint32_t lane_addr = tmem_base + ((uint32_t)((fw & 3) * 32) << 16);
    for (int lt = 0; lt < my_tiles; ++lt) {
      const int tile = blockIdx.x + lt * gridDim.x;
      mbar_wait(bar(kBarAcc2Full), (uint32_t)(lt & 1));
      tc_fence_after();
      const long m = (long)tile * kM + row;
      const bool valid = m < (long)P.M;
      long pix = m;
      if (P.transpose_xy && valid) {                               // occ_pred.permute(0, 3, 2, 1): (n, y, x) -> (n, x, y)
        const int n = (int)(m / P.HW), rem = (int)(m - (long)n * P.HW);
        const int y = rem / P.W, x = rem - y * P.W;
        pix = ((long)n * P.W + x) * P.H + y;
      }
      float* lo = P.logits != nullptr && valid ? P.logits + pix * kN2 + grp * kN2Half : nullptr;
      const float* b2 = s_b2 + grp * kN2Half;
      const uint32_t t2 = lane_addr + kAcc2Col + (uint32_t)(grp * kN2Half);
      uint32_t occ_lo = 0, occ_hi = 0, tie_mask = 0;
      float best = 0.f, second = 0.f;
      int arg = 0;
#pragma unroll
      for (int blk = 0; blk < kN2Half / 16; ++blk) {               // 9 x 16 columns
        float v[16];
        tmem_ld16(t2 + blk * 16, v);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int col = blk * 16 + j;                            // compile-time: 0 .. 143
          const int k = col % kCls, z = col / kCls;
          v[j] += b2[col];
          if (k == 0) {
            best = v[j];
            second = -INFINITY;
            arg = 0;
          } else {
            second = fmaxf(second, fminf(v[j], best));
            if (v[j] > best) {                                     // strict: the first maximum wins, as torch.argmax
              best = v[j];
              arg = k;
            }
          }
          if (k == kCls - 1) {
            if (z < 4) occ_lo |= (uint32_t)arg << (8 * z);
            else occ_hi |= (uint32_t)arg << (8 * (z - 4));
            if (best - second <= kSoftmaxTieGap) tie_mask |= 1u << z;   // softmax rounding may merge the two: exact path
          }
        }
        if (lo != nullptr) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            st_cs(reinterpret_cast<float4*>(lo + blk * 16) + j, make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]));
        }
      }
      if (P.occ != nullptr && __any_sync(kFull, tie_mask != 0)) {
        // rare (top two logits of a z plane within 1e-6): re-read that plane's 18 logits and let the restatement of
        // torch's softmax kernel pick the class.  Warp-uniform control flow: tcgen05.ld is .sync.aligned.
#pragma unroll 1
        for (int z = 0; z < kDz / 2; ++z) {
          if (!__any_sync(kFull, (tie_mask >> z) & 1u)) continue;
          float x[kCls];
#pragma unroll
          for (int i = 0; i < kCls / 2; ++i) {
            tmem_ld2(t2 + (uint32_t)(z * kCls + 2 * i), x[2 * i], x[2 * i + 1]);
            x[2 * i] += b2[z * kCls + 2 * i];
            x[2 * i + 1] += b2[z * kCls + 2 * i + 1];
          }
          if ((tie_mask >> z) & 1u) {
            const uint32_t a = (uint32_t)softmax_argmax_torch<kCls>(x);
            if (z < 4) occ_lo = (occ_lo & ~(0xffu << (8 * z))) | (a << (8 * z));
            else occ_hi = (occ_hi & ~(0xffu << (8 * (z - 4)))) | (a << (8 * (z - 4)));
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar(kBarAcc2Empty));
      if (P.occ != nullptr && valid)
        *reinterpret_cast<uint2*>(P.occ + pix * kDz + grp * 8) = make_uint2(occ_lo, occ_hi);
    }
  }
  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

typedef CUresult (*EncodeTiledFnT)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
void* conv_encode_fn();            // conv_igemm.cu

static int encode_2d(EncodeTiledFnT enc, CUtensorMap* map, const void* ptr, uint64_t inner, uint64_t rows, uint64_t row_elems,
                     uint32_t box_rows, const char* what) {
  cuuint64_t dims[2] = {inner, rows};
  cuuint64_t strides[1] = {row_elems * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t es[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(DHD_EINVAL, "cuTensorMapEncodeTiled(%s) failed: %ld", what, (long)r);
  return DHD_OK;
}

}  // namespace dhd

using namespace dhd;

extern "C" int dhd_predictor_tail(const dhd_predictor_tail_desc* d, void* stream) {
  DHD_REQUIRE(d != nullptr, "predictor tail desc is null");
  DHD_REQUIRE(d->in != nullptr && d->w1 != nullptr && d->w2 != nullptr, "null input / weight pointer");
  DHD_REQUIRE(d->logits != nullptr || d->occ != nullptr, "neither logits nor occ requested");
  DHD_REQUIRE(d->N > 0 && d->H > 0 && d->W > 0 && (long)d->N * d->H * d->W < (1L << 31) - 256, "bad image shape");
  DHD_REQUIRE(d->K1 == pt::kK1 && d->N1 == pt::kN1 && d->Dz == pt::kDz && d->n_cls == pt::kCls,
              "fused tail is built for 256 -> 512 -> 16 x 18 (DHD-S/M/L occ_head); use the layer-by-layer path otherwise");
  DHD_REQUIRE(d->in_ld % 8 == 0 && d->in_coff % 8 == 0 && d->in_coff + pt::kK1 <= d->in_ld,
              "input channel offsets must be multiples of 8 (16-byte TMA alignment)");
  DHD_REQUIRE(((uintptr_t)d->in & 15) == 0 && ((uintptr_t)d->w1 & 15) == 0 && ((uintptr_t)d->w2 & 15) == 0,
              "input / weights must be 16-byte aligned");
  DHD_REQUIRE(d->logits == nullptr || ((uintptr_t)d->logits & 15) == 0, "logits must be 16-byte aligned");
  DHD_REQUIRE(d->occ == nullptr || ((uintptr_t)d->occ & 7) == 0, "occ must be 8-byte aligned");
  EncodeTiledFnT enc = (EncodeTiledFnT)conv_encode_fn();
  if (enc == nullptr) return fail(DHD_EUNSUPPORTED, "%s", "cuTensorMapEncodeTiled is unavailable");
  TailMaps maps;
  TailParams P;
  P.M = d->N * d->H * d->W;
  P.HW = d->H * d->W;
  P.W = d->W;
  P.H = d->H;
  P.in_coff = d->in_coff;
  P.n_tiles = (P.M + pt::kM - 1) / pt::kM;
  P.b1 = d->b1;
  P.b2 = d->b2;
  P.logits = d->logits;
  P.occ = d->occ;
  P.transpose_xy = d->transpose_xy;
  P.hidden = (__nv_bfloat16*)d->hidden;
  P.hidden_ld = d->hidden_ld;
  P.hidden_coff = d->hidden_coff;
  if (d->hidden != nullptr)
    DHD_REQUIRE(d->hidden_ld % 8 == 0 && d->hidden_coff % 8 == 0 && d->hidden_coff + pt::kN1 <= d->hidden_ld &&
                    ((uintptr_t)d->hidden & 15) == 0, "hidden rows must be 16-byte aligned and hold N1 channels");
  int rc = encode_2d(enc, &maps.a, d->in, (uint64_t)d->in_ld, (uint64_t)P.M, (uint64_t)d->in_ld, pt::kM, "A");
  if (rc != DHD_OK) return rc;
  rc = encode_2d(enc, &maps.w1, d->w1, pt::kK1, pt::kN1, pt::kK1, pt::kChunk, "W1");
  if (rc != DHD_OK) return rc;
  rc = encode_2d(enc, &maps.w2, d->w2, pt::kN1, pt::kN2, pt::kN1, pt::kN2Half, "W2");
  if (rc != DHD_OK) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(predictor_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pt::kSmem);
    if (e != cudaSuccess) return fail((int)e, "%s: %ld", "cudaFuncSetAttribute(predictor_tail)", (long)e);
    attr_set = true;
  }
  const int grid = min(P.n_tiles, sm_count());
  predictor_tail_kernel<<<grid, pt::kThreads, pt::kSmem, (cudaStream_t)stream>>>(maps, P);
  DHD_CUDA_LAUNCH_CHECK("predictor_tail");
  return DHD_OK;
}
