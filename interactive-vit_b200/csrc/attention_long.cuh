// Multi-head self-attention for sequences / head sizes outside the fused single-tile kernel (attention.cuh):
// any token count N (keys are processed in blocks of 128) and head dims D = 64 .. 128 in steps of 16 (ViT-H: 80).
// Used by the 384 px models of BASELINE.json (577 tokens) and by the 1280-wide "ViT-H" configuration.
//
// Arithmetic: torch.nn.functional.multi_head_attention_forward, weights branch (torch/nn/functional.py:6630-6659):
// P = softmax((q / sqrt(D)) k^T), O = P v, optional head mean of P -- the same contract as attention.cuh.
//
// Two kernels, both tcgen05 / TMEM / TMA, 384 threads (warp 0 TMA, warp 1 MMA issue, warp 2 TMEM allocator, warps 4..11
// softmax with two threads per query row = two 64-key halves of a 128-key block):
//
//   attention_long_ctx_kernel   one CTA per (image, 128-query tile, head).  Pass A recomputes nothing but the row
//       maximum (S = Q K^T per key block, max only); pass B recomputes S, forms e = exp2((s - max) c), accumulates the
//       row sum and O += bf16(e) V in TMEM.  No online rescaling: the maximum is final before the first exponential.
//       Writes the context rows (O / sum) and the row statistics (max * c, 1 / sum) for the map kernel.
//
//   attention_long_maps_kernel  one CTA per (image, 128-query tile, 128-key block), looping over the heads:
//       S = Q_h K_h^T, p = exp2(s c - max c) / sum from the stored statistics; the head average accumulates in
//       REGISTERS (64 fp32 per thread) and is written once; per-head class-token rows and (opt-in) full per-head maps
//       are written as they are produced.  Only launched when a map output is requested.
//
// Head dims above 64 use two 64-column TMA boxes per operand tile (the second box's surplus columns belong to the
// next head and are simply not multiplied: the MMAs step over K = D in units of 16).
#pragma once
#include <cuda.h>
#include "ptx.cuh"

namespace vitb200 {

struct AttnLongParams {
  int B, N, H, D;       // images, tokens, heads, head dim (multiple of 16, 64..128)
  int d;                // model width = H * D
  int q_tiles, k_blocks;  // ceil(N / 128) each
  float scale_log2;     // (1 / sqrt(D)) * log2(e)
  __nv_bfloat16* ctx;   // [B*N, d]
  __nv_bfloat16* ctx_lo;  // fp32x3 mode: low halves of the context (nullptr otherwise)
  float2* stats;        // [B, H, N] (max * scale_log2, 1 / sum)
  float* avg_map;       // [B, N, ldmap] or nullptr
  float* head_map;      // [B, H, N, ldmap] or nullptr
  float* cls_map;       // [B, H, N] or nullptr
  int ldmap;
};

namespace attn_long_cfg {
constexpr int kThreads = 384;
constexpr int BM = 128;                     // query rows per CTA
constexpr int BK = 128;                     // keys per block
constexpr int kBoxBytes = 128 * 128;        // one [128 rows x 64 bf16] SWIZZLE_128B box
constexpr int kTileBytes = 2 * kBoxBytes;   // operand tile: columns [0,64) and [64,128) of the head
constexpr int kTmemS = 0;                   // two S buffers of 128 fp32 columns
constexpr int kTmemO = 256;                 // O: up to 128 columns
// ctx kernel: Q + 2 K stages + 2 V stages + P = 32 * 6 = 192 KB; maps kernel: 2 Q stages + 2 K stages = 128 KB
constexpr int kSmemCtx = 6 * kTileBytes + 2 * 2 * BM * 4 + 256;
// fp32x3 mode (head dim 64 only): every 32 KB operand slot holds [hi box | lo box]; one more 32 KB tile for P_lo
constexpr int kSmemCtxSplit = 7 * kTileBytes + 2 * 2 * BM * 4 + 256;
constexpr int kSmemMaps = 4 * kTileBytes + 256;
// compact context kernel (head dim 64, bf16 mode): one 16 KB box per operand, ONE V stage, one S buffer -> 96 KB of
// smem and 192 TMEM columns, so that TWO CTAs share an SM and one CTA's exponential pass runs under the other's
// loads, maxima and epilogue (the per-CTA chain is serial: ~22k cycles of which the SM's pipes are busy a third)
// compact map kernel (head dim 64, bf16 mode): 64-key blocks -> 32 accumulators per thread, 2 x (16 + 8) KB of smem and
// 128 TMEM columns: two CTAs per SM here too
constexpr int kSmemMapsCompact = 2 * kBoxBytes + 2 * (kBoxBytes / 2) + 256;
constexpr int kSmemCtxCompact = 4 * kBoxBytes + 2 * kBoxBytes + 2 * 2 * BM * 4 + 256;   // Q + 2 K + V, P (32 KB)
}  // namespace attn_long_cfg

__device__ __forceinline__ uint32_t pack_bf16x2_f(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// S[128 x 128] (+)= Q[128 x D] K[128 x D]^T from two-box operand tiles: K = 64 from box 0, D - 64 from box 1.
__device__ __forceinline__ void issue_qk_long(uint32_t tmem_s, uint32_t sq, uint32_t sk, int D, int keys = attn_long_cfg::BK) {
  using namespace attn_long_cfg;
  const uint32_t idesc = ptx::make_idesc_bf16(BM, static_cast<uint32_t>(keys), 0, 0);
  const uint64_t dq = ptx::make_smem_desc_sw128(sq, 16, 1024);
  const uint64_t dk = ptx::make_smem_desc_sw128(sk, 16, 1024);
  const int ksteps = D >> 4;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (k < ksteps) {
      // box 1 starts kBoxBytes (= 1024 descriptor units) after box 0; +2 units per 16 columns inside a box
      const uint64_t off = static_cast<uint64_t>((k >> 2) * (kBoxBytes >> 4) + 2 * (k & 3));
      ptx::umma_bf16_ss(tmem_s, dq + off, dk + off, idesc, k != 0 ? 1u : 0u);
    }
  }
}

// fp32x3 mode, D = 64: operand slots are [hi box | lo box]; S = Q_hi K_hi^T + Q_lo K_hi^T + Q_hi K_lo^T.
__device__ __forceinline__ void issue_qk_long_split(uint32_t tmem_s, uint32_t sq, uint32_t sk) {
  using namespace attn_long_cfg;
  const uint32_t idesc = ptx::make_idesc_bf16(BM, BK, 0, 0);
  const uint64_t dq = ptx::make_smem_desc_sw128(sq, 16, 1024);
  const uint64_t dk = ptx::make_smem_desc_sw128(sk, 16, 1024);
  constexpr uint64_t kLo = kBoxBytes >> 4;
#pragma unroll
  for (int g = 0; g < 3; ++g) {
    const uint64_t oq = g == 1 ? kLo : 0, ok = g == 2 ? kLo : 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      ptx::umma_bf16_ss(tmem_s, dq + oq + 2 * k, dk + ok + 2 * k, idesc, (g | k) != 0 ? 1u : 0u);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
template <bool kSplit, bool kCompact = false, bool kOnline = true>
__global__ void __launch_bounds__(attn_long_cfg::kThreads, kCompact ? 2 : 1)
attention_long_ctx_kernel(const __grid_constant__ CUtensorMap tmap_qkv,  // box 64 x 128 over qkv viewed as [B][N][3d]
                          const __grid_constant__ CUtensorMap tmap_qkv_lo,  // kSplit: the low halves, same geometry
                          AttnLongParams p) {
  using namespace attn_long_cfg;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();
  static_assert(!(kSplit && kCompact), "the compact layout is for plain bf16 operands");
  constexpr int kOp = kCompact ? kBoxBytes : kTileBytes;   // bytes of one operand tile slot
  constexpr int kVStages = kCompact ? 1 : 2;
  constexpr int kSBufs = kCompact ? 1 : 2;                 // S buffers in TMEM
  constexpr uint32_t kTmemOc = kCompact ? 128 : kTmemO;    // O columns behind the S buffer(s)
  // kOnline: one O accumulator per 64-key half of the blocks (the two threads of a row keep independent reference
  // maxima; the halves meet in the epilogue), the second one kOStride columns after the first
  constexpr uint32_t kOStride = kCompact ? 64 : 128;
  uint8_t* s_q = smem;
  uint8_t* s_k = smem + kOp;                       // 2 stages
  uint8_t* s_v = smem + 3 * kOp;                   // kVStages stages
  uint8_t* s_p = smem + (3 + kVStages) * kOp;      // 2 K-blocks of 64 keys: [128 rows x 128 B] each
  uint8_t* s_p_lo = s_p + kTileBytes;              // kSplit only
  float* red = reinterpret_cast<float*>(s_p + (kSplit ? 2 : 1) * kTileBytes);  // [2 kinds][2 halves][128 rows]
  uint64_t* bars = reinterpret_cast<uint64_t*>(red + 2 * 2 * BM);
  uint64_t* q_full = bars;          // Q landed
  uint64_t* k_full = bars + 1;      // [2]
  uint64_t* k_empty = bars + 3;     // [2]
  uint64_t* v_full = bars + 5;      // [2]
  uint64_t* v_empty = bars + 7;     // [2]
  uint64_t* s_full = bars + 9;      // [2] S buffer written by the MMAs
  uint64_t* s_free = bars + 11;     // [2] S buffer copied to registers (8 warps)
  uint64_t* p_full = bars + 13;     // P tile written (8 warps)
  uint64_t* p_free = bars + 14;     // P V retired
  uint64_t* o_full = bars + 15;     // all P V retired
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x % p.H;
  const int qt = (blockIdx.x / p.H) % p.q_tiles;
  const int b = blockIdx.x / (p.H * p.q_tiles);
  const int nb = p.k_blocks;
  const int D = p.D;
  const bool two_box = !kSplit && !kCompact && D > 64;
  const uint32_t tile_tx = (kSplit || two_box) ? kTileBytes : kBoxBytes;

  if (warp == 0 && lane == 0) ptx::prefetch_tmap(&tmap_qkv);
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&k_full[i], 1), ptx::mbar_init(&k_empty[i], 1);
      ptx::mbar_init(&v_full[i], 1), ptx::mbar_init(&v_empty[i], 1);
      ptx::mbar_init(&s_full[i], 1), ptx::mbar_init(&s_free[i], 8);
    }
    ptx::mbar_init(p_full, 8), ptx::mbar_init(p_free, 1), ptx::mbar_init(o_full, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc<kCompact ? 256 : 512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    // rows beyond the image (3-D map: [B][N][3d]) are zero-filled: S = 0 there, masked by the softmax threads
    auto load_tile = [&](uint8_t* dst, uint64_t* bar, int col, int row) {
      ptx::mbar_arrive_expect_tx(bar, tile_tx);
      ptx::tma_load_3d(dst, &tmap_qkv, bar, col, row, b);
      if (kSplit) ptx::tma_load_3d(dst + kBoxBytes, &tmap_qkv_lo, bar, col, row, b);
      else if (two_box) ptx::tma_load_3d(dst + kBoxBytes, &tmap_qkv, bar, col + 64, row, b);
    };
    if (ptx::elect_one()) load_tile(s_q, q_full, h * D, qt * BM);
    __syncwarp();
    int it = 0;  // K stage uses: pass A blocks (not kOnline) then pass B blocks
    for (int pass = kOnline ? 1 : 0; pass < 2; ++pass) {
      for (int blk = 0; blk < nb; ++blk, ++it) {
        const int st = it & 1;
        ptx::mbar_wait(&k_empty[st], ((it >> 1) & 1) ^ 1);
        if (ptx::elect_one()) load_tile(s_k + st * kOp, &k_full[st], p.d + h * D, blk * BK);
        __syncwarp();
        if (pass == 1) {
          const int vs = kVStages == 2 ? (blk & 1) : 0;
          const uint32_t vph = kVStages == 2 ? ((blk >> 1) & 1) : (blk & 1);
          ptx::mbar_wait(&v_empty[vs], vph ^ 1);
          if (ptx::elect_one()) load_tile(s_v + vs * kOp, &v_full[vs], 2 * p.d + h * D, blk * BK);
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ UMMA issuer
    const uint32_t sq = ptx::smem_u32(s_q);
    ptx::mbar_wait(q_full, 0);
    auto issue_qk = [&](int it) {  // it: running S / K use index over both passes
      const int st = it & 1;                        // K stage
      const int sb = kSBufs == 2 ? st : 0;          // S buffer
      ptx::mbar_wait(&k_full[st], (it >> 1) & 1);
      if (kSBufs == 2) {
        if (it >= 2) ptx::mbar_wait(&s_free[sb], ((it - 2) >> 1) & 1);
      } else {
        if (it >= 1) ptx::mbar_wait(&s_free[0], (it - 1) & 1);   // the single buffer: the previous block has been read
      }
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        if (kSplit) issue_qk_long_split(tmem_base + kTmemS + sb * BK, sq, ptx::smem_u32(s_k + st * kOp));
        else issue_qk_long(tmem_base + kTmemS + sb * BK, sq, ptx::smem_u32(s_k + st * kOp), D);
        ptx::umma_commit(&k_empty[st]);
        ptx::umma_commit(&s_full[sb]);
      }
      __syncwarp();
    };
    // pass A: maxima only
    const int it0 = kOnline ? 0 : nb;   // S / K use index of pass B's first block
    if (!kOnline)
      for (int blk = 0; blk < nb; ++blk) issue_qk(blk);
    // pass B: Q K^T of block blk+1 is issued before P V of block blk
    issue_qk(it0);
    const uint32_t idesc_pv0 = ptx::make_idesc_bf16(BM, 64, 0, 1);                       // V columns [0, 64)
    const uint32_t idesc_pv1 = ptx::make_idesc_bf16(BM, static_cast<uint32_t>(two_box ? D - 64 : 16), 0, 1);
    for (int blk = 0; blk < nb; ++blk) {
      if (blk + 1 < nb) issue_qk(it0 + blk + 1);
      const int vs = kVStages == 2 ? (blk & 1) : 0;
      ptx::mbar_wait(&v_full[vs], kVStages == 2 ? ((blk >> 1) & 1) : (blk & 1));
      ptx::mbar_wait(p_full, blk & 1);
      ptx::tc_fence_after();
      const uint64_t dp0 = ptx::make_smem_desc_sw128(ptx::smem_u32(s_p), 16, 1024);
      const uint64_t dv0 = ptx::make_smem_desc_sw128(ptx::smem_u32(s_v + vs * kOp), 1024, 1024);
      if (ptx::elect_one()) {
#pragma unroll
        for (int ks = 0; ks < BK / 16; ++ks) {
          // A: P K-block ks / 4 (+16 KB), +32 B per 16 keys; B: V rows [16 ks, +16) MN-major, +2 KB per 16 keys
          const uint64_t dp = dp0 + static_cast<uint64_t>((ks >> 2) * (kBoxBytes >> 4) + 2 * (ks & 3));
          const uint64_t dv = dv0 + static_cast<uint64_t>(ks * (2048 >> 4));
          // kOnline: keys [0, 64) of the block (ks < 4) accumulate into O_0, keys [64, 128) into O_1
          const uint32_t acc = (kOnline ? (blk | (ks & 3)) : (blk | ks)) != 0 ? 1u : 0u;
          const uint32_t t_o = tmem_base + kTmemOc + (kOnline ? (ks >> 2) * kOStride : 0u);
          ptx::umma_bf16_ss(t_o, dp, dv, idesc_pv0, acc);
          if (kSplit) {
            // + P_lo V_hi + P_hi V_lo (the P_lo tile sits one 32 KB slot after P_hi, V_lo one box after V_hi)
            ptx::umma_bf16_ss(t_o, dp + (kTileBytes >> 4), dv, idesc_pv0, 1u);
            ptx::umma_bf16_ss(t_o, dp, dv + (kBoxBytes >> 4), idesc_pv0, 1u);
          } else if (two_box) {
            ptx::umma_bf16_ss(t_o + 64, dp, dv + (kBoxBytes >> 4), idesc_pv1, acc);
          }
        }
        ptx::umma_commit(&v_empty[vs]);
        ptx::umma_commit(p_free);
        if (blk + 1 == nb) ptx::umma_commit(o_full);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ softmax
    const int half = (warp - 4) >> 2, quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int qrow = qt * BM + r;
    const bool row_ok = qrow < p.N;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const uint32_t p_row = ptx::smem_u32(s_p) + half * kBoxBytes + r * 128;
    const uint32_t p_row_lo = ptx::smem_u32(s_p_lo) + half * kBoxBytes + r * 128;
    const int sw = r & 7;
    float* red_max = red;
    float* red_sum = red + 2 * BM;
    const int it0s = kOnline ? 0 : nb;   // S use index of pass B's first block

    uint32_t s[4][16];
    // S block `it` -> registers; release = hand the S buffer back to the MMA issuer
    auto read_s = [&](int it, bool wait) {
      const int st = kSBufs == 2 ? (it & 1) : 0;
      if (wait) ptx::mbar_wait(&s_full[st], kSBufs == 2 ? ((it >> 1) & 1) : (it & 1));
      ptx::tc_fence_after();
#pragma unroll
      for (int c = 0; c < 4; ++c) ptx::tmem_ld_x16(lane_base + kTmemS + st * BK + half * 64 + c * 16, s[c]);
      ptx::tmem_ld_wait();
    };
    auto release_s = [&](int it) {
      const int st = kSBufs == 2 ? (it & 1) : 0;
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&s_free[st]);
    };
    auto load_s = [&](int it) { read_s(it, true), release_s(it); };
    auto mask_s = [&](int blk) {
      const int key0 = blk * BK + half * 64;
      if (key0 + 64 > p.N) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (key0 + c * 16 + j >= p.N) s[c][j] = 0xff800000u;
      }
    };
    float mxs, sum = 0.f;   // reference maximum (scaled: log2 units) and the row sum relative to it
    float a_self = 1.f, a_oth = 0.f;   // kOnline: weights of this half's / the other half's O accumulator in the epilogue
    if (!kOnline) {
      // ---- pass A: row maximum
      float mx = -INFINITY;
      for (int blk = 0; blk < nb; ++blk) {
        load_s(blk);
        mask_s(blk);
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int j = 0; j < 16; j += 2) mx = ptx::fmax3(mx, __uint_as_float(s[c][j]), __uint_as_float(s[c][j + 1]));
      }
      red_max[half * BM + r] = mx;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mx = fmaxf(mx, red_max[(half ^ 1) * BM + r]);
      mxs = mx * p.scale_log2;
    } else {
      mxs = -INFINITY;
    }
    // ---- pass B: e = exp2(s c - ref) -> row sum, bf16 P tile -> O += P V.  Without pass A (kOnline) `ref` is this
    //      THREAD's running reference: the maximum of its 64-key half of the first block, raised only when a later block
    //      exceeds it by more than kLazy (2^8: e stays far inside bf16 / fp32 range), in which case the thread rescales
    //      its own row of its own O accumulator in TMEM between the P V of the previous block and its next P tile.
    constexpr float kLazy = 8.0f;
    for (int blk = 0; blk < nb; ++blk) {
      if (kOnline) read_s(it0s + blk, true);   // the buffer is handed back below: the rescale path reads the block twice
      else load_s(it0s + blk);
      mask_s(blk);
      if (kOnline) {
        float bm = -INFINITY;
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int j = 0; j < 16; j += 2) bm = ptx::fmax3(bm, __uint_as_float(s[c][j]), __uint_as_float(s[c][j + 1]));
        const float bms = bm * p.scale_log2;
        const bool raise = bms > mxs + kLazy;   // also true for the first unmasked block of the thread (mxs = -inf)
        if (blk > 0 && __any_sync(0xffffffffu, raise)) {
          // (warp-uniform branch: tcgen05.ld / st are warp-collective; lanes that keep their reference multiply by 1)
          const float f = raise ? ptx::ex2_approx(mxs - bms) : 1.0f;   // exp2(-inf) = 0: nothing accumulated yet
          ptx::mbar_wait(p_free, (blk - 1) & 1);   // O is complete up to the previous block
          ptx::tc_fence_after();
          const uint32_t t_o = lane_base + kTmemOc + half * kOStride;
          for (int c0 = 0; c0 < D; c0 += 16) {
            uint32_t o[16];
            ptx::tmem_ld_x16(t_o + c0, o);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = __float_as_uint(__uint_as_float(o[j]) * f);
            ptx::tmem_st_x16(t_o + c0, o);
          }
          ptx::tmem_st_wait();
          ptx::tc_fence_before();
          sum *= f;
          // (the block is read again instead of being kept in 64 registers across this branch: with it live here the
          // compact instantiation -- 80 registers, two CTAs per SM -- parked the whole block in local memory on EVERY
          // iteration)
          read_s(it0s + blk, false);
          mask_s(blk);
        }
        release_s(it0s + blk);
        if (raise) mxs = bms;
      }
      if (blk > 0) ptx::mbar_wait(p_free, (blk - 1) & 1);  // the previous block's P V has read the tile
      const float ref = (kOnline && mxs == -INFINITY) ? 0.f : mxs;   // a thread whose keys are all masked so far: e = 0
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        float e[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          e[j] = ptx::ex2_approx(fmaf(__uint_as_float(s[c][j]), p.scale_log2, -ref));
          sum += e[j];
        }
        // keys [16 c, 16 c + 16) of this half: 16-byte chunks 2c, 2c+1 of the row in K-block `half`
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p_row + (((2 * c) ^ sw) << 4)),
                     "r"(pack_bf16x2_f(e[0], e[1])), "r"(pack_bf16x2_f(e[2], e[3])), "r"(pack_bf16x2_f(e[4], e[5])),
                     "r"(pack_bf16x2_f(e[6], e[7]))
                     : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p_row + (((2 * c + 1) ^ sw) << 4)),
                     "r"(pack_bf16x2_f(e[8], e[9])), "r"(pack_bf16x2_f(e[10], e[11])), "r"(pack_bf16x2_f(e[12], e[13])),
                     "r"(pack_bf16x2_f(e[14], e[15]))
                     : "memory");
        if (kSplit) {
          float l[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) l[j] = e[j] - __bfloat162float(__float2bfloat16_rn(e[j]));
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p_row_lo + (((2 * c) ^ sw) << 4)),
                       "r"(pack_bf16x2_f(l[0], l[1])), "r"(pack_bf16x2_f(l[2], l[3])), "r"(pack_bf16x2_f(l[4], l[5])),
                       "r"(pack_bf16x2_f(l[6], l[7]))
                       : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(p_row_lo + (((2 * c + 1) ^ sw) << 4)),
                       "r"(pack_bf16x2_f(l[8], l[9])), "r"(pack_bf16x2_f(l[10], l[11])), "r"(pack_bf16x2_f(l[12], l[13])),
                       "r"(pack_bf16x2_f(l[14], l[15]))
                       : "memory");
        }
      }
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(p_full);
    }
    float inv;
    if (!kOnline) {
      red_sum[half * BM + r] = sum;
      asm volatile("bar.sync 2, 256;" ::: "memory");
      sum += red_sum[(half ^ 1) * BM + r];
      inv = 1.0f / sum;
    } else {
      // the two halves of the row meet: common reference M = max of the two, each half's sum and O weighted by
      // 2^(own reference - M).  Formed from (half 0, half 1) in that order by both threads: the same bits in both.
      red_max[half * BM + r] = mxs, red_sum[half * BM + r] = sum;
      asm volatile("bar.sync 2, 256;" ::: "memory");
      const float m0 = red_max[r], m1 = red_max[BM + r], s0 = red_sum[r], s1 = red_sum[BM + r];
      const float M = fmaxf(m0, m1);   // finite: key 0 is never masked
      const float a0 = m0 == -INFINITY ? 0.f : ptx::ex2_approx(m0 - M), a1 = m1 == -INFINITY ? 0.f : ptx::ex2_approx(m1 - M);
      inv = 1.0f / __fadd_rn(__fmul_rn(s0, a0), __fmul_rn(s1, a1));
      a_self = half == 0 ? a0 : a1, a_oth = half == 0 ? a1 : a0;
      mxs = M;
    }
    if (half == 0 && row_ok) p.stats[(static_cast<size_t>(b) * p.H + h) * p.N + qrow] = make_float2(mxs, inv);
    // ---- context rows: this half owns D / 2 of the D columns (D / 2 is a multiple of 8)
    ptx::mbar_wait(o_full, 0);
    ptx::tc_fence_after();
    const int cols = D >> 1;
    __nv_bfloat16* op = p.ctx + (static_cast<size_t>(b) * p.N + qrow) * p.d + h * D + half * cols;
    const float w_self = a_self * inv, w_oth = a_oth * inv;
    const float inv_e = kOnline ? 1.0f : inv;   // kOnline: 1 / sum is part of the weights
    for (int c0 = 0; c0 < cols; c0 += 8) {
      uint32_t o[8];
      if (!kOnline) {
        ptx::tmem_ld_x8(lane_base + kTmemOc + half * cols + c0, o);
        ptx::tmem_ld_wait();
      } else {
        // O = (O_half0 2^(m0 - M) + O_half1 2^(m1 - M)) / sum; the scaling by inv below is folded into the weights
        uint32_t o2[8];
        ptx::tmem_ld_x8(lane_base + kTmemOc + half * kOStride + half * cols + c0, o);
        ptx::tmem_ld_x8(lane_base + kTmemOc + (half ^ 1) * kOStride + half * cols + c0, o2);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j)
          o[j] = __float_as_uint(fmaf(__uint_as_float(o2[j]), w_oth, __uint_as_float(o[j]) * w_self));
      }
      if (row_ok) {
        uint4 v;
        v.x = pack_bf16x2_f(__uint_as_float(o[0]) * inv_e, __uint_as_float(o[1]) * inv_e);
        v.y = pack_bf16x2_f(__uint_as_float(o[2]) * inv_e, __uint_as_float(o[3]) * inv_e);
        v.z = pack_bf16x2_f(__uint_as_float(o[4]) * inv_e, __uint_as_float(o[5]) * inv_e);
        v.w = pack_bf16x2_f(__uint_as_float(o[6]) * inv_e, __uint_as_float(o[7]) * inv_e);
        *reinterpret_cast<uint4*>(op + c0) = v;
        if (kSplit) {
          const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
          uint4 l;
          l.x = pack_bf16x2_f(__uint_as_float(o[0]) * inv_e - __low2float(h[0]), __uint_as_float(o[1]) * inv_e - __high2float(h[0]));
          l.y = pack_bf16x2_f(__uint_as_float(o[2]) * inv_e - __low2float(h[1]), __uint_as_float(o[3]) * inv_e - __high2float(h[1]));
          l.z = pack_bf16x2_f(__uint_as_float(o[4]) * inv_e - __low2float(h[2]), __uint_as_float(o[5]) * inv_e - __high2float(h[2]));
          l.w = pack_bf16x2_f(__uint_as_float(o[6]) * inv_e - __low2float(h[3]), __uint_as_float(o[7]) * inv_e - __high2float(h[3]));
          *reinterpret_cast<uint4*>(p.ctx_lo + (op - p.ctx) + c0) = l;
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<kCompact ? 256 : 512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// kCompact (head dim 64, plain bf16): key blocks of 64 (p.k_blocks counts THOSE), K tiles through `tmap_qkv_lo`, which
// the host then builds as a 64-row box map of the same tensor.
template <bool kHeads, bool kSplit, bool kCompact = false>
__global__ void __launch_bounds__(attn_long_cfg::kThreads, kCompact ? 2 : 1)
attention_long_maps_kernel(const __grid_constant__ CUtensorMap tmap_qkv, const __grid_constant__ CUtensorMap tmap_qkv_lo,
                           AttnLongParams p) {
  using namespace attn_long_cfg;
  static_assert(!(kSplit && kCompact), "the compact layout is for plain bf16 operands");
  constexpr int kKeys = kCompact ? 64 : BK;                    // keys per block
  constexpr int kQb = kCompact ? kBoxBytes : kTileBytes;       // bytes of a Q stage
  constexpr int kKb = kCompact ? kBoxBytes / 2 : kTileBytes;   // bytes of a K stage
  constexpr int kPer = kKeys / 2;                              // keys per thread (two threads per row)
  constexpr int kC = kPer / 16;                                // 16-column TMEM loads per thread
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* s_q = smem;                     // 2 stages over heads
  uint8_t* s_k = smem + 2 * kQb;           // 2 stages over heads
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kQb + 2 * kKb);
  uint64_t* full = bars;         // [2] Q_h and K_h landed
  uint64_t* empty = bars + 2;    // [2]
  uint64_t* s_full = bars + 4;   // [2]
  uint64_t* s_free = bars + 6;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb = blockIdx.x % p.k_blocks;
  const int qt = (blockIdx.x / p.k_blocks) % p.q_tiles;
  const int b = blockIdx.x / (p.k_blocks * p.q_tiles);
  const int D = p.D;
  const bool two_box = !kSplit && !kCompact && D > 64;
  const uint32_t stage_tx = kCompact ? (kQb + kKb) : 2 * ((kSplit || two_box) ? kTileBytes : kBoxBytes);

  if (warp == 0 && lane == 0) ptx::prefetch_tmap(&tmap_qkv);
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&full[i], 1), ptx::mbar_init(&empty[i], 1);
      ptx::mbar_init(&s_full[i], 1), ptx::mbar_init(&s_free[i], 8);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc<kCompact ? 128 : 256>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    for (int h = 0; h < p.H; ++h) {
      const int st = h & 1;
      ptx::mbar_wait(&empty[st], ((h >> 1) & 1) ^ 1);
      if (ptx::elect_one()) {
        uint8_t* q = s_q + st * kQb;
        uint8_t* k = s_k + st * kKb;
        ptx::mbar_arrive_expect_tx(&full[st], stage_tx);
        ptx::tma_load_3d(q, &tmap_qkv, &full[st], h * D, qt * BM, b);
        if (kCompact) ptx::tma_load_3d(k, &tmap_qkv_lo, &full[st], p.d + h * D, kb * kKeys, b);   // 64-row box map
        else ptx::tma_load_3d(k, &tmap_qkv, &full[st], p.d + h * D, kb * BK, b);
        if (kSplit) {
          ptx::tma_load_3d(q + kBoxBytes, &tmap_qkv_lo, &full[st], h * D, qt * BM, b);
          ptx::tma_load_3d(k + kBoxBytes, &tmap_qkv_lo, &full[st], p.d + h * D, kb * BK, b);
        } else if (two_box) {
          ptx::tma_load_3d(q + kBoxBytes, &tmap_qkv, &full[st], h * D + 64, qt * BM, b);
          ptx::tma_load_3d(k + kBoxBytes, &tmap_qkv, &full[st], p.d + h * D + 64, kb * BK, b);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    for (int h = 0; h < p.H; ++h) {
      const int st = h & 1;
      ptx::mbar_wait(&full[st], (h >> 1) & 1);
      if (h >= 2) ptx::mbar_wait(&s_free[st], ((h - 2) >> 1) & 1);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        if (kSplit) issue_qk_long_split(tmem_base + st * BK, ptx::smem_u32(s_q + st * kTileBytes), ptx::smem_u32(s_k + st * kTileBytes));
        else issue_qk_long(tmem_base + st * kKeys, ptx::smem_u32(s_q + st * kQb), ptx::smem_u32(s_k + st * kKb), D, kKeys);
        ptx::umma_commit(&empty[st]);
        ptx::umma_commit(&s_full[st]);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    const int half = (warp - 4) >> 2, quarter = warp & 3;
    const int r = quarter * 32 + lane;
    const int qrow = qt * BM + r;
    const bool row_ok = qrow < p.N;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const int key0 = kb * kKeys + half * kPer;
    const bool want_avg = p.avg_map != nullptr;
    const bool tail = key0 + kPer > p.N;   // this thread's keys include padding (warp-uniform)
    float acc[kC][16];
#pragma unroll
    for (int c = 0; c < kC; ++c)
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[c][j] = 0.f;
    // row statistics of the next head are fetched one head ahead (a dependent ~1 us global load per head otherwise
    // sits in front of every exponential pass)
    float2 ms_next = make_float2(0.f, 0.f);
    if (row_ok) ms_next = p.stats[(static_cast<size_t>(b) * p.H) * p.N + qrow];
    for (int h = 0; h < p.H; ++h) {
      const int st = h & 1;
      const float2 ms = ms_next;
      if (row_ok && h + 1 < p.H) ms_next = p.stats[(static_cast<size_t>(b) * p.H + h + 1) * p.N + qrow];
      ptx::mbar_wait(&s_full[st], (h >> 1) & 1);
      ptx::tc_fence_after();
      uint32_t s[kC][16];
#pragma unroll
      for (int c = 0; c < kC; ++c) ptx::tmem_ld_x16(lane_base + st * kKeys + half * kPer + c * 16, s[c]);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&s_free[st]);
      float* hp = nullptr;
      if (kHeads && p.head_map != nullptr && row_ok)
        hp = p.head_map + ((static_cast<size_t>(b) * p.H + h) * p.N + qrow) * p.ldmap + key0;
      float* cp = (p.cls_map != nullptr && qrow == 0) ? p.cls_map + (static_cast<size_t>(b) * p.H + h) * p.N : nullptr;
      // Common case -- every key of the block is real, no per-head output from this thread: three instructions per
      // probability (scale-and-shift, exp2, accumulate).  The general path below costs ~15 (bounds predicates, the
      // separate product for the per-head outputs), which made this kernel issue-bound.
      if (!tail && hp == nullptr && cp == nullptr) {
#pragma unroll
        for (int c = 0; c < kC; ++c)
#pragma unroll
          for (int j = 0; j < 16; ++j)
            acc[c][j] = fmaf(ptx::ex2_approx(fmaf(__uint_as_float(s[c][j]), p.scale_log2, -ms.x)), ms.y, acc[c][j]);
        continue;
      }
#pragma unroll
      for (int c = 0; c < kC; ++c) {
        float pr[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const bool valid = key0 + c * 16 + j < p.N;
          pr[j] = valid ? ptx::ex2_approx(fmaf(__uint_as_float(s[c][j]), p.scale_log2, -ms.x)) * ms.y : 0.f;
          acc[c][j] += pr[j];
        }
        if (kHeads && hp != nullptr) {
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            if (key0 + c * 16 + j < p.ldmap)
              *reinterpret_cast<float4*>(hp + c * 16 + j) = make_float4(pr[j], pr[j + 1], pr[j + 2], pr[j + 3]);
        }
        if (cp != nullptr) {
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (key0 + c * 16 + j < p.N) cp[key0 + c * 16 + j] = pr[j];
        }
      }
    }
    if (want_avg && row_ok) {
      const float inv_h = 1.0f / static_cast<float>(p.H);
      float* ap = p.avg_map + (static_cast<size_t>(b) * p.N + qrow) * p.ldmap + key0;
#pragma unroll
      for (int c = 0; c < kC; ++c)
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          if (key0 + c * 16 + j < p.ldmap)
            *reinterpret_cast<float4*>(ap + c * 16 + j) =
                make_float4(acc[c][j] * inv_h, acc[c][j + 1] * inv_h, acc[c][j + 2] * inv_h, acc[c][j + 3] * inv_h);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<kCompact ? 128 : 256>(tmem_base);
  }
}

}  // namespace vitb200
