// Fused multi-head self-attention with attention-map emission (tcgen05 + TMEM), for sequences whose padded
// key count KP = round_up(N, 16) fits one UMMA N (<= 256): N = 197 (224 px / patch 16) -> KP = 208.
//
// Arithmetic follows torch.nn.functional.multi_head_attention_forward's weights branch
// (torch/nn/functional.py:6630-6659: q * 1/sqrt(D), bmm(q, k^T), softmax, bmm(P, v), optional head mean)
// which is what the oracle obtains from torchvision's EncoderBlock (vision_transformer.py:110-119) with
// need_weights=True.  The 1/sqrt(D) scale is applied to the fp32 scores instead of to q (identical in
// exact arithmetic; exact in floating point too when D = 64 since the factor is a power of two).
//
// One CTA per (image b, 128-row query tile qt); the CTA loops over all H heads so that the head-averaged
// probabilities can be accumulated on chip (TMEM) and written to HBM exactly once:
//   warp 0 lane 0 : TMA producer  Q_h [128 x D], K_h [KP x D], V_h [KP x D] tiles of the packed qkv
//                   activation, 2-stage ring over heads (head h+1 streams in while head h computes)
//   warp 1 lane 0 : UMMA issuer   S = Q K^T (M=128, N=KP, K=D)   -> TMEM cols [64, 64+KP)
//                                 O = P V   (M=128, N=D,  K=KP)  -> TMEM cols [0, D), V is the MN-major B
//   warp 2        : TMEM allocator (all 512 columns)
//   warps 4..7    : softmax, one thread per query row (32x32b TMEM loads: no cross-lane reductions):
//                   pass 1 row max; pass 2 e = exp2(s*c - m*c) -> bf16 P tile in shared memory (128-B
//                   swizzled K-major A operand for P V) and fp32 e back into the S columns; pass 3
//                   Pbar += e / (sum * H) in TMEM cols [288, 288+KP), optional per-head rows to HBM;
//                   then O * 1/sum -> bf16 context rows.
#pragma once
#include <cuda.h>
#include "ptx.cuh"

namespace vitb200 {

struct AttnParams {
  int B, N, H;          // images, tokens per image, heads
  int d;                // model width = H * D
  int KP;               // keys padded to a multiple of 16 (<= 208 for this kernel's TMEM plan, see below)
  float scale_log2;     // (1/sqrt(D)) * log2(e)
  __nv_bfloat16* ctx;   // [B*N, d] attention context (input of out_proj)
  float* avg_map;       // [B, N, ldmap] head-averaged probabilities, or nullptr
  float* head_map;      // [B, H, N, ldmap] per-head probabilities, or nullptr (opt-in, large)
  float* cls_map;       // [B, H, N] per-head probabilities of query token 0, or nullptr
  int ldmap;            // row stride of avg_map/head_map in floats (>= KP, multiple of 4)
  int q_tiles;          // 128-row query tiles per image: ceil(N / 128) (1 or 2)
};

namespace attn_cfg {
constexpr int kThreads = 256;
constexpr int BM = 128;
constexpr int D = 64;
constexpr int KP_MAX = 208;  // TMEM plan: O [0,64) | S [64,64+KP) | Pbar [288,288+KP)  -> KP <= 208
constexpr int kTmemO = 0;
constexpr int kTmemS = 64;
constexpr int kTmemAvg = 288;
constexpr int kQBytes = BM * D * 2;                // 16 KB
constexpr int kKVBytes = KP_MAX * D * 2;           // 26 KB (two TMA boxes of KP/2 rows)
constexpr int kStageBytes = kQBytes + 2 * kKVBytes;
constexpr int kPBlockBytes = BM * 128;             // one 64-key K-block of P: 128 rows x 128 B
constexpr int kPBlocks = (KP_MAX + 63) / 64;       // 4
constexpr int kPBytes = kPBlocks * kPBlockBytes;   // 64 KB
constexpr int kClsStageBytes = 256 * 4;
constexpr int kSmemBytes = 2 * kStageBytes + kPBytes + kClsStageBytes + 1024 + 256;
}  // namespace attn_cfg

__global__ void __launch_bounds__(attn_cfg::kThreads, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmap_q,   // box 64 x 128 over qkv [B*N, 3d]
                 const __grid_constant__ CUtensorMap tmap_kv,  // box 64 x (KP/2) over the same tensor
                 AttnParams p) {
  using namespace attn_cfg;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_p = smem + 2 * kStageBytes;
  float* cls_stage = reinterpret_cast<float*>(smem_p + kPBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_p + kPBytes + kClsStageBytes);
  uint64_t* full_bar = bars;        // [2] Q/K/V of a head landed
  uint64_t* empty_bar = bars + 2;   // [2] Q/K/V stage consumed by the MMAs
  uint64_t* s_full = bars + 4;      // S = QK^T complete
  uint64_t* p_full = bars + 5;      // bf16 P tile written to smem (128 arrivals)
  uint64_t* o_full = bars + 6;      // O = PV complete
  uint64_t* s_free = bars + 7;      // softmax done with S / O / Pbar columns of this head (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.x / p.q_tiles;
  const int qt = blockIdx.x - b * p.q_tiles;
  const int KP = p.KP;
  const int half_rows = KP >> 1;
  const uint32_t stage_tx = kQBytes + 2 * static_cast<uint32_t>(KP) * D * 2;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_q);
    ptx::prefetch_tmap(&tmap_kv);
  }
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(&full_bar[0], 1);
    ptx::mbar_init(&full_bar[1], 1);
    ptx::mbar_init(&empty_bar[0], 1);
    ptx::mbar_init(&empty_bar[1], 1);
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(p_full, 128);
    ptx::mbar_init(o_full, 1);
    ptx::mbar_init(s_free, 128);
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int row0 = b * p.N;  // first token row of this image in the [B*N, 3d] activation

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------ TMA producer
    for (int h = 0; h < p.H; ++h) {
      const int st = h & 1;
      const uint32_t ph = (h >> 1) & 1;
      ptx::mbar_wait(&empty_bar[st], ph ^ 1);
      uint8_t* sq = smem + st * kStageBytes;
      uint8_t* sk = sq + kQBytes;
      uint8_t* sv = sk + kKVBytes;
      ptx::mbar_arrive_expect_tx(&full_bar[st], stage_tx);
      ptx::tma_load_2d(sq, &tmap_q, &full_bar[st], h * D, row0 + qt * BM);
      ptx::tma_load_2d(sk, &tmap_kv, &full_bar[st], p.d + h * D, row0);
      ptx::tma_load_2d(sk + half_rows * 128, &tmap_kv, &full_bar[st], p.d + h * D, row0 + half_rows);
      ptx::tma_load_2d(sv, &tmap_kv, &full_bar[st], 2 * p.d + h * D, row0);
      ptx::tma_load_2d(sv + half_rows * 128, &tmap_kv, &full_bar[st], 2 * p.d + h * D, row0 + half_rows);
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------------------ UMMA issuer
    const uint32_t idesc_qk = ptx::make_idesc_bf16(BM, static_cast<uint32_t>(KP), 0, 0);
    const uint32_t idesc_pv = ptx::make_idesc_bf16(BM, D, 0, 1);  // B (= V) is MN-major
    const uint32_t sp = ptx::smem_u32(smem_p);
    const int ksteps = KP >> 4;
    for (int h = 0; h < p.H; ++h) {
      const int st = h & 1;
      const uint32_t ph = (h >> 1) & 1;
      ptx::mbar_wait(&full_bar[st], ph);
      if (h > 0) ptx::mbar_wait(s_free, (h - 1) & 1);
      ptx::tc_fence_after();
      const uint32_t sq = ptx::smem_u32(smem + st * kStageBytes);
      const uint32_t sk = sq + kQBytes;
      const uint32_t sv = sk + kKVBytes;
      const uint64_t dq = ptx::make_smem_desc_sw128(sq, 16, 1024);
      const uint64_t dk = ptx::make_smem_desc_sw128(sk, 16, 1024);
#pragma unroll
      for (int k = 0; k < D / 16; ++k)
        ptx::umma_bf16_ss(tmem_base + kTmemS, dq + 2 * k, dk + 2 * k, idesc_qk, k != 0 ? 1u : 0u);
      ptx::umma_commit(s_full);

      ptx::mbar_wait(p_full, h & 1);
      ptx::tc_fence_after();
      for (int ks = 0; ks < ksteps; ++ks) {
        // A: P K-block (ks / 4), +32 B per 16 keys inside the swizzle span.  B: V rows [16 ks, 16 ks + 16),
        // MN-major: 8-key groups 1024 B apart (SBO); the single 64-wide MN group makes LBO irrelevant.
        const uint64_t dp = ptx::make_smem_desc_sw128(sp + (ks >> 2) * kPBlockBytes, 16, 1024) + 2 * (ks & 3);
        const uint64_t dv = ptx::make_smem_desc_sw128(sv + ks * 2048, 1024, 1024);
        ptx::umma_bf16_ss(tmem_base + kTmemO, dp, dv, idesc_pv, ks != 0 ? 1u : 0u);
      }
      ptx::umma_commit(&empty_bar[st]);
      ptx::umma_commit(o_full);
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ softmax / epilogue (thread = query row)
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;          // row inside the tile
    const int qrow = qt * BM + r;               // token index inside the image
    const bool row_ok = qrow < p.N;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const int nchunks = KP >> 4;
    const float inv_h = 1.0f / static_cast<float>(p.H);
    const bool want_avg = p.avg_map != nullptr;
    const bool want_cls = p.cls_map != nullptr && qt == 0;
    const uint32_t p_row = ptx::smem_u32(smem_p) + r * 128;
    const int sw = r & 7;

    for (int h = 0; h < p.H; ++h) {
      ptx::mbar_wait(s_full, h & 1);
      ptx::tc_fence_after();
      // pass 1: row maximum over the valid keys
      float mx = -INFINITY;
      for (int c = 0; c < nchunks; ++c) {
        uint32_t s[16];
        ptx::tmem_ld_x16(lane_base + kTmemS + c * 16, s);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (c * 16 + j < p.N) mx = fmaxf(mx, __uint_as_float(s[j]));
      }
      const float mxs = mx * p.scale_log2;
      // pass 2: e = exp2(s*c - max*c); bf16 P to smem (swizzled K-major), fp32 e back to TMEM
      float sum = 0.f;
      for (int c = 0; c < nchunks; ++c) {
        uint32_t s[16];
        ptx::tmem_ld_x16(lane_base + kTmemS + c * 16, s);
        ptx::tmem_ld_wait();
        float e[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float v = exp2f(fmaf(__uint_as_float(s[j]), p.scale_log2, -mxs));
          e[j] = (c * 16 + j < p.N) ? v : 0.f;
          sum += e[j];
          s[j] = __float_as_uint(e[j]);
        }
        ptx::tmem_st_x16(lane_base + kTmemS + c * 16, s);
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          __nv_bfloat162 t = __floats2bfloat162_rn(e[2 * j], e[2 * j + 1]);
          pk[j] = *reinterpret_cast<uint32_t*>(&t);
        }
        // keys [16c, 16c+16): K-block kb = c/4, 16-byte chunks 2*(c%4) and 2*(c%4)+1 of this row
        const uint32_t blk = p_row + (c >> 2) * kPBlockBytes;
        const int ch0 = 2 * (c & 3);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(blk + ((ch0 ^ sw) << 4)), "r"(pk[0]),
                     "r"(pk[1]), "r"(pk[2]), "r"(pk[3])
                     : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(blk + (((ch0 + 1) ^ sw) << 4)), "r"(pk[4]),
                     "r"(pk[5]), "r"(pk[6]), "r"(pk[7])
                     : "memory");
      }
      ptx::tmem_st_wait();
      ptx::fence_proxy_async_smem();
      ptx::mbar_arrive(p_full);

      // pass 3 (overlaps the P V MMAs): normalised probabilities -> head average / per-head rows
      const float inv = 1.0f / sum;
      if (want_avg || want_cls || p.head_map != nullptr) {
        const float wavg = inv * inv_h;
        for (int c = 0; c < nchunks; ++c) {
          uint32_t e[16];
          ptx::tmem_ld_x16(lane_base + kTmemS + c * 16, e);
          if (want_avg) {
            uint32_t a[16];
            if (h > 0) {
              ptx::tmem_ld_x16(lane_base + kTmemAvg + c * 16, a);
              ptx::tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j)
                a[j] = __float_as_uint(fmaf(__uint_as_float(e[j]), wavg, __uint_as_float(a[j])));
            } else {
              ptx::tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 16; ++j) a[j] = __float_as_uint(__uint_as_float(e[j]) * wavg);
            }
            ptx::tmem_st_x16(lane_base + kTmemAvg + c * 16, a);
          } else {
            ptx::tmem_ld_wait();
          }
          if (want_cls && r == 0) {
#pragma unroll
            for (int j = 0; j < 16; ++j) cls_stage[c * 16 + j] = __uint_as_float(e[j]) * inv;
          }
          if (p.head_map != nullptr && row_ok) {
            float* hp = p.head_map + ((static_cast<size_t>(b) * p.H + h) * p.N + qrow) * p.ldmap + c * 16;
#pragma unroll
            for (int j = 0; j < 16; j += 4)
              *reinterpret_cast<float4*>(hp + j) =
                  make_float4(__uint_as_float(e[j]) * inv, __uint_as_float(e[j + 1]) * inv,
                              __uint_as_float(e[j + 2]) * inv, __uint_as_float(e[j + 3]) * inv);
          }
        }
        if (want_avg) ptx::tmem_st_wait();
        if (want_cls) {
          asm volatile("bar.sync 1, 128;" ::: "memory");
          float* cp = p.cls_map + (static_cast<size_t>(b) * p.H + h) * p.N;
          for (int j = threadIdx.x - 128; j < p.N; j += 128) cp[j] = cls_stage[j];
          asm volatile("bar.sync 1, 128;" ::: "memory");
        }
      }

      // O epilogue: context rows = (P_unnormalised V) / sum
      ptx::mbar_wait(o_full, h & 1);
      ptx::tc_fence_after();
      {
        uint32_t o[4][16];
#pragma unroll
        for (int c = 0; c < 4; ++c) ptx::tmem_ld_x16(lane_base + kTmemO + c * 16, o[c]);
        ptx::tmem_ld_wait();
        if (row_ok) {
          __nv_bfloat16* op = p.ctx + (static_cast<size_t>(row0) + qrow) * p.d + h * D;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
#pragma unroll
            for (int j = 0; j < 16; j += 8) {
              uint4 pk;
              __nv_bfloat162 t0 = __floats2bfloat162_rn(__uint_as_float(o[c][j]) * inv, __uint_as_float(o[c][j + 1]) * inv);
              __nv_bfloat162 t1 = __floats2bfloat162_rn(__uint_as_float(o[c][j + 2]) * inv, __uint_as_float(o[c][j + 3]) * inv);
              __nv_bfloat162 t2 = __floats2bfloat162_rn(__uint_as_float(o[c][j + 4]) * inv, __uint_as_float(o[c][j + 5]) * inv);
              __nv_bfloat162 t3 = __floats2bfloat162_rn(__uint_as_float(o[c][j + 6]) * inv, __uint_as_float(o[c][j + 7]) * inv);
              pk.x = *reinterpret_cast<uint32_t*>(&t0);
              pk.y = *reinterpret_cast<uint32_t*>(&t1);
              pk.z = *reinterpret_cast<uint32_t*>(&t2);
              pk.w = *reinterpret_cast<uint32_t*>(&t3);
              *reinterpret_cast<uint4*>(op + c * 16 + j) = pk;
            }
          }
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(s_free);
    }

    // head-averaged map rows -> HBM, once per (image, query row)
    if (want_avg) {
      for (int c = 0; c < nchunks; ++c) {
        uint32_t a[16];
        ptx::tmem_ld_x16(lane_base + kTmemAvg + c * 16, a);
        ptx::tmem_ld_wait();
        if (row_ok) {
          float* ap = p.avg_map + (static_cast<size_t>(b) * p.N + qrow) * p.ldmap + c * 16;
#pragma unroll
          for (int j = 0; j < 16; j += 4)
            *reinterpret_cast<float4*>(ap + j) = make_float4(__uint_as_float(a[j]), __uint_as_float(a[j + 1]),
                                                             __uint_as_float(a[j + 2]), __uint_as_float(a[j + 3]));
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace vitb200
