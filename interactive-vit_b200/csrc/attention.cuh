// Fused multi-head self-attention with attention-map emission (tcgen05 + TMEM), for sequences whose padded
// key count KP = round_up(N, 16) fits one UMMA N and this kernel's TMEM plan (KP <= 208): N = 197
// (224 px / patch 16) -> KP = 208.
//
// Arithmetic follows torch.nn.functional.multi_head_attention_forward's weights branch
// (torch/nn/functional.py:6630-6659: q * 1/sqrt(D), bmm(q, k^T), softmax, bmm(P, v), optional head mean)
// which is what the oracle obtains from torchvision's EncoderBlock (vision_transformer.py:110-119) with
// need_weights=True.  The 1/sqrt(D) scale is applied to the fp32 scores instead of to q (identical in
// exact arithmetic; exact in floating point too when D = 64 since the factor is a power of two).
//
// One CTA per (image b, 128-row query tile qt); the CTA loops over all H heads so that the head-averaged
// probabilities can be accumulated on chip (TMEM) and written to HBM exactly once.  384 threads:
//   warp 0 lane 0 : TMA producer  Q_h [128 x D], K_h [KP x D], V_h [KP x D] tiles of the packed qkv
//                   activation, 2-stage ring over heads
//   warp 1 lane 0 : UMMA issuer   S = Q K^T (M=128, N=KP, K=D)   -> TMEM cols [64, 64+KP)
//                                 O = P V   (M=128, N=D,  K=KP)  -> TMEM cols [0, D), V is the MN-major B.
//                   Issue order QK(0), QK(1), PV(0), QK(2), PV(1), ...: the S columns are released as soon
//                   as the softmax warps have copied them to registers, so QK^T of head h+1 and P V of
//                   head h run on the tensor pipe underneath the softmax of head h.
//   warp 2        : TMEM allocator (all 512 columns)
//   warps 4..11   : softmax.  Two threads per query row: warps 4..7 own the first half of the key columns,
//                   warps 8..11 the second half (a warp may only touch the TMEM lane quarter warp%4).  The
//                   half row (<= 112 scores) is read from TMEM ONCE and stays in registers for: row max
//                   (exchanged with the partner thread through smem) -> e = exp2(s*c - m*c) -> bf16 P tile
//                   in shared memory (128-B swizzled K-major A operand of P V) -> row sum (exchanged) ->
//                   Pbar += e / (sum * H) in TMEM cols [288, 288+KP) -> optional per-head rows to HBM.
//                   Then O * 1/sum -> bf16 context rows (each half owns 32 of the 64 columns).
#pragma once
#include <cuda.h>
#include "ptx.cuh"

namespace vitb200 {

struct AttnParams {
  int B, N, H;          // images, tokens per image, heads
  int d;                // model width = H * D
  int KP;               // keys padded to a multiple of 16 (<= 208)
  float scale_log2;     // (1/sqrt(D)) * log2(e)
  __nv_bfloat16* ctx;   // [B*N, d] attention context (input of out_proj)
  float* avg_map;       // [B, N, ldmap] head-averaged probabilities, or nullptr
  float* head_map;      // [B, H, N, ldmap] per-head probabilities, or nullptr (opt-in, large)
  float* cls_map;       // [B, H, N] per-head probabilities of query token 0, or nullptr
  int ldmap;            // row stride of avg_map/head_map in floats (>= KP, multiple of 4)
  int q_tiles;          // 128-row query tiles per image: ceil(N / 128) (1 or 2)
};

// Optional phase tracing (built only with -DVITB200_ATTN_TRACE into a separate library, tools/attn_trace.py):
// CTA 0's softmax warp 4 / lane 0 and the MMA warp write clock64() stamps per head to a global buffer.
#ifdef VITB200_ATTN_TRACE
__device__ long long g_attn_trace[64 * 32];
#define ATTN_TS(slot) do { if (blockIdx.x == 0 && lane == 0 && (warp == 4 || warp == 1)) g_attn_trace[(h) * 32 + (slot)] = clock64(); } while (0)
#else
#define ATTN_TS(slot) do { } while (0)
#endif

namespace attn_cfg {
constexpr int kThreads = 384;
constexpr int kSoftmaxThreads = 256;
constexpr int BM = 128;
constexpr int D = 64;
constexpr int KP_MAX = 208;  // TMEM plan: O [0,64) | S [64,64+KP) | Pbar [288,288+KP)  -> KP <= 208
constexpr int kMaxChunks = 7;                      // 16-column chunks per softmax thread: ceil(13 / 2)
constexpr int kTmemO = 0;
constexpr int kTmemS = 64;
constexpr int kTmemAvg = 288;
constexpr int kQBytes = BM * D * 2;                // 16 KB
constexpr int kKVBytes = KP_MAX * D * 2;           // 26 KB (two TMA boxes of KP/2 rows)
constexpr int kStageBytes = kQBytes + 2 * kKVBytes;
constexpr int kPBlockBytes = BM * 128;             // one 64-key K-block of P: 128 rows x 128 B
constexpr int kPBlocks = (KP_MAX + 63) / 64;       // 4
constexpr int kPBytes = kPBlocks * kPBlockBytes;   // 64 KB
constexpr int kRedBytes = 2 * 2 * BM * 4;          // row max / row sum exchange: [2 kinds][2 halves][128 rows]
constexpr int kClsStageBytes = 256 * 4;            // normalised probabilities of query row 0
constexpr int kSmemBytes = 2 * kStageBytes + kPBytes + kRedBytes + kClsStageBytes + 256;
}  // namespace attn_cfg

// kHeads: also write the full per-head probabilities (opt-in; a separate instantiation keeps that code out of
// the instruction stream of the common variant, which has to stay inside the 32 KB instruction cache).
template <bool kHeads>
__global__ void __launch_bounds__(attn_cfg::kThreads, 1)
attention_kernel(const __grid_constant__ CUtensorMap tmap_q,   // box 64 x 128 over qkv [B*N, 3d]
                 const __grid_constant__ CUtensorMap tmap_kv,  // box 64 x (KP/2) over the same tensor
                 AttnParams p) {
  using namespace attn_cfg;
  // Dynamic smem starts 1024-B aligned (it follows the 1 KB the driver reserves); keeping the array typed lets
  // the compiler emit LDS/STS instead of generic loads for the exchange buffers.
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* smem_p = smem + 2 * kStageBytes;
  float* red = reinterpret_cast<float*>(smem_p + kPBytes);  // [kind][half][row]
  float* cls_stage = red + 2 * 2 * BM;                       // [256] normalised probabilities of query row 0
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_p + kPBytes + kRedBytes + kClsStageBytes);
  uint64_t* full_bar = bars;        // [2] Q/K/V of a head landed
  uint64_t* empty_bar = bars + 2;   // [2] Q/K/V stage consumed by the MMAs
  uint64_t* s_full = bars + 4;      // S = QK^T complete
  uint64_t* s_free = bars + 5;      // S copied to registers by all softmax threads (256 arrivals)
  uint64_t* p_full = bars + 6;      // bf16 P tile written to smem (256 arrivals)
  uint64_t* o_full = bars + 7;      // O = PV complete
  uint64_t* o_free = bars + 8;      // O columns read by all softmax threads (256 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 9);

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
    // one arrival per softmax WARP (lane 0 after __syncwarp): 256 lanes arriving on one smem word serialise
    ptx::mbar_init(s_free, kSoftmaxThreads / 32);
    ptx::mbar_init(p_full, kSoftmaxThreads / 32);
    ptx::mbar_init(o_full, 1);
    ptx::mbar_init(o_free, kSoftmaxThreads / 32);
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int row0 = b * p.N;  // first token row of this image in the [B*N, 3d] activation

  // Producer and MMA loops run on whole warps with one elected lane issuing, so that addresses and descriptors
  // stay in uniform registers (a single-lane loop pays an R2UR chain in front of every UTMALDG / UTCHMMA).
  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    for (int h = 0; h < p.H; ++h) {
      const int st = h & 1;
      const uint32_t ph = (h >> 1) & 1;
      ptx::mbar_wait(&empty_bar[st], ph ^ 1);
      uint8_t* sq = smem + st * kStageBytes;
      uint8_t* sk = sq + kQBytes;
      uint8_t* sv = sk + kKVBytes;
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(&full_bar[st], stage_tx);
        ptx::tma_load_2d(sq, &tmap_q, &full_bar[st], h * D, row0 + qt * BM);
        ptx::tma_load_2d(sk, &tmap_kv, &full_bar[st], p.d + h * D, row0);
        ptx::tma_load_2d(sk + half_rows * 128, &tmap_kv, &full_bar[st], p.d + h * D, row0 + half_rows);
        ptx::tma_load_2d(sv, &tmap_kv, &full_bar[st], 2 * p.d + h * D, row0);
        ptx::tma_load_2d(sv + half_rows * 128, &tmap_kv, &full_bar[st], 2 * p.d + h * D, row0 + half_rows);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ UMMA issuer
    const uint32_t idesc_qk = ptx::make_idesc_bf16(BM, static_cast<uint32_t>(KP), 0, 0);
    const uint32_t idesc_pv = ptx::make_idesc_bf16(BM, D, 0, 1);  // B (= V) is MN-major
    const uint32_t sp = ptx::smem_u32(smem_p);
    const int ksteps = KP >> 4;
    auto issue_qk = [&](int h) {
      const int st = h & 1;
      ptx::mbar_wait(&full_bar[st], (h >> 1) & 1);
      if (h > 0) ptx::mbar_wait(s_free, (h - 1) & 1);
      ptx::tc_fence_after();
      const uint32_t sq = ptx::smem_u32(smem + st * kStageBytes);
      const uint64_t dq = ptx::make_smem_desc_sw128(sq, 16, 1024);
      const uint64_t dk = ptx::make_smem_desc_sw128(sq + kQBytes, 16, 1024);
      if (ptx::elect_one()) {
#pragma unroll
        for (int k = 0; k < D / 16; ++k)
          ptx::umma_bf16_ss(tmem_base + kTmemS, dq + 2 * k, dk + 2 * k, idesc_qk, k != 0 ? 1u : 0u);
        ptx::umma_commit(s_full);
      }
      __syncwarp();
    };
    issue_qk(0);
    for (int h = 0; h < p.H; ++h) {
      if (h + 1 < p.H) issue_qk(h + 1);
      const int st = h & 1;
      ATTN_TS(16);
      ptx::mbar_wait(p_full, h & 1);
      ATTN_TS(17);
      if (h > 0) ptx::mbar_wait(o_free, (h - 1) & 1);
      ATTN_TS(18);
      ptx::tc_fence_after();
      const uint32_t sv = ptx::smem_u32(smem + st * kStageBytes) + kQBytes + kKVBytes;
      // A: P K-block (ks / 4), +32 B per 16 keys inside the swizzle span.  B: V rows [16 ks, 16 ks + 16),
      // MN-major: 8-key groups 1024 B apart (SBO); the single 64-wide MN group makes LBO irrelevant.
      // Descriptor start addresses are in 16-B units: +2 per 16 keys of P inside a K-block, +1024 (= 16 KB) to the
      // next K-block, +128 (= 2 KB) per 16 keys of V.
      const uint64_t dp0 = ptx::make_smem_desc_sw128(sp, 16, 1024);
      const uint64_t dv0 = ptx::make_smem_desc_sw128(sv, 1024, 1024);
      if (ptx::elect_one()) {
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint64_t dp = dp0 + static_cast<uint64_t>((ks >> 2) * (kPBlockBytes >> 4) + 2 * (ks & 3));
          const uint64_t dv = dv0 + static_cast<uint64_t>(ks * (2048 >> 4));
          ptx::umma_bf16_ss(tmem_base + kTmemO, dp, dv, idesc_pv, ks != 0 ? 1u : 0u);
        }
        ptx::umma_commit(&empty_bar[st]);
        ptx::umma_commit(o_full);
      }
      __syncwarp();
      ATTN_TS(19);
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ softmax / epilogue
    const int half = (warp - 4) >> 2;           // which half of the key columns
    const int quarter = warp & 3;               // TMEM lane quarter
    const int r = quarter * 32 + lane;          // row inside the tile
    const int qrow = qt * BM + r;               // token index inside the image
    const bool row_ok = qrow < p.N;
    const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    const int nchunks = KP >> 4;
    const int split = (nchunks + 1) >> 1;       // 16-key chunks [0, split) -> half 0, [split, nchunks) -> half 1
    const int nmy = half ? nchunks - split : split;
    // local chunk c <-> global chunk gc: half 0 walks up from 0, half 1 walks DOWN from the last chunk, so the
    // only chunk that can hold padded keys (>= N) is always local chunk 0 of half 1 (a static register index)
    const int gc_base = half ? nchunks - 1 : 0;
    const int gc_step = half ? -1 : 1;
    const int tail_valid = p.N - (nchunks - 1) * 16;  // valid keys in the last chunk (1..16)
    const float inv_h = 1.0f / static_cast<float>(p.H);
    const bool want_avg = p.avg_map != nullptr;
    const bool want_cls = p.cls_map != nullptr && qt == 0;
    const bool want_maps = want_avg || want_cls || kHeads;
    const uint32_t p_row = ptx::smem_u32(smem_p) + r * 128;
    const int sw = r & 7;
    float* red_max = red;                        // [half][row]
    float* red_sum = red + 2 * BM;
    const uint32_t t_s = lane_base + kTmemS + gc_base * 16;
    const uint32_t t_avg = lane_base + kTmemAvg + gc_base * 16;
    const int t_step = gc_step * 16;

    for (int h = 0; h < p.H; ++h) {
      // ---- S half-row -> registers (single TMEM read), then release the S columns
      uint32_t s[kMaxChunks][16];
      ATTN_TS(0);
      ptx::mbar_wait(s_full, h & 1);
      ATTN_TS(1);
      ptx::tc_fence_after();
#pragma unroll
      for (int c = 0; c < kMaxChunks; ++c)
        if (c < nmy) ptx::tmem_ld_x16(t_s + c * t_step, s[c]);
      ptx::tmem_ld_wait();
      ATTN_TS(2);
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(s_free);
      if (half && tail_valid < 16) {
#pragma unroll
        for (int j = 0; j < 16; ++j)
          if (j >= tail_valid) s[0][j] = 0xff800000u;  // -inf: exp2 gives 0, max ignores it
      }

      // ---- row max (own half, then partner's through smem)
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < kMaxChunks; ++c) {
        if (c < nmy) {
#pragma unroll
          for (int j = 0; j < 16; ++j) mx = fmaxf(mx, __uint_as_float(s[c][j]));
        }
      }
      red_max[half * BM + r] = mx;
      ATTN_TS(3);
      asm volatile("bar.sync 1, 256;" ::: "memory");
      ATTN_TS(4);
      mx = fmaxf(mx, red_max[(half ^ 1) * BM + r]);
      const float mxs = mx * p.scale_log2;

      // ---- e = exp2(s*c - max*c) in place; bf16 P to smem (swizzled K-major)
      float ps0 = 0.f, ps1 = 0.f, ps2 = 0.f, ps3 = 0.f;
#pragma unroll
      for (int c = 0; c < kMaxChunks; ++c) {
        if (c < nmy) {
          const int gc = gc_base + c * gc_step;
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float v0 = ptx::ex2_approx(fmaf(__uint_as_float(s[c][j]), p.scale_log2, -mxs));
            const float v1 = ptx::ex2_approx(fmaf(__uint_as_float(s[c][j + 1]), p.scale_log2, -mxs));
            const float v2 = ptx::ex2_approx(fmaf(__uint_as_float(s[c][j + 2]), p.scale_log2, -mxs));
            const float v3 = ptx::ex2_approx(fmaf(__uint_as_float(s[c][j + 3]), p.scale_log2, -mxs));
            ps0 += v0, ps1 += v1, ps2 += v2, ps3 += v3;
            s[c][j] = __float_as_uint(v0), s[c][j + 1] = __float_as_uint(v1);
            s[c][j + 2] = __float_as_uint(v2), s[c][j + 3] = __float_as_uint(v3);
          }
          uint32_t pk[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            __nv_bfloat162 t = __floats2bfloat162_rn(__uint_as_float(s[c][2 * j]), __uint_as_float(s[c][2 * j + 1]));
            pk[j] = *reinterpret_cast<uint32_t*>(&t);
          }
          // keys [16 gc, 16 gc + 16): K-block gc/4, 16-byte chunks 2*(gc%4) and 2*(gc%4)+1 of this row
          const uint32_t blk = p_row + (gc >> 2) * kPBlockBytes;
          const int ch0 = 2 * (gc & 3);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(blk + ((ch0 ^ sw) << 4)), "r"(pk[0]),
                       "r"(pk[1]), "r"(pk[2]), "r"(pk[3])
                       : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(blk + (((ch0 + 1) ^ sw) << 4)), "r"(pk[4]),
                       "r"(pk[5]), "r"(pk[6]), "r"(pk[7])
                       : "memory");
        }
      }
      ATTN_TS(5);
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(p_full);
      ATTN_TS(6);

      // ---- row sum exchange
      float sum = (ps0 + ps1) + (ps2 + ps3);
      red_sum[half * BM + r] = sum;
      asm volatile("bar.sync 2, 256;" ::: "memory");
      ATTN_TS(7);
      sum += red_sum[(half ^ 1) * BM + r];
      const float inv = ptx::rcp_approx(sum);

      // ---- normalised probabilities -> head average (TMEM) / per-head rows (HBM); overlaps the P V MMAs
      if (want_maps) {
        const float wavg = inv * inv_h;
        if (want_avg) {
#pragma unroll
          for (int c = 0; c < kMaxChunks; ++c) {
            if (c < nmy) {
              uint32_t a[16];
              if (h > 0) {
                ptx::tmem_ld_x16(t_avg + c * t_step, a);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  a[j] = __float_as_uint(fmaf(__uint_as_float(s[c][j]), wavg, __uint_as_float(a[j])));
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) a[j] = __float_as_uint(__uint_as_float(s[c][j]) * wavg);
              }
              ptx::tmem_st_x16(t_avg + c * t_step, a);
            }
          }
          ptx::tmem_st_wait();
        }
        if (want_cls && quarter == 0) {
          // query row 0 lives in lane 0 of warps 4 (first half of the keys) and 8 (second half): it stages its
          // normalised values, then the whole warp writes that key range out (no cross-warp dependency)
          if (lane == 0) {
#pragma unroll
            for (int c = 0; c < kMaxChunks; ++c) {
              if (c < nmy) {
                float* dst = cls_stage + (gc_base + c * gc_step) * 16;
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                  *reinterpret_cast<float4*>(dst + j) =
                      make_float4(__uint_as_float(s[c][j]) * inv, __uint_as_float(s[c][j + 1]) * inv,
                                  __uint_as_float(s[c][j + 2]) * inv, __uint_as_float(s[c][j + 3]) * inv);
              }
            }
          }
          __syncwarp();
          const int lo = half ? split * 16 : 0;
          const int hi = half ? p.N : min(split * 16, p.N);
          float* cp = p.cls_map + (static_cast<size_t>(b) * p.H + h) * p.N;
          for (int j = lo + lane; j < hi; j += 32) cp[j] = cls_stage[j];
          __syncwarp();
        }
        if (kHeads) {
          if (p.head_map != nullptr && row_ok) {
            float* hp = p.head_map + ((static_cast<size_t>(b) * p.H + h) * p.N + qrow) * p.ldmap;
#pragma unroll
            for (int c = 0; c < kMaxChunks; ++c) {
              if (c < nmy) {
                float* dst = hp + (gc_base + c * gc_step) * 16;
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                  *reinterpret_cast<float4*>(dst + j) =
                      make_float4(__uint_as_float(s[c][j]) * inv, __uint_as_float(s[c][j + 1]) * inv,
                                  __uint_as_float(s[c][j + 2]) * inv, __uint_as_float(s[c][j + 3]) * inv);
              }
            }
          }
        }
      }

      // ---- O epilogue: context rows = (P_unnormalised V) / sum; this half owns 32 of the 64 columns
      ATTN_TS(8);
      ptx::mbar_wait(o_full, h & 1);
      ATTN_TS(9);
      ptx::tc_fence_after();
      {
        uint32_t o[32];
        ptx::tmem_ld_x32(lane_base + kTmemO + half * 32, o);
        ptx::tmem_ld_wait();
        ATTN_TS(10);
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(o_free);
        if (row_ok) {
          __nv_bfloat16* op = p.ctx + (static_cast<size_t>(row0) + qrow) * p.d + h * D + half * 32;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 pk;
            __nv_bfloat162 t0 = __floats2bfloat162_rn(__uint_as_float(o[j]) * inv, __uint_as_float(o[j + 1]) * inv);
            __nv_bfloat162 t1 = __floats2bfloat162_rn(__uint_as_float(o[j + 2]) * inv, __uint_as_float(o[j + 3]) * inv);
            __nv_bfloat162 t2 = __floats2bfloat162_rn(__uint_as_float(o[j + 4]) * inv, __uint_as_float(o[j + 5]) * inv);
            __nv_bfloat162 t3 = __floats2bfloat162_rn(__uint_as_float(o[j + 6]) * inv, __uint_as_float(o[j + 7]) * inv);
            pk.x = *reinterpret_cast<uint32_t*>(&t0);
            pk.y = *reinterpret_cast<uint32_t*>(&t1);
            pk.z = *reinterpret_cast<uint32_t*>(&t2);
            pk.w = *reinterpret_cast<uint32_t*>(&t3);
            *reinterpret_cast<uint4*>(op + j) = pk;
          }
        }
      }
    }

    // head-averaged map rows -> HBM, once per (image, query row)
    if (want_avg) {
#pragma unroll 1
      for (int c = 0; c < nmy; ++c) {
        uint32_t a[16];
        ptx::tmem_ld_x16(t_avg + c * t_step, a);
        ptx::tmem_ld_wait();
        if (row_ok) {
          float* ap = p.avg_map + (static_cast<size_t>(b) * p.N + qrow) * p.ldmap + (gc_base + c * gc_step) * 16;
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
