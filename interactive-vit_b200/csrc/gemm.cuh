// Persistent, warp-specialised tcgen05 GEMM for every linear map on the ViT forward path:
//     out[m, n] = epilogue( sum_k A[m, k] * W[n, k] + bias[n] )
// A = activations [M, K] bf16 row-major (K-major), W = nn.Linear weight [N, K] bf16 (K-major), fp32
// accumulation in TMEM.  It replaces, on the reference's hot path (Model.compute -> sub(x),
// main/context.py:79-88), the ATen CPU sgemm calls made by torchvision's
//   conv_proj (as a GEMM over im2col'ed patches)  vision_transformer.py:268-287
//   MultiheadAttention in_proj / out_proj         torch/nn/functional.py:6244 ff.
//   MLPBlock Linear -> GELU(erf) -> Linear        vision_transformer.py:40-47
//   heads.head                                    vision_transformer.py:302-304
//
// kPair = 2 (default): a CTA pair (cluster of 2, cta_group::2) computes a 256 x BN output tile with
// UMMA M = 256: each CTA stages its own 128 rows of A and HALF of the W tile, so the L2 -> SMEM traffic per
// FLOP is 2/3 of the single-CTA kernel's (the 128 x 256 single-CTA tile is L2-bandwidth bound at ~900 TF/s).
// kPair = 1: one CTA computes 128 x BN on its own (kept as the A/B baseline).
// kPairs = 2 (BN = 256 only, opt-in: VITB200_GEMM_PAIR=4): a cluster of FOUR CTAs = two pairs computing two tiles that
// share one operand; each CTA loads 64 rows of the shared 128-row block and TMA-multicasts them to its counterpart
// in the other pair.  Normally the pairs take vertically adjacent tiles (same W tile, "share W"); an odd last row of
// tiles is covered with horizontally adjacent tiles (same A tile, "share A").  Built on the hypothesis that the pair
// kernel was bound by the L2 -> SMEM feed; measured, it is NOT faster (multicast does not reduce the bytes delivered
// into each SM, and only 33 four-CTA clusters are co-resident on 148 SMs) -- kept as a correct, tested negative result.
//
// Roles (4 + 8 warps, or 4 + 16 for the GELU epilogues; 1 CTA / SM, persistent over output tiles; role = hardware warp
// rotated so that the control roles own the highest warp ids):
//   role 0 (elected lane) : TMA producer  (A tile 128x64, W tile (BN/kPair)x64, SWIZZLE_128B, mbarrier ring; in a pair
//                   both CTAs' loads complete_tx on the LEADER's full barrier)
//   role 1 (elected lane) : UMMA issuer (leader CTA only in a pair): 4 x tcgen05.mma (K = 16) per stage, commit ->
//                   empty barrier of both CTAs; accumulator-complete commit -> both CTAs' epilogues
//   role 2        : TMEM allocator (2 accumulator buffers of BN fp32 columns: MMA of tile i+1 overlaps the
//                   epilogue of tile i)
//   role 3        : folded-LayerNorm GEMMs only: stages (rstd, -rstd * mean) of the next tile's rows in smem
//   roles 4..     : epilogue, 2 (4 with GELU) warps per TMEM lane quarter, each owning a slice of the BN columns.  Per
//                   32-column chunk: tcgen05.ld -> warp-private smem slab -> re-read transposed so that a warp store
//                   covers 4 rows x 128 contiguous bytes -> +bias | folded-LayerNorm affine [-> GELU] [+ fp32 residual
//                   (cp.async-prefetched for short-K GEMMs) / pos-embedding] -> coalesced bf16 / fp32 global stores
//                   [+ bf16 copy and per-row partial LayerNorm sums for the next GEMM].
//                   The accumulator is handed back with a RELAXED cluster-scope arrive: a release there compiled to
//                   MEMBAR + ERRBAR and waited for every global store in flight (65% -> 82.7% tensor-pipe activity).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include "ptx.cuh"
#include "rowwise.cuh"

namespace vitb200 {

struct GemmEpilogue {
  const float* bias = nullptr;  // [N] fp32 or nullptr
  void* out = nullptr;          // bf16 (kOutF32=false) or fp32 (kOutF32=true), row stride ldo elements
  int ldo = 0;
  const float* resid = nullptr;  // fp32 addend, row stride ldr elements (may alias `out` when kOutF32)
  int ldr = 0;
  // Row remap (used by the patch-embedding GEMM to scatter patch rows behind each image's class token):
  //   g = row / group_rows, i = row % group_rows
  //   out_row   = g * out_group_stride + out_row_offset + i
  //   resid_row = resid_broadcast ? resid_row_offset + i : out_row
  // group_rows == 0 means identity (out_row = resid_row = row).
  int group_rows = 0;
  int out_group_stride = 0;
  int out_row_offset = 0;
  int resid_broadcast = 0;
  int resid_row_offset = 0;
  // LayerNorm folding (see engine.cu, "LayerNorm is folded into the GEMMs"):
  //  * producer side (kResid epilogues, optional): besides the fp32 result also write its bf16 copy `xb` (the A
  //    operand of the next GEMM) and, per row and COLUMN GROUP of one epilogue warp (ln_slot_width(N) = 128 or 64
  //    columns), the partial sums (sum x, sum x^2) into row_stats_out[out_row * stats_slots + col / slot_width] --
  //    fixed slots, no atomics, so the statistics are bit-reproducible.  (Round 1 wrote one slot per 32-column chunk:
  //    4x the bytes, and the consumer needed several dependent L2 round trips per tile to reduce them.)
  //  * consumer side (kLnIn epilogues): out = rstd_m * (acc - mean_m * colsum[n]) + bias[n].  (rstd_m, -rstd_m * mean_m)
  //    is reduced from the producer's partial sums row_stats_in[m * stats_in_slots + s] BY THIS KERNEL (role 3, one tile
  //    ahead of the epilogue, in a fixed slot order: bit-reproducible); round 1 ran a separate finalize kernel in front
  //    of every consuming GEMM (24 launches of ~10 us per ViT-B forward).  The weight operand holds gamma-scaled
  //    weights, `bias` the beta-folded bias, colsum[n] = sum_k W'[n, k].
  __nv_bfloat16* xb = nullptr;
  int ldxb = 0;
  // split-bf16 mode: low halves of the bf16 outputs (out for bf16 epilogues, xb for residual epilogues):
  // lo = bf16(value - float(bf16(value)))
  __nv_bfloat16* out_lo = nullptr;
  __nv_bfloat16* xb_lo = nullptr;
  float2* row_stats_out = nullptr;
  const float2* row_stats_in = nullptr;
  const float* colsum = nullptr;
  // bf16 outputs only: columns >= f16_from_col are written as FP16 instead of bf16 (same 2 bytes per element,
  // saturating at +-65504).  The fused attention kernels multiply fp16 probabilities with the V third of the qkv
  // activation, and kind::f16 wants both operands in one format.  A multiple of 32 (warp-uniform per chunk).
  // (Q and K in fp16 as well was measured: the maps' error against the fp32 oracle moves by < 10 % -- the bf16
  // rounding of the GEMM operands that PRODUCE q and k weighs as much as the rounding of q and k -- not worth bf16's
  // range.)
  int f16_from_col = 0x7fffffff;
  int stats_slots = 0;      // producer side: slots per row of row_stats_out (= N / ln_slot_width(N))
  int stats_in_slots = 0;   // consumer side: slots per row of row_stats_in (= K / slot width, even)
  int stats_in_pairs = 0;   // consumer side: 64-column slots of a K that is a multiple of 256: add adjacent slots first
  float ln_eps = 1e-6f;
};

struct GemmShape {
  int M, N, K;
  // 1: bf16 operands.  3: split-bf16 ("fp32x3") operands: every fp32 operand value is carried as hi + lo (two bf16
  // matrices), and the product as A_hi W_hi + A_lo W_hi + A_hi W_lo -- the main loop runs three K passes over the
  // (hi, hi), (lo, hi), (hi, lo) operand pairs into the same fp32 accumulator (relative error ~2^-17 per product).
  int split = 1;
  int prefetch_w = 0;   // small launches: pull this CTA's W panel into L2 ahead of the grid dependency (W = weights)
};

namespace gemm_cfg {
constexpr int BM = 128;  // rows per CTA (a pair covers 256)
constexpr int BK = 64;   // 64 bf16 = one 128-byte swizzle span
constexpr int kEpiWarp0 = 4;
// Epilogue warps: 2 per TMEM lane quarter normally; the GELU epilogue is issue/latency bound (about 24 instructions
// per output element against a 6144-cycle main loop per tile), so it gets 4 per quarter = 4 per SM sub-partition.
#ifndef VITB200_GELU_EPI_WARPS
#define VITB200_GELU_EPI_WARPS 16
#endif
__host__ __device__ constexpr int epi_warps(bool gelu) { return gelu ? VITB200_GELU_EPI_WARPS : 8; }
constexpr int kSlabStride = 36;                           // floats per slab row: 32 + 4 pad (16-B bank skew)
constexpr int kSlabBytes = 32 * kSlabStride * 4;          // one warp's 32 x 32 fp32 transpose slab
constexpr int kResidStageBytes = 32 * 8 * 16;             // one warp's residual prefetch buffer: 8 x 16 B per lane
constexpr int kLnBufs = 4;                                // folded LayerNorm: the statistics warps run up to three tiles ahead
template <int BN, int kPair, int kEpiWarps, bool kResidStage = false>
struct Cfg {
  static constexpr int kThreads = 32 * (kEpiWarp0 + kEpiWarps);
  static constexpr int kStageBytesA = BM * BK * 2;
  static constexpr int kStageBytesB = (BN / kPair) * BK * 2;
  static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
  static constexpr int kResidBytes = kResidStage ? kEpiWarps * kResidStageBytes : 0;
  static constexpr int kSmemBudget = 227 * 1024 - 256 /*barriers*/ - kLnBufs * BM * 8 /*LN rows*/ - kEpiWarps * kSlabBytes - kResidBytes;
  static constexpr int kStagesFit = kSmemBudget / kStageBytes;
  static constexpr int kStages = kStagesFit > 8 ? 8 : kStagesFit;
  static constexpr int kTmemCols = 2 * BN;  // two accumulator buffers
  static constexpr int kLnBytes = kLnBufs * BM * 8;  // folded LayerNorm: (rstd, -rstd * mean) of the tile's rows, kLnBufs tiles deep
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiWarps * kSlabBytes + 256 + kLnBytes + kResidBytes;
  static_assert(kSmemBytes <= 227 * 1024, "gemm: shared memory budget");
  static_assert(kStages >= 3, "pipeline too shallow");
};
}  // namespace gemm_cfg

__device__ __forceinline__ uint32_t pack2_bf16(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// Exact-erf GELU (torch.nn.GELU default, vision_transformer.py:40-47 MLPBlock): x * Phi(x), Phi the normal CDF.
// With u = |x| and e(u) = erfc(u / sqrt 2) = 2 Phi(-u):   gelu(x) = max(x, 0) - 0.5 * u * e(u).
// e(u) = exp2(-u * q(u)) with q a degree-4 polynomial fitted to -log2(erfc(u / sqrt 2)) / u on [0, 6] (weighted
// minimax; |abs error of gelu| <= 6e-7 for every x, checked in tests/test_gpu_kernels.py): 4 FFMA + 3 FMUL +
// 1 FMNMX + 1 FFMA and ONE MUFU.EX2 per element (Abramowitz-Stegun 7.1.26 needed MUFU.RCP + MUFU.EX2 and ~15
// instructions; the fc1 epilogue has to keep pace with a 6144-cycle MMA main loop per 128x256 tile).
__device__ __forceinline__ float gelu_erf(float x) {
  const float u = fabsf(x);
  float q = fmaf(-4.881035730e-04f, u, 7.198722885e-03f);
  q = fmaf(q, u, -5.214662994e-02f);
  q = fmaf(q, u, -4.595958578e-01f);
  q = fmaf(q, u, -1.151000535e+00f);
  const float e = ptx::ex2_approx(q * u);   // erfc(u / sqrt 2); -> 0 for large u (q * u -> -inf)
  return fmaf(u * e, -0.5f, fmaxf(x, 0.0f));
}

// kGelu / kOutF32 / kResid / kRemap select the epilogue at compile time (the fc1 epilogue is issue-bound: every
// instruction that a runtime flag would leave in its inner loop costs ~1% of the kernel).
// Work decomposition shared by the three roles.  A "unit" is what one persistent scheduler slot (CTA, pair, or
// 4-CTA cluster) processes per iteration.  kPairs == 1: unit i = tile i (m = i / n_tiles, n = i % n_tiles).
// kPairs == 2: units [0, full_rows * n_tiles) are "share W" (pair p takes m = 2 * (i / n_tiles) + p, n = i % n_tiles);
// if m_tiles is odd, the last row of tiles follows as "share A" units (m = m_tiles - 1, pair p takes n = 2 j + p;
// an n beyond the last tile is computed on zero-filled operands and dropped by the epilogue's column check).
struct GemmWork {
  int m_tiles, n_tiles, full_rows, w_units, num_units;
  __device__ GemmWork(int M, int N, int tile_m, int bn, int pairs) {
    m_tiles = (M + tile_m - 1) / tile_m;
    n_tiles = (N + bn - 1) / bn;
    if (pairs == 1) {
      full_rows = m_tiles, w_units = m_tiles * n_tiles, num_units = w_units;
    } else {
      full_rows = m_tiles >> 1, w_units = full_rows * n_tiles;
      num_units = w_units + ((m_tiles & 1) ? ((n_tiles + 1) >> 1) : 0);
    }
  }
  // returns true when the unit shares A between the pairs (else it shares W, or nothing for pairs == 1)
  __device__ bool decode(int unit, int pairs, int pair_id, int& m_blk, int& n_blk) const {
    if (pairs == 1) {
      m_blk = unit / n_tiles, n_blk = unit - m_blk * n_tiles;
      return false;
    }
    if (unit < w_units) {
      const int row = unit / n_tiles;
      m_blk = 2 * row + pair_id, n_blk = unit - row * n_tiles;
      return false;
    }
    m_blk = m_tiles - 1, n_blk = 2 * (unit - w_units) + pair_id;
    return true;
  }
};

template <int BN, int kPair, int kPairs, bool kGelu, bool kOutF32, bool kResid, bool kRemap, bool kLnIn = false,
          bool kPrefetch = false>
__global__ void __launch_bounds__((gemm_cfg::Cfg<BN, kPair, gemm_cfg::epi_warps(kGelu), kPrefetch>::kThreads), 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                 const __grid_constant__ CUtensorMap tmap_a_lo, const __grid_constant__ CUtensorMap tmap_w_lo,
                 GemmShape shape, GemmEpilogue ep) {
  using namespace gemm_cfg;
  constexpr int kEpiWarps = epi_warps(kGelu);
  constexpr int kColGroups = kEpiWarps / 4;  // warps per TMEM lane quarter; each owns BN / kColGroups columns
  static_assert(!kPrefetch || kResid, "the residual prefetch belongs to residual epilogues");
  using C = Cfg<BN, kPair, kEpiWarps, kPrefetch>;
  constexpr int kStages = C::kStages;
  constexpr int kTileM = BM * kPair;

  // Dynamic smem starts 1024-B aligned (it follows the 1 KB the driver reserves); keeping the array typed lets
  // the compiler emit LDS/STS (not generic LD/ST) for the epilogue slabs.
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();
  float* slabs = reinterpret_cast<float*>(smem + kStages * C::kStageBytes);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * C::kStageBytes + kEpiWarps * kSlabBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint64_t* ln_full_bar = tmem_empty_bar + 2;   // [kLnBufs] (rstd, -rstd * mean) of a tile's rows staged by warps 2 and 3
  uint64_t* ln_empty_bar = ln_full_bar + kLnBufs;     // [kLnBufs] ... and read by every epilogue warp
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ln_empty_bar + kLnBufs);
  static_assert((2 * 8 + 4 + 2 * kLnBufs) * 8 + 8 <= 256, "barrier area");
  float2* ln_rows = reinterpret_cast<float2*>(smem + kStages * C::kStageBytes + kEpiWarps * kSlabBytes + 256);  // [2][BM]
  float4* resid_stage = reinterpret_cast<float4*>(smem + kStages * C::kStageBytes + kEpiWarps * kSlabBytes + 256 + C::kLnBytes);

  // Role index, not the hardware warp id: the four control roles (0 TMA, 1 MMA issue, 2 TMEM allocator, 3 LayerNorm
  // rows) sit on the HIGHEST hardware warp ids, the epilogue roles (4..) on the lowest.  The warp scheduler favours
  // higher warp ids among eligible warps, and the TMA / MMA issue loops are the ones nobody should be queueing in
  // front of.  role % 4 == hardware warp % 4, so the TMEM lane-quarter rule is unaffected.
  const int warp = static_cast<int>((threadIdx.x >> 5) + kEpiWarp0) % (C::kThreads / 32);
  const int lane = threadIdx.x & 31;
  static_assert(kPairs == 1 || (kPair == 2 && BN == 256), "4-CTA clusters: pair kernels with BN = 256 only");
  constexpr int kClusterCtas = kPair * kPairs;
  const uint32_t cluster_rank = (kClusterCtas > 1) ? ptx::cluster_ctarank() : 0u;
  const uint32_t cta_rank = cluster_rank & (kPair - 1);    // rank inside the CTA pair
  const int pair_id = static_cast<int>(cluster_rank >> 1);  // which pair of the cluster (0 when kPairs == 1)
  const bool leader = cta_rank == 0;
  const int slot = blockIdx.x / kClusterCtas;               // persistent scheduler slot: CTA, pair or cluster
  const int num_slots = gridDim.x / kClusterCtas;

  const GemmWork work(shape.M, shape.N, kTileM, BN, kPairs);
  const int num_units = work.num_units;
  const int num_kb = (shape.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_w);
    if (shape.split > 1) {
      ptx::prefetch_tmap(&tmap_a_lo);
      ptx::prefetch_tmap(&tmap_w_lo);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], kPairs);  // a stage is written for both pairs: both pairs' MMAs must release it
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full_bar[a], 1);
      ptx::mbar_init(&tmem_empty_bar[a], kEpiWarps * kPair);  // one elected lane per epilogue warp (of both CTAs)
    }
    for (int a = 0; a < kLnBufs; ++a) {
      ptx::mbar_init(&ln_full_bar[a], 2);   // the two statistics warps (roles 2 and 3), 64 rows each
      ptx::mbar_init(&ln_empty_bar[a], kEpiWarps);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    if (kPair == 2) ptx::tmem_alloc_pair<C::kTmemCols>(tmem_slot);
    else ptx::tmem_alloc<C::kTmemCols>(tmem_slot);
  }
  ptx::tc_fence_before();
  if (kClusterCtas > 1) ptx::cluster_sync();  // peer barriers initialised before any multicast commit / remote arrive
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ptx::grid_dep_launch();   // PDL: the next kernel's CTAs may take the SMs this grid leaves free / frees at its tail ...
  // Small launches (one unit per CTA: the single-image request): W is a weight matrix, not a result of the previous
  // kernel, so its panel for this CTA's unit is pulled into L2 while that kernel is still running -- at B = 1 a GEMM is
  // 12-48 CTAs each streaming its W panel at what ONE SM can keep in flight against HBM latency (fc2: 786 KB per CTA,
  // 19 us); from L2 the same ring of loads turns around ~2.5 x faster.
  if (kPairs == 1 && shape.prefetch_w && warp == 0 && num_units <= num_slots && slot < num_units) {
    int m_blk, n_blk;
    work.decode(slot, kPairs, pair_id, m_blk, n_blk);
    const int w_row = n_blk * BN + static_cast<int>(cta_rank) * (BN / kPair);
    for (int kb = lane; kb < num_kb; kb += 32) ptx::tma_prefetch_l2_2d(&tmap_w, kb * BK, w_row);
  }
  ptx::grid_dep_wait();     // ... and everything below waits for the previous kernel's results

  // The producer and MMA loops are executed by WHOLE warps with one elected lane issuing: loop counters, smem
  // addresses and descriptors then stay warp-uniform (uniform registers), instead of being moved from vector to
  // uniform registers (R2UR) in front of every UTMALDG / UTCHMMA, which throttles a single-thread issue loop.
  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (every CTA)
    int stage = 0;
    uint32_t phase = 0;
    // kPairs == 2: the counterpart of this CTA in the other pair has the same rank-in-pair
    const uint16_t mc_mask = static_cast<uint16_t>((1u << cta_rank) | (1u << (2 + cta_rank)));
    for (int unit = slot; unit < num_units; unit += num_slots) {
      int m_blk, n_blk;
      const bool share_a = work.decode(unit, kPairs, pair_id, m_blk, n_blk);
      const int a_row = m_blk * kTileM + static_cast<int>(cta_rank) * BM;
      const int w_row = n_blk * BN + static_cast<int>(cta_rank) * (BN / kPair);
      for (int pass = 0; pass < shape.split; ++pass) {
      // operand pair of this K pass: (hi, hi), (lo, hi), (hi, lo)
      const CUtensorMap* ta = pass == 1 ? &tmap_a_lo : &tmap_a;
      const CUtensorMap* tw = pass == 2 ? &tmap_w_lo : &tmap_w;
      for (int kb = 0; kb < num_kb; ++kb) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * C::kStageBytes;
        uint8_t* sb = sa + C::kStageBytesA;
        if (ptx::elect_one()) {
          if (kPairs == 2) {
            // 64-row boxes.  Own operand: both halves; shared operand: this pair's half, multicast to the counterpart.
            constexpr int kHalf = 64 * BK * 2;
            if (leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * C::kStageBytes);
            if (!share_a) {
              ptx::tma_load_2d_pair(sa, ta, &full_bar[stage], kb * BK, a_row);
              ptx::tma_load_2d_pair(sa + kHalf, ta, &full_bar[stage], kb * BK, a_row + 64);
              ptx::tma_load_2d_pair_mc(sb + pair_id * kHalf, tw, &full_bar[stage], kb * BK, w_row + pair_id * 64,
                                       mc_mask);
            } else {
              ptx::tma_load_2d_pair_mc(sa + pair_id * kHalf, ta, &full_bar[stage], kb * BK, a_row + pair_id * 64,
                                       mc_mask);
              ptx::tma_load_2d_pair(sb, tw, &full_bar[stage], kb * BK, w_row);
              ptx::tma_load_2d_pair(sb + kHalf, tw, &full_bar[stage], kb * BK, w_row + 64);
            }
          } else if (kPair == 2) {
            // one barrier (the leader's) tracks the bytes of both CTAs' halves
            if (leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * C::kStageBytes);
            ptx::tma_load_2d_pair(sa, ta, &full_bar[stage], kb * BK, a_row);
            ptx::tma_load_2d_pair(sb, tw, &full_bar[stage], kb * BK, w_row);
          } else {
            ptx::mbar_arrive_expect_tx(&full_bar[stage], C::kStageBytes);
            ptx::tma_load_2d(sa, ta, &full_bar[stage], kb * BK, a_row);
            ptx::tma_load_2d(sb, tw, &full_bar[stage], kb * BK, w_row);
          }
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      }  // K passes
    }
  } else if (warp == 1 && leader) {
    // ------------------------------------------------------------ UMMA issuer (leader CTA)
    constexpr uint32_t idesc = ptx::make_idesc_bf16(kTileM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int local = 0;
    // commits: the smem stage is released in every CTA of the cluster, the accumulator goes to this pair's epilogues
    constexpr uint16_t kEmptyMask = static_cast<uint16_t>((1u << kClusterCtas) - 1);
    const uint16_t full_mask = static_cast<uint16_t>(0x3u << (2 * pair_id));
    for (int unit = slot; unit < num_units; unit += num_slots, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      ptx::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
      ptx::tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      const int total_kb = num_kb * shape.split;  // split-bf16: three K passes into the same accumulator
      for (int kb = 0; kb < total_kb; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        const uint32_t sa = ptx::smem_u32(smem + stage * C::kStageBytes);
        const uint32_t sb = sa + C::kStageBytesA;
        const uint64_t da = ptx::make_smem_desc_sw128(sa, 16, 1024);
        const uint64_t db = ptx::make_smem_desc_sw128(sb, 16, 1024);
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // +32 bytes per K=16 step inside the 128-B swizzle span (descriptor address is in 16-B units)
            if (kPair == 2) ptx::umma_bf16_ss_pair(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            else ptx::umma_bf16_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          // smem slot reusable (in both CTAs) once these MMAs retire
          if (kPair == 2) ptx::umma_commit_pair(&empty_bar[stage], kEmptyMask);
          else ptx::umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      // accumulator complete -> epilogue warps (of both CTAs)
      if (ptx::elect_one()) {
        if (kPair == 2) ptx::umma_commit_pair(&tmem_full_bar[acc], full_mask);
        else ptx::umma_commit(&tmem_full_bar[acc]);
      }
      __syncwarp();
    }
  } else if (kLnIn && (warp == 3 || warp == 2)) {
    // ------------------------------------------------------------ LayerNorm row statistics (folded-LN GEMMs)
    // Runs one tile ahead of the epilogue: reduces the partial sums (sum x, sum x^2 per 32-column chunk, written by the
    // GEMM that produced the rows) of the tile's 128 rows (this CTA's) to (rstd, -rstd * mean) and stages them in smem,
    // so that the epilogue never waits for an L2 round trip (exposed, that latency cost ~2,000 cycles per tile).
    // Four lanes per row: lane q of a row's quad sums the float4 pairs of slots q, q + 4, ... in ascending order, then
    // two xor-shuffles -- the order is fixed, so the statistics are bit-reproducible (and identical to what round 1's
    // separate row_stats_finalize kernel produced).  With one slot per 128 (64) columns a ViT-B row has 6 slots = 3
    // float4: lanes q < 3 of a row's quad load one each, so a tile costs one L2 round trip of 16 independent loads per
    // lane (2.4 MB per forward pass over the statistics, L2-resident behind the producing GEMM).
    // (roles 2 and 3 split the tile's 128 rows: with short K -- ViT-S, six k-blocks per tile -- ONE warp reducing all 128
    //  rows took as long as the tile's main loop and the qkv / fc1 GEMMs ran 50 % / 15 % slower than with round 1's
    //  separate finalize kernel: measured)
    int local = 0;
    const int half = warp - 2;
    const int q = lane & 3, rq = lane >> 2;
    // stats_in_pairs: the slots are 64 columns wide although K is a multiple of 256 (small batches, see engine.cu:
    // stats_width): adjacent slots are added first -- exactly the 128-column slot the wide tiles write -- and the rest of
    // the reduction is the same, so the result does not depend on the slot width.
    const bool pairs = ep.stats_in_pairs != 0;
    const int n4 = pairs ? ep.stats_in_slots >> 2 : ep.stats_in_slots >> 1;   // float4 = two (combined) slots
    const float inv_w = 1.0f / static_cast<float>(shape.K);
    for (int unit = slot; unit < num_units; unit += num_slots, ++local) {
      int m_blk, n_blk;
      work.decode(unit, kPairs, pair_id, m_blk, n_blk);
      const int buf = local % kLnBufs;
      ptx::mbar_wait(&ln_empty_bar[buf], ((local / kLnBufs) & 1) ^ 1);
      const int row0 = m_blk * kTileM + static_cast<int>(cta_rank) * BM;
      if (pairs) {
        // small batches only.  Same order of additions as below with u = (slot 2i) + (slot 2i+1) in place of the 128-column
        // slot i; 8 row groups (16 independent 16-byte loads per lane) per L2 round trip
        constexpr int kG = 8;
#pragma unroll 1
        for (int g0 = half * (BM / 16); g0 < (half + 1) * (BM / 16); g0 += kG) {
          float s1[kG], s2[kG];
#pragma unroll
          for (int g = 0; g < kG; ++g) s1[g] = 0.f, s2[g] = 0.f;
          for (int j = q; j < n4; j += 4) {
            float4 lo[kG], hi[kG];
#pragma unroll
            for (int g = 0; g < kG; ++g) {
              const int row = row0 + (g0 + g) * 8 + rq;
              lo[g] = make_float4(0.f, 0.f, 0.f, 0.f), hi[g] = lo[g];
              if (row < shape.M) {
                const float4* sp = reinterpret_cast<const float4*>(ep.row_stats_in + static_cast<long>(row) * ep.stats_in_slots);
                lo[g] = sp[2 * j], hi[g] = sp[2 * j + 1];
              }
            }
#pragma unroll
            for (int g = 0; g < kG; ++g)
              s1[g] += lo[g].x + lo[g].z, s2[g] += lo[g].y + lo[g].w, s1[g] += hi[g].x + hi[g].z, s2[g] += hi[g].y + hi[g].w;
          }
#pragma unroll
          for (int g = 0; g < kG; ++g) {
            float a = s1[g], b = s2[g];
            a += __shfl_xor_sync(0xffffffffu, a, 1), b += __shfl_xor_sync(0xffffffffu, b, 1);
            a += __shfl_xor_sync(0xffffffffu, a, 2), b += __shfl_xor_sync(0xffffffffu, b, 2);
            if (q == 0) {
              const float mean = a * inv_w;
              const float rstd = rsqrtf(fmaxf(b * inv_w - mean * mean, 0.f) + ep.ln_eps);
              ln_rows[buf * BM + (g0 + g) * 8 + rq] = make_float2(rstd, -rstd * mean);
            }
          }
        }
      }
      constexpr int kGroups = 8;             // this warp's 8 row groups (of 8 rows) at once: ONE L2 round trip per tile and j
#pragma unroll 1
      for (int g0 = half * (BM / 16); g0 < (pairs ? 0 : (half + 1) * (BM / 16)); g0 += kGroups) {
        float s1[kGroups], s2[kGroups];
#pragma unroll
        for (int g = 0; g < kGroups; ++g) s1[g] = 0.f, s2[g] = 0.f;
        for (int j = q; j < n4; j += 4) {
          float4 v[kGroups];
#pragma unroll
          for (int g = 0; g < kGroups; ++g) {
            const int row = row0 + (g0 + g) * 8 + rq;
            v[g] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < shape.M)
              v[g] = reinterpret_cast<const float4*>(ep.row_stats_in + static_cast<long>(row) * ep.stats_in_slots)[j];
          }
#pragma unroll
          for (int g = 0; g < kGroups; ++g) s1[g] += v[g].x, s2[g] += v[g].y, s1[g] += v[g].z, s2[g] += v[g].w;
        }
#pragma unroll
        for (int g = 0; g < kGroups; ++g) {
          float a = s1[g], b = s2[g];
          a += __shfl_xor_sync(0xffffffffu, a, 1), b += __shfl_xor_sync(0xffffffffu, b, 1);
          a += __shfl_xor_sync(0xffffffffu, a, 2), b += __shfl_xor_sync(0xffffffffu, b, 2);
          if (q == 0) {
            const float mean = a * inv_w;
            const float rstd = rsqrtf(fmaxf(b * inv_w - mean * mean, 0.f) + ep.ln_eps);
            ln_rows[buf * BM + (g0 + g) * 8 + rq] = make_float2(rstd, -rstd * mean);
          }
        }
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&ln_full_bar[buf]);
    }
  } else if (warp >= kEpiWarp0) {
    // ------------------------------------------------------------ epilogue
    const int quarter = warp & 3;                   // TMEM lane quarter this warp may access
    const int col_grp = (warp - kEpiWarp0) >> 2;    // which group of the BN accumulator columns
    constexpr int kGroupCols = BN / kColGroups;
    constexpr int kChunks = kGroupCols / 32;        // 32-column chunks per warp
    static_assert(kChunks >= 1, "column group narrower than one 32-column chunk");
    float* slab = slabs + (warp - kEpiWarp0) * (32 * kSlabStride);
    const int trow = lane >> 3;                     // transposed mapping: rows trow + 4 i, 4 columns at 4 * tcol
    const int tcol = lane & 7;
    // Residual epilogues of short-K GEMMs (kPrefetch; out_proj): the fp32 addend of chunk c+1 (or of the next tile's
    // first chunk) is fetched with cp.async into a per-thread staging area while chunk c is processed.  (Long-K GEMMs
    // such as fc2 hide the load behind their main loop already and prefer the extra smem stage: measured.)  Loading it at the point of use left one HBM round
    // trip exposed per chunk: 4 KB in flight per warp bounded out_proj at 4.3 TB/s.  Every thread reads back only
    // what it staged itself, so no warp synchronisation is involved.
    float4* rstage = resid_stage + (warp - kEpiWarp0) * (kResidStageBytes / 16);
    auto prefetch_resid = [&](int unit2, int c2) {
      if (unit2 < num_units) {
        int m2, n2;
        work.decode(unit2, kPairs, pair_id, m2, n2);
        const int col = n2 * BN + col_grp * kGroupCols + c2 * 32 + 4 * tcol;
        const int rb = m2 * kTileM + static_cast<int>(cta_rank) * BM + quarter * 32;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = rb + trow + 4 * i;
          if (row < shape.M && col < shape.N) {
            long resid_row = row;
            if (kRemap) {
              const int g = row / ep.group_rows;
              const int gi = row - g * ep.group_rows;
              resid_row = ep.resid_broadcast ? (ep.resid_row_offset + gi)
                                             : static_cast<long>(g) * ep.out_group_stride + ep.out_row_offset + gi;
            }
            ptx::cp_async_16(&rstage[i * 32 + lane], ep.resid + resid_row * ep.ldr + col);
          }
        }
      }
      ptx::cp_async_commit();
    };
    if (kPrefetch) prefetch_resid(slot, 0);
    int local = 0;
    for (int unit = slot; unit < num_units; unit += num_slots, ++local) {
      int m_blk, n_blk;
      work.decode(unit, kPairs, pair_id, m_blk, n_blk);
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const int row_base = m_blk * kTileM + static_cast<int>(cta_rank) * BM + quarter * 32;
      // folded LayerNorm: scale a = rstd and offset c = -rstd * mean of this thread's 8 rows (trow + 4 i), staged by warp 3
      float ln_a[8], ln_c[8];
      if (kLnIn) {
        const int lbuf = local % kLnBufs;
        ptx::mbar_wait(&ln_full_bar[lbuf], (local / kLnBufs) & 1);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 v = ln_rows[lbuf * BM + quarter * 32 + trow + 4 * i];
          ln_a[i] = v.x, ln_c[i] = v.y;
        }
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&ln_empty_bar[lbuf]);
      }
      // bias (and folded-LayerNorm column sums) of every chunk of this tile, fetched before the accumulator wait: loaded
      // inside the chunk loop their L2 latency was exposed once per chunk (9 % of the epilogue's stall samples)
      float4 bias_c[kChunks], csum_c[kChunks];
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        const int col = n_blk * BN + col_grp * kGroupCols + c * 32 + 4 * tcol;
        bias_c[c] = make_float4(0.f, 0.f, 0.f, 0.f), csum_c[c] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col < shape.N) {
          if (ep.bias != nullptr) bias_c[c] = *reinterpret_cast<const float4*>(ep.bias + col);
          if (kLnIn) csum_c[c] = *reinterpret_cast<const float4*>(ep.colsum + col);
        }
      }
      // producer side of the LayerNorm folding: this lane's row (trow + 4 * tcol of the quarter) summed over the warp's chunks.
      // A 128-column slot is formed as (chunk 0 + chunk 1) + (chunk 2 + chunk 3), i.e. as the sum of the two 64-column
      // slots a 128-wide tile would have written: the statistics -- and with them every output of the forward -- are then
      // bit-identical whichever tile width the engine picked for the batch size (stats_width in engine.cu).
      float st1 = 0.f, st2 = 0.f, stb1 = 0.f, stb2 = 0.f;
      ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t taddr0 =
          tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + col_grp * kGroupCols;
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        uint32_t r[32];
        ptx::tmem_ld_x32(taddr0 + c * 32, r);
        ptx::tmem_ld_wait();
        if (c + 1 == kChunks) {
          // every TMEM read of this accumulator has landed in registers: hand it back to the MMA issuer
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (kPair == 2) ptx::mbar_arrive_cluster(&tmem_empty_bar[acc], cluster_rank & ~1u);
            else ptx::mbar_arrive(&tmem_empty_bar[acc]);
          }
        }
        const int col0 = n_blk * BN + col_grp * kGroupCols + c * 32;
        if (col0 >= shape.N) {  // warp-uniform
          if (kPrefetch) {  // keep the prefetch chain going
            ptx::cp_async_wait_all();
            prefetch_resid(c + 1 < kChunks ? unit : unit + num_slots, c + 1 < kChunks ? c + 1 : 0);
          }
          continue;
        }
        // registers (thread = row, 32 columns) -> slab
        float* srow = slab + lane * kSlabStride;
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<uint4*>(srow + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
        __syncwarp();
        const int col = col0 + 4 * tcol;
        const bool col_ok = col < shape.N;
        const float4 bias4 = bias_c[c], csum4 = csum_c[c];
        // transposed pass: 8 rows (trow + 4 i) x 4 columns per thread.  With a residual all 8 rows form one batch so
        // that 8 independent 16-byte loads per thread are in flight (the out_proj / fc2 epilogues are bound by
        // HBM latency x outstanding bytes); otherwise two batches of 4 keep the instruction footprint small.
        constexpr int kRowBatch = kResid ? 8 : 4;
        const float* sl = slab + trow * kSlabStride + 4 * tcol;
        // Addresses: without a row remap the 8 rows of a thread are 4 * ld apart, so every pointer is one 64-bit base per
        // chunk plus a 32-bit multiple of the row step (one IMAD.WIDE per access; forming row * ld + col in 64 bits per
        // row cost ~17 integer instructions per 4 outputs -- 40 % of the GELU epilogue's instruction stream).
        constexpr unsigned kOutEsz = kOutF32 ? 4u : 2u;
        const long first_row = row_base + trow;
        char* const out0 = reinterpret_cast<char*>(ep.out) + (first_row * ep.ldo + col) * kOutEsz;
        const unsigned out_st = 4u * static_cast<unsigned>(ep.ldo) * kOutEsz;
        const char* const res0 = reinterpret_cast<const char*>(ep.resid) + (first_row * ep.ldr + col) * 4;
        const unsigned res_st = 16u * static_cast<unsigned>(ep.ldr);
        char* const xb0 = reinterpret_cast<char*>(ep.xb) + (first_row * ep.ldxb + col) * 2;
        const unsigned xb_st = 8u * static_cast<unsigned>(ep.ldxb);
        // (unrolled for the folded-LayerNorm epilogue: its per-row scale / offset arrays need static indices)
        constexpr int kBatchUnroll = kLnIn ? 8 / kRowBatch : 1;
#pragma unroll kBatchUnroll
        for (int i0 = 0; i0 < 8; i0 += kRowBatch) {
          float4 v[kRowBatch], q[kRowBatch];
          long orow[kRowBatch];
          bool ok[kRowBatch];
          float p1[kRowBatch], p2[kRowBatch];  // producer side of the LayerNorm folding: per-row partial sums
#pragma unroll
          for (int i = 0; i < kRowBatch; ++i) {
            const int row = row_base + trow + 4 * (i0 + i);
            v[i] = *reinterpret_cast<const float4*>(sl + 4 * (i0 + i) * kSlabStride);
            ok[i] = row < shape.M && col_ok;
            long out_row = row, resid_row = row;
            if (kRemap) {
              const int g = row / ep.group_rows;
              const int gi = row - g * ep.group_rows;
              out_row = static_cast<long>(g) * ep.out_group_stride + ep.out_row_offset + gi;
              resid_row = ep.resid_broadcast ? (ep.resid_row_offset + gi) : out_row;
            }
            orow[i] = out_row;
            if (kResid && !kPrefetch) {
              q[i] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (ok[i])
                q[i] = kRemap ? *reinterpret_cast<const float4*>(ep.resid + resid_row * ep.ldr + col)
                              : *reinterpret_cast<const float4*>(res0 + static_cast<unsigned>(i0 + i) * res_st);
            }
          }
          if (kPrefetch) {
            ptx::cp_async_wait_all();
#pragma unroll
            for (int i = 0; i < kRowBatch; ++i) q[i] = ok[i] ? rstage[(i0 + i) * 32 + lane] : make_float4(0.f, 0.f, 0.f, 0.f);
            prefetch_resid(c + 1 < kChunks ? unit : unit + num_slots, c + 1 < kChunks ? c + 1 : 0);
          }
#pragma unroll
          for (int i = 0; i < kRowBatch; ++i) {
            float4 t = v[i];
            if (kLnIn) {
              const float a = ln_a[i0 + i], cc = ln_c[i0 + i];
              t.x = fmaf(t.x, a, fmaf(cc, csum4.x, bias4.x)), t.y = fmaf(t.y, a, fmaf(cc, csum4.y, bias4.y));
              t.z = fmaf(t.z, a, fmaf(cc, csum4.z, bias4.z)), t.w = fmaf(t.w, a, fmaf(cc, csum4.w, bias4.w));
            } else {
              t.x += bias4.x, t.y += bias4.y, t.z += bias4.z, t.w += bias4.w;
            }
            if (kGelu) t.x = gelu_erf(t.x), t.y = gelu_erf(t.y), t.z = gelu_erf(t.z), t.w = gelu_erf(t.w);
            if (kResid) t.x += q[i].x, t.y += q[i].y, t.z += q[i].z, t.w += q[i].w;
            if (kResid && ep.xb != nullptr) {
              // bf16 copy for the next GEMM; this thread's share of the row's partial sums (reduced after the loop)
              p1[i] = ok[i] ? quad_sum(t) : 0.f;      // the canonical order of rowwise.cuh ("LayerNorm partial sums")
              p2[i] = ok[i] ? quad_sumsq(t) : 0.f;
              if (ok[i]) {
                const __nv_bfloat162 h01 = __floats2bfloat162_rn(t.x, t.y), h23 = __floats2bfloat162_rn(t.z, t.w);
                uint2 pk;
                pk.x = *reinterpret_cast<const uint32_t*>(&h01), pk.y = *reinterpret_cast<const uint32_t*>(&h23);
                char* const xbp = kRemap ? reinterpret_cast<char*>(ep.xb + orow[i] * ep.ldxb + col)
                                         : xb0 + static_cast<unsigned>(i0 + i) * xb_st;
                *reinterpret_cast<uint2*>(xbp) = pk;
                if (ep.xb_lo != nullptr) {
                  uint2 pl;
                  pl.x = pack2_bf16(t.x - __low2float(h01), t.y - __high2float(h01));
                  pl.y = pack2_bf16(t.z - __low2float(h23), t.w - __high2float(h23));
                  *reinterpret_cast<uint2*>(reinterpret_cast<char*>(ep.xb_lo) + (xbp - reinterpret_cast<char*>(ep.xb))) = pl;
                }
              }
            }
            if (ok[i]) {
              char* const outp = kRemap ? reinterpret_cast<char*>(ep.out) + (orow[i] * ep.ldo + col) * kOutEsz
                                        : out0 + static_cast<unsigned>(i0 + i) * out_st;
              if (kOutF32) {
                *reinterpret_cast<float4*>(outp) = t;
              } else {
                __nv_bfloat162 lo = __floats2bfloat162_rn(t.x, t.y);
                __nv_bfloat162 hi = __floats2bfloat162_rn(t.z, t.w);
                uint2 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&lo);
                pk.y = *reinterpret_cast<uint32_t*>(&hi);
                if (!kGelu && col0 >= ep.f16_from_col) {   // warp-uniform (the qkv GEMM's V third; never a GELU epilogue)
                  pk.x = pack_f16x2_sat(t.x, t.y), pk.y = pack_f16x2_sat(t.z, t.w);
                }
                *reinterpret_cast<uint2*>(outp) = pk;
                if (ep.out_lo != nullptr) {  // split-bf16: the rounding residue as a second bf16 matrix
                  uint2 pl;
                  pl.x = pack2_bf16(t.x - __low2float(lo), t.y - __high2float(lo));
                  pl.y = pack2_bf16(t.z - __low2float(hi), t.w - __high2float(hi));
                  *reinterpret_cast<uint2*>(reinterpret_cast<char*>(ep.out_lo) + (outp - reinterpret_cast<char*>(ep.out))) = pl;
                }
              }
            }
          }
          if (kResid && ep.xb != nullptr) {
            // The 8 lanes that share a row group (tcol = 0..7, consecutive lanes) hold 8 rows x (sum, sum of squares)
            // each.  Recursive halving: exchange the half of the rows the partner keeps (xor 4, 2, 1), add, so that
            // lane tcol ends with the complete sums of row trow + 4 * tcol: 14 shuffles instead of 48.
            static_assert(!kResid || kRowBatch == 8, "the residual epilogue reduces all 8 rows of a chunk at once");
            float a1[4], a2[4];
            const bool b2 = (tcol & 4) != 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float s1 = b2 ? p1[i] : p1[i + 4], s2 = b2 ? p2[i] : p2[i + 4];   // rows the partner keeps
              const float k1 = b2 ? p1[i + 4] : p1[i], k2 = b2 ? p2[i + 4] : p2[i];
              a1[i] = k1 + __shfl_xor_sync(0xffffffffu, s1, 4);
              a2[i] = k2 + __shfl_xor_sync(0xffffffffu, s2, 4);
            }
            float c1[2], c2[2];
            const bool b1 = (tcol & 2) != 0;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const float s1 = b1 ? a1[i] : a1[i + 2], s2 = b1 ? a2[i] : a2[i + 2];
              const float k1 = b1 ? a1[i + 2] : a1[i], k2 = b1 ? a2[i + 2] : a2[i];
              c1[i] = k1 + __shfl_xor_sync(0xffffffffu, s1, 2);
              c2[i] = k2 + __shfl_xor_sync(0xffffffffu, s2, 2);
            }
            const bool b0 = (tcol & 1) != 0;
            const float f1 = (b0 ? c1[1] : c1[0]) + __shfl_xor_sync(0xffffffffu, b0 ? c1[0] : c1[1], 1);
            const float f2 = (b0 ? c2[1] : c2[0]) + __shfl_xor_sync(0xffffffffu, b0 ? c2[0] : c2[1], 1);
            // this lane now owns row index tcol of the batch (bit 2 chose rows 4..7, bit 1 rows +2, bit 0 rows +1):
            // chunk sums are added in ascending chunk order
            if (kGroupCols == 128 && c >= 2) stb1 += f1, stb2 += f2;
            else st1 += f1, st2 += f2;
          }
        }
        __syncwarp();  // slab is rewritten by the next chunk
      }
      if (kResid && ep.xb != nullptr) {
        static_assert(!kResid || (kGroupCols == 64 || kGroupCols == 128), "one statistics slot per epilogue-warp column group");
        const int row = row_base + trow + 4 * tcol;
        const int gcol = n_blk * BN + col_grp * kGroupCols;
        if (row < shape.M && gcol < shape.N) {
          long out_row = row;
          if (kRemap) {
            const int g = row / ep.group_rows;
            out_row = static_cast<long>(g) * ep.out_group_stride + ep.out_row_offset + (row - g * ep.group_rows);
          }
          ep.row_stats_out[out_row * ep.stats_slots + gcol / kGroupCols] = make_float2(st1 + stb1, st2 + stb2);
        }
      }
    }
  }

  __syncwarp();
  ptx::tc_fence_before();
  if (kClusterCtas > 1) ptx::cluster_sync();  // the peer may still be signalling this CTA's barriers / reading its smem
  else __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    if (kPair == 2) ptx::tmem_dealloc_pair<C::kTmemCols>(tmem_base);
    else ptx::tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

}  // namespace vitb200
