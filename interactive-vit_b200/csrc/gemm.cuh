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
//
// Roles (4 + 8 warps, or 4 + 16 for the GELU epilogue; 1 CTA / SM, persistent over output tiles):
//   warp 0 lane 0 : TMA producer  (A tile 128x64, W tile (BN/kPair)x64, SWIZZLE_128B, mbarrier ring; in a pair both
//                   CTAs' loads complete_tx on the LEADER's full barrier)
//   warp 1 lane 0 : UMMA issuer (leader CTA only in a pair): 4 x tcgen05.mma (K = 16) per stage, commit ->
//                   empty barrier of both CTAs; accumulator-complete commit -> both CTAs' epilogues
//   warp 2        : TMEM allocator (2 accumulator buffers of BN fp32 columns: MMA of tile i+1 overlaps the
//                   epilogue of tile i)
//   warps 4..     : epilogue, 2 (4 with GELU) warps per TMEM lane quarter, each owning a slice of the BN columns.  Per 32-column
//                   chunk: tcgen05.ld -> warp-private smem slab -> re-read transposed so that a warp store
//                   covers 4 rows x 128 contiguous bytes -> +bias [-> GELU] [+ fp32 residual / pos-embedding]
//                   -> coalesced bf16 / fp32 global stores.
#pragma once
#include <cuda.h>
#include "ptx.cuh"

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
};

struct GemmShape {
  int M, N, K;
};

namespace gemm_cfg {
constexpr int BM = 128;  // rows per CTA (a pair covers 256)
constexpr int BK = 64;   // 64 bf16 = one 128-byte swizzle span
constexpr int kEpiWarp0 = 4;
// Epilogue warps: 2 per TMEM lane quarter normally; the GELU epilogue is issue/latency bound (about 24 instructions
// per output element against a 6144-cycle main loop per tile), so it gets 4 per quarter = 4 per SM sub-partition.
__host__ __device__ constexpr int epi_warps(bool gelu) { return gelu ? 16 : 8; }
constexpr int kSlabStride = 36;                           // floats per slab row: 32 + 4 pad (16-B bank skew)
constexpr int kSlabBytes = 32 * kSlabStride * 4;          // one warp's 32 x 32 fp32 transpose slab
template <int BN, int kPair, int kEpiWarps>
struct Cfg {
  static constexpr int kThreads = 32 * (kEpiWarp0 + kEpiWarps);
  static constexpr int kStageBytesA = BM * BK * 2;
  static constexpr int kStageBytesB = (BN / kPair) * BK * 2;
  static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
  static constexpr int kSmemBudget = 227 * 1024 - 256 /*barriers*/ - kEpiWarps * kSlabBytes;
  static constexpr int kStagesFit = kSmemBudget / kStageBytes;
  static constexpr int kStages = kStagesFit > 8 ? 8 : kStagesFit;
  static constexpr int kTmemCols = 2 * BN;  // two accumulator buffers
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiWarps * kSlabBytes + 256;
  static_assert(kStages >= 3, "pipeline too shallow");
};
}  // namespace gemm_cfg

// Exact-erf GELU (torch.nn.GELU default, vision_transformer.py:40-47 MLPBlock): 0.5 x (1 + erf(x / sqrt 2)).
// erf by Abramowitz-Stegun 7.1.26 (|abs error| <= 1.5e-7, i.e. fp32 round-off level) evaluated with one
// MUFU.RCP and one MUFU.EX2 instead of the ~25-instruction libdevice erff: the fc1 epilogue must keep pace with
// a 6144-cycle MMA main loop per 128x256 tile.
__device__ __forceinline__ float gelu_erf(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = ptx::rcp_approx(fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  poly *= t;
  const float e = ptx::ex2_approx(-1.4426950408889634f * z * z);
  const float erf_abs = fmaf(-poly, e, 1.0f);           // erf(|x|/sqrt2)
  const float half_x = 0.5f * x;
  return fmaf(copysignf(erf_abs, x), half_x, half_x);   // 0.5 x (1 + erf)
}

// kGelu / kOutF32 / kResid / kRemap select the epilogue at compile time (the fc1 epilogue is issue-bound: every
// instruction that a runtime flag would leave in its inner loop costs ~1% of the kernel).
template <int BN, int kPair, bool kGelu, bool kOutF32, bool kResid, bool kRemap>
__global__ void __launch_bounds__((gemm_cfg::Cfg<BN, kPair, gemm_cfg::epi_warps(kGelu)>::kThreads), 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                 GemmShape shape, GemmEpilogue ep) {
  using namespace gemm_cfg;
  constexpr int kEpiWarps = epi_warps(kGelu);
  constexpr int kColGroups = kEpiWarps / 4;  // warps per TMEM lane quarter; each owns BN / kColGroups columns
  using C = Cfg<BN, kPair, kEpiWarps>;
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
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (kPair == 2) ? ptx::cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int unit = blockIdx.x / kPair;       // CTA (kPair=1) or CTA-pair index
  const int num_units = gridDim.x / kPair;

  const int m_tiles = (shape.M + kTileM - 1) / kTileM;
  const int n_tiles = (shape.N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = (shape.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_a);
    ptx::prefetch_tmap(&tmap_w);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      ptx::mbar_init(&tmem_full_bar[a], 1);
      ptx::mbar_init(&tmem_empty_bar[a], kEpiWarps * kPair);  // one elected lane per epilogue warp (of both CTAs)
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    if (kPair == 2) ptx::tmem_alloc_pair<C::kTmemCols>(tmem_slot);
    else ptx::tmem_alloc<C::kTmemCols>(tmem_slot);
  }
  ptx::tc_fence_before();
  if (kPair == 2) ptx::cluster_sync();  // peer barriers initialised before any multicast commit / remote arrive
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // The producer and MMA loops are executed by WHOLE warps with one elected lane issuing: loop counters, smem
  // addresses and descriptors then stay warp-uniform (uniform registers), instead of being moved from vector to
  // uniform registers (R2UR) in front of every UTMALDG / UTCHMMA, which throttles a single-thread issue loop.
  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (every CTA)
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = unit; tile < num_tiles; tile += num_units) {
      const int m_blk = tile / n_tiles;
      const int n_blk = tile % n_tiles;
      const int a_row = m_blk * kTileM + static_cast<int>(cta_rank) * BM;
      const int w_row = n_blk * BN + static_cast<int>(cta_rank) * (BN / kPair);
      for (int kb = 0; kb < num_kb; ++kb) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * C::kStageBytes;
        uint8_t* sb = sa + C::kStageBytesA;
        if (ptx::elect_one()) {
          if (kPair == 2) {
            // one barrier (the leader's) tracks the bytes of both CTAs' halves
            if (leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * C::kStageBytes);
            ptx::tma_load_2d_pair(sa, &tmap_a, &full_bar[stage], kb * BK, a_row);
            ptx::tma_load_2d_pair(sb, &tmap_w, &full_bar[stage], kb * BK, w_row);
          } else {
            ptx::mbar_arrive_expect_tx(&full_bar[stage], C::kStageBytes);
            ptx::tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BK, a_row);
            ptx::tma_load_2d(sb, &tmap_w, &full_bar[stage], kb * BK, w_row);
          }
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1 && leader) {
    // ------------------------------------------------------------ UMMA issuer (leader CTA)
    constexpr uint32_t idesc = ptx::make_idesc_bf16(kTileM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int local = 0;
    for (int tile = unit; tile < num_tiles; tile += num_units, ++local) {
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      ptx::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);
      ptx::tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
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
          if (kPair == 2) ptx::umma_commit_pair(&empty_bar[stage], 0x3);
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
        if (kPair == 2) ptx::umma_commit_pair(&tmem_full_bar[acc], 0x3);
        else ptx::umma_commit(&tmem_full_bar[acc]);
      }
      __syncwarp();
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
    int local = 0;
    for (int tile = unit; tile < num_tiles; tile += num_units, ++local) {
      const int m_blk = tile / n_tiles;
      const int n_blk = tile % n_tiles;
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const int row_base = m_blk * kTileM + static_cast<int>(cta_rank) * BM + quarter * 32;
      ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t taddr0 =
          tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + col_grp * kGroupCols;
#pragma unroll 1
      for (int c = 0; c < kChunks; ++c) {
        uint32_t r[32];
        ptx::tmem_ld_x32(taddr0 + c * 32, r);
        ptx::tmem_ld_wait();
        if (c + 1 == kChunks) {
          // every TMEM read of this accumulator has landed in registers: hand it back to the MMA issuer
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (kPair == 2) ptx::mbar_arrive_cluster(&tmem_empty_bar[acc], 0);
            else ptx::mbar_arrive(&tmem_empty_bar[acc]);
          }
        }
        const int col0 = n_blk * BN + col_grp * kGroupCols + c * 32;
        if (col0 >= shape.N) continue;  // warp-uniform
        // registers (thread = row, 32 columns) -> slab
        float* srow = slab + lane * kSlabStride;
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<uint4*>(srow + j) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
        __syncwarp();
        const int col = col0 + 4 * tcol;
        const bool col_ok = col < shape.N;
        float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ep.bias != nullptr && col_ok) bias4 = *reinterpret_cast<const float4*>(ep.bias + col);
        // transposed pass: 8 rows (trow + 4 i) x 4 columns per thread.  With a residual all 8 rows form one batch so
        // that 8 independent 16-byte loads per thread are in flight (the out_proj / fc2 epilogues are bound by
        // HBM latency x outstanding bytes); otherwise two batches of 4 keep the instruction footprint small.
        constexpr int kRowBatch = kResid ? 8 : 4;
        const float* sl = slab + trow * kSlabStride + 4 * tcol;
#pragma unroll 1
        for (int i0 = 0; i0 < 8; i0 += kRowBatch) {
          float4 v[kRowBatch], q[kRowBatch];
          long orow[kRowBatch];
          bool ok[kRowBatch];
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
            if (kResid) {
              q[i] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (ok[i]) q[i] = *reinterpret_cast<const float4*>(ep.resid + resid_row * ep.ldr + col);
            }
          }
#pragma unroll
          for (int i = 0; i < kRowBatch; ++i) {
            float4 t = v[i];
            t.x += bias4.x, t.y += bias4.y, t.z += bias4.z, t.w += bias4.w;
            if (kGelu) t.x = gelu_erf(t.x), t.y = gelu_erf(t.y), t.z = gelu_erf(t.z), t.w = gelu_erf(t.w);
            if (kResid) t.x += q[i].x, t.y += q[i].y, t.z += q[i].z, t.w += q[i].w;
            if (ok[i]) {
              if (kOutF32) {
                *reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + orow[i] * ep.ldo + col) = t;
              } else {
                __nv_bfloat162 lo = __floats2bfloat162_rn(t.x, t.y);
                __nv_bfloat162 hi = __floats2bfloat162_rn(t.z, t.w);
                uint2 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&lo);
                pk.y = *reinterpret_cast<uint32_t*>(&hi);
                *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(ep.out) + orow[i] * ep.ldo + col) = pk;
              }
            }
          }
        }
        __syncwarp();  // slab is rewritten by the next chunk
      }
    }
  }

  __syncwarp();
  ptx::tc_fence_before();
  if (kPair == 2) ptx::cluster_sync();  // the peer may still be signalling this CTA's barriers / reading its smem
  else __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    if (kPair == 2) ptx::tmem_dealloc_pair<C::kTmemCols>(tmem_base);
    else ptx::tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

}  // namespace vitb200
