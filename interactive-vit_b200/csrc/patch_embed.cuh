// Patch embedding with the im2col done by TMA: torchvision's conv_proj (a p x p stride-p convolution,
// vision_transformer.py:268-287 _process_input) + positional embedding, straight from the fp32 images.
//
//     x[b, 1 + py*g + px, n] = sum_{c,ky,kx} img[b, c, py*p + ky, px*p + kx] * W[n, c, ky, kx] + bias[n] + pos[1 + py*g + px, n]
//
// The image is viewed as a 5-D tensor (kx, px, ky, py, b*3 + c) with the strides of [B, 3, S, S]; ONE tiled TMA load
// with the box (16, g, 1, pr, 1) lands the A tile of a k-block -- rows = pr patch rows x g patches, columns = the 16
// pixels of one kernel row -- in shared memory as K-major SWIZZLE_128B rows a tcgen05 operand descriptor reads.  (With the
// 128-byte swizzle TMA gives every innermost box row its own 128-byte line, whatever the box's inner extent -- measured,
// `tools/micro/tma5d_probe.cu`: a (16, 2, ...) box does NOT put two kernel rows side by side -- so a k-block is 16 fp32 =
// the first half of each line, and the weight tile is loaded with a 16-wide box the same way.)  No patch matrix in HBM (round 1: a patchify kernel wrote 77 MB of bf16 patches
// that the GEMM read back), no conversion pass: the MMAs are kind::tf32 on the fp32 pixels and fp32 weights (10
// mantissa bits against bf16's 7; half the bf16 rate, which a 59-GFLOP GEMM can afford).
//
// Persistent CTAs (one per SM) over units = (image, group of pr patch rows, 256-column tile); UMMA M = 128 (the
// pr * g <= 128 patch rows of the unit; the rest of the 128 rows is stale shared memory whose accumulator rows are never
// read), N = 256, K = 8 per instruction, 48 k-blocks of 16 (channel c, kernel row ky).  Two TMEM accumulators: the
// epilogue of unit i overlaps the MMAs of unit i + 1.  Epilogue (4 warps, thread = patch row, TMEM-native layout):
// + bias + positional embedding -> fp32 token stream, its bf16 copy and the per-row LayerNorm partial sums that the first
// qkv GEMM consumes (one slot per slot_width columns, a 128-column slot formed as the sum of its 64-column halves: see
// gemm.cuh), staged in swizzled shared memory and TMA-stored; a warp whose 32 rows are not all valid stores directly.
#pragma once
#include <cuda.h>
#include "ptx.cuh"
#include "gemm.cuh"

namespace vitb200 {

struct PatchEmbedParams {
  int B, S, g, N, d;        // images, image side, patches per side, tokens per image (g*g + 1), width
  int pr;                   // patch rows per unit (pr * g <= 128)
  int groups;               // units per image along py: ceil(g / pr)
  const float* bias;        // [d]
  const float* pos;         // [N, d]
  float* x;                 // [B*N, d]
  __nv_bfloat16* xb;        // [B*N, d]
  float2* row_stats;        // [B*N, d / slot_width]
  int slot_width;           // 64 or 128
};

namespace patch_cfg {
constexpr int BM = 128, BN = 256, BK = 16;          // BK fp32 = 64 bytes = the used half of every 128-byte smem row
constexpr int kStages = 3;
constexpr int kBytesA = BM * 128, kBytesW = BN * 128, kStageBytes = kBytesA + kBytesW;
constexpr int kTileF32 = 32 * 128, kTileBf16 = 32 * 64;
constexpr int kWarpBytes = 2 * (kTileF32 + kTileBf16);                    // two result tile pairs per epilogue warp
constexpr int kThreads = 256;
constexpr int kSmemBytes = kStages * kStageBytes + 4 * kWarpBytes + 256;
static_assert(kSmemBytes <= 227 * 1024, "patch embedding: shared memory budget");
}  // namespace patch_cfg

__device__ __forceinline__ void umma_tf32_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* smem_dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(ptx::smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(ptx::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3), "r"(c4)
      : "memory");
}

__global__ void __launch_bounds__(patch_cfg::kThreads, 1)
patch_embed_kernel(const __grid_constant__ CUtensorMap tmap_img,   // fp32 5-D (kx 16, px g, ky 16, py g, B*3), box (16, g, 1, pr, 1)
                   const __grid_constant__ CUtensorMap tmap_w,     // fp32 [d, 768], box 16 x 256
                   const __grid_constant__ CUtensorMap tmap_x,     // fp32 [B][N][d], box 32 x 32 x 1, SWIZZLE_128B
                   const __grid_constant__ CUtensorMap tmap_xb,    // bf16 [B][N][d], box 32 x 32 x 1, SWIZZLE_64B
                   PatchEmbedParams p) {
  using namespace patch_cfg;
  extern __shared__ __align__(1024) uint8_t smem[];
  if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* epi_smem = smem + kStages * kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(epi_smem + 4 * kWarpBytes);
  uint64_t* full_bar = bars;                       // [kStages]
  uint64_t* empty_bar = bars + kStages;            // [kStages]
  uint64_t* tmem_full_bar = bars + 2 * kStages;    // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = p.d / BN;
  const int num_units = p.B * p.groups * n_tiles;
  constexpr int num_kb = 3 * 16;                   // (channel, kernel row)
  const uint32_t a_tx = static_cast<uint32_t>(p.pr) * p.g * 64u;    // bytes one image box delivers (OOB patch rows are zero-filled)
  constexpr uint32_t w_tx = BN * 64u;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmap_img);
    ptx::prefetch_tmap(&tmap_w);
    ptx::prefetch_tmap(&tmap_x);
    ptx::prefetch_tmap(&tmap_xb);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < kStages; ++s) ptx::mbar_init(&full_bar[s], 1), ptx::mbar_init(&empty_bar[s], 1);
    for (int a = 0; a < 2; ++a) ptx::mbar_init(&tmem_full_bar[a], 1), ptx::mbar_init(&tmem_empty_bar[a], 4);
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int unit, int& b, int& grp, int& n_blk) {
    n_blk = unit % n_tiles;
    const int t = unit / n_tiles;
    grp = t % p.groups, b = t / p.groups;
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
      int b, grp, n_blk;
      decode(unit, b, grp, n_blk);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int c = kb >> 4, ky = kb & 15;
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * kStageBytes;
        if (ptx::elect_one()) {
          ptx::mbar_arrive_expect_tx(&full_bar[stage], a_tx + w_tx);
          tma_load_5d(sa, &tmap_img, &full_bar[stage], 0, 0, ky, grp * p.pr, b * 3 + c);
          ptx::tma_load_2d(sa + kBytesA, &tmap_w, &full_bar[stage], kb * BK, n_blk * BN);
        }
        __syncwarp();
        if (++stage == kStages) stage = 0, phase ^= 1;
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ UMMA issuer (kind::tf32: a_format = b_format = 2)
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((BN >> 3) << 17) | ((BM >> 4) << 24);
    int stage = 0, local = 0;
    uint32_t phase = 0;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++local) {
      const int acc = local & 1;
      ptx::mbar_wait(&tmem_empty_bar[acc], ((local >> 1) & 1) ^ 1);
      ptx::tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        const uint32_t sa = ptx::smem_u32(smem + stage * kStageBytes);
        const uint64_t da = ptx::make_smem_desc_sw128(sa, 16, 1024);
        const uint64_t db = ptx::make_smem_desc_sw128(sa + kBytesA, 16, 1024);
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / 8; ++k)   // K = 8 fp32 = 32 bytes per instruction: +2 descriptor units
            umma_tf32_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          ptx::umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == kStages) stage = 0, phase ^= 1;
      }
      if (ptx::elect_one()) ptx::umma_commit(&tmem_full_bar[acc]);
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue: warp q = TMEM lane quarter q, tile rows 32 q ..
    const int q = warp - 4;
    uint8_t* my = epi_smem + q * kWarpBytes;
    const uint32_t s_f32 = ptx::smem_u32(my), s_bf16 = s_f32 + 2 * kTileF32;
    const int sw = lane & 7, sw64 = (lane >> 1) & 3;
    const uint32_t row_f32 = static_cast<uint32_t>(lane) * 128u, row_bf16 = static_cast<uint32_t>(lane) * 64u;
    const int chunks_per_slot = p.slot_width >> 5;
    const int slots = p.d / p.slot_width;
    int local = 0, n_store = 0;
    for (int unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++local) {
      int b, grp, n_blk;
      decode(unit, b, grp, n_blk);
      const int acc = local & 1;
      const int py0 = grp * p.pr;
      const int valid = min(p.pr, p.g - py0) * p.g;          // patch rows of this unit that exist
      const int vq = min(32, max(0, valid - 32 * q));        // ... among this warp's 32
      const int tok0 = 1 + py0 * p.g + 32 * q;               // token index (inside the image) of this warp's first row
      const bool row_ok = lane < vq;
      const long grow = static_cast<long>(b) * p.N + tok0 + lane;   // global token row of this thread
      ptx::mbar_wait(&tmem_full_bar[acc], (local >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * BN;
      uint32_t r[2][32];
      ptx::tmem_ld_x32(taddr0, r[0]);
      float st1 = 0.f, st2 = 0.f, stb1 = 0.f, stb2 = 0.f;
#pragma unroll
      for (int c = 0; c < BN / 32; ++c) {
        const int col0 = n_blk * BN + c * 32;
        ptx::tmem_ld_wait();
        if (c + 1 < BN / 32) {
          ptx::tmem_ld_x32(taddr0 + (c + 1) * 32, r[(c + 1) & 1]);
        } else {
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&tmem_empty_bar[acc]);
        }
        const uint32_t (&a)[32] = r[c & 1];
        const bool bulk = vq == 32;                           // warp-uniform: a full tile leaves with two TMA stores
        const int buf = n_store & 1;
        if (bulk) {
          // the stores that used this tile pair two chunks ago have read it (only lane 0 has bulk groups)
          if (lane == 0) ptx::tma_store_wait_read<1>();
          __syncwarp();
        }
        const uint32_t t_f32 = s_f32 + buf * kTileF32 + row_f32, t_bf16 = s_bf16 + buf * kTileBf16 + row_bf16;
        const float* pos_row = p.pos + static_cast<long>(tok0 + lane) * p.d + col0;
        uint32_t xbw[16];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0) + j);
          float4 ps = make_float4(0.f, 0.f, 0.f, 0.f);
          if (row_ok) ps = __ldg(reinterpret_cast<const float4*>(pos_row) + j);
          float4 t;
          t.x = (__uint_as_float(a[4 * j]) + b4.x) + ps.x, t.y = (__uint_as_float(a[4 * j + 1]) + b4.y) + ps.y;
          t.z = (__uint_as_float(a[4 * j + 2]) + b4.z) + ps.z, t.w = (__uint_as_float(a[4 * j + 3]) + b4.w) + ps.w;
          s1 += (t.x + t.y) + (t.z + t.w);
          s2 += (t.x * t.x + t.y * t.y) + (t.z * t.z + t.w * t.w);
          xbw[2 * j] = pack2_bf16(t.x, t.y), xbw[2 * j + 1] = pack2_bf16(t.z, t.w);
          if (bulk) {
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(t_f32 + (static_cast<uint32_t>(j ^ sw) << 4)), "f"(t.x),
                         "f"(t.y), "f"(t.z), "f"(t.w)
                         : "memory");
          } else if (row_ok) {
            reinterpret_cast<float4*>(p.x + grow * p.d + col0)[j] = t;
          }
        }
        if (bulk) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(t_bf16 + (static_cast<uint32_t>(j ^ sw64) << 4)),
                         "r"(xbw[4 * j]), "r"(xbw[4 * j + 1]), "r"(xbw[4 * j + 2]), "r"(xbw[4 * j + 3])
                         : "memory");
          ptx::fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            ptx::tma_store_3d(&tmap_x, my + buf * kTileF32, col0, tok0, b);
            ptx::tma_store_3d(&tmap_xb, my + 2 * kTileF32 + buf * kTileBf16, col0, tok0, b);
            ptx::tma_store_commit();
          }
          ++n_store;
        } else if (row_ok) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            reinterpret_cast<uint4*>(p.xb + grow * p.d + col0)[j] = make_uint4(xbw[4 * j], xbw[4 * j + 1], xbw[4 * j + 2], xbw[4 * j + 3]);
        }
        // LayerNorm partial sums: chunk sums in ascending order inside a 64-column half, halves added last
        if (p.slot_width == 128 && (c & 3) >= 2) stb1 += s1, stb2 += s2;
        else st1 += s1, st2 += s2;
        if (((c + 1) % chunks_per_slot) == 0) {
          if (row_ok) p.row_stats[grow * slots + (col0 + 32 - p.slot_width) / p.slot_width] = make_float2(st1 + stb1, st2 + stb2);
          st1 = st2 = stb1 = stb2 = 0.f;
        }
      }
    }
    if (lane == 0) ptx::tma_store_wait_read<0>();
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace vitb200
