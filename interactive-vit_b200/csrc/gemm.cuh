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
// Roles (256 threads, 1 CTA / SM, persistent over output tiles):
//   warp 0 lane 0 : TMA producer  (A tile 128x64, W tile BNx64, SWIZZLE_128B, kStages-deep mbarrier ring)
//   warp 1 lane 0 : UMMA issuer   (tcgen05.mma cta_group::1 kind::f16, M=128, N=BN, K=16 x4 per stage)
//   warp 2        : TMEM allocator (2 accumulator buffers of BN fp32 columns -> MMA of tile i+1 overlaps
//                   the epilogue of tile i)
//   warps 4..7    : epilogue (tcgen05.ld 32x32b -> +bias [-> GELU] [+ fp32 residual / pos-embedding]
//                   -> bf16 or fp32 global stores)
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
constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = one 128-byte swizzle span
constexpr int kThreads = 256;
constexpr int kEpiWarp0 = 4;
template <int BN>
struct Cfg {
  static constexpr int kStageBytesA = BM * BK * 2;
  static constexpr int kStageBytesB = BN * BK * 2;
  static constexpr int kStageBytes = kStageBytesA + kStageBytesB;
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kTmemCols = 2 * BN;  // two accumulator buffers
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};
}  // namespace gemm_cfg

// Exact-erf GELU (torch.nn.GELU default, vision_transformer.py:40-47 MLPBlock).
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

template <int BN, bool kGelu, bool kOutF32>
__global__ void __launch_bounds__(gemm_cfg::kThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w,
                 GemmShape shape, GemmEpilogue ep) {
  using namespace gemm_cfg;
  using C = Cfg<BN>;
  constexpr int kStages = C::kStages;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * C::kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int m_tiles = (shape.M + BM - 1) / BM;
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
      ptx::mbar_init(&tmem_empty_bar[a], 128);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc<C::kTmemCols>(tmem_slot);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------ TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / n_tiles;
      const int n_blk = tile % n_tiles;
      for (int kb = 0; kb < num_kb; ++kb) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem + stage * C::kStageBytes;
        uint8_t* sb = sa + C::kStageBytesA;
        ptx::mbar_arrive_expect_tx(&full_bar[stage], C::kStageBytes);
        ptx::tma_load_2d(sa, &tmap_a, &full_bar[stage], kb * BK, m_blk * BM);
        ptx::tma_load_2d(sb, &tmap_w, &full_bar[stage], kb * BK, n_blk * BN);
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------------------ UMMA issuer
    constexpr uint32_t idesc = ptx::make_idesc_bf16(BM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int local = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
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
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // +32 bytes per K=16 step inside the 128-B swizzle span (descriptor address is in 16-B units)
          ptx::umma_bf16_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        ptx::umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs retire
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
      ptx::umma_commit(&tmem_full_bar[acc]);  // accumulator complete -> epilogue
    }
  } else if (warp >= kEpiWarp0) {
    // ------------------------------------------------------------ epilogue
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    int local = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++local) {
      const int m_blk = tile / n_tiles;
      const int n_blk = tile % n_tiles;
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      const int row = m_blk * BM + quarter * 32 + lane;
      const bool row_ok = row < shape.M;
      long out_row = row, resid_row = row;
      if (ep.group_rows > 0) {
        const int g = row / ep.group_rows;
        const int i = row - g * ep.group_rows;
        out_row = static_cast<long>(g) * ep.out_group_stride + ep.out_row_offset + i;
        resid_row = ep.resid_broadcast ? (ep.resid_row_offset + i) : out_row;
      }
      ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
      ptx::tc_fence_after();
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN;
#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int col0 = n_blk * BN + c * 32;
        uint32_t r[32];
        ptx::tmem_ld_x32(taddr0 + c * 32, r);
        ptx::tmem_ld_wait();
        if (col0 >= shape.N) continue;  // warp-uniform
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (ep.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (col0 + j < shape.N) {
              const float4 b = *reinterpret_cast<const float4*>(ep.bias + col0 + j);
              v[j] += b.x, v[j + 1] += b.y, v[j + 2] += b.z, v[j + 3] += b.w;
            }
          }
        }
        if (kGelu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
        }
        if (row_ok) {
          if (ep.resid != nullptr) {
            const float* rp = ep.resid + resid_row * ep.ldr + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (col0 + j < shape.N) {
                const float4 b = *reinterpret_cast<const float4*>(rp + j);
                v[j] += b.x, v[j + 1] += b.y, v[j + 2] += b.z, v[j + 3] += b.w;
              }
            }
          }
          if (kOutF32) {
            float* op = reinterpret_cast<float*>(ep.out) + out_row * ep.ldo + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (col0 + j < shape.N) *reinterpret_cast<float4*>(op + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
          } else {
            __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(ep.out) + out_row * ep.ldo + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              if (col0 + j < shape.N) {
                uint4 pk;
                __nv_bfloat162 t0 = __floats2bfloat162_rn(v[j], v[j + 1]);
                __nv_bfloat162 t1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
                __nv_bfloat162 t2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]);
                __nv_bfloat162 t3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
                pk.x = *reinterpret_cast<uint32_t*>(&t0);
                pk.y = *reinterpret_cast<uint32_t*>(&t1);
                pk.z = *reinterpret_cast<uint32_t*>(&t2);
                pk.w = *reinterpret_cast<uint32_t*>(&t3);
                *reinterpret_cast<uint4*>(op + j) = pk;
              }
            }
          }
        }
      }
      ptx::tc_fence_before();
      ptx::mbar_arrive(&tmem_empty_bar[acc]);
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<C::kTmemCols>(tmem_base);
  }
}

}  // namespace vitb200
