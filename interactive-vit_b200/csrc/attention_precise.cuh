// Multi-head self-attention of the fp32x3 ("precise") mode for head dims other than 64 (ViT-H: 80): plain fp32
// arithmetic on the CUDA cores.  The precise mode carries every GEMM operand as hi + lo bf16 halves; for head dim 64 the
// key-blocked tensor-core kernels take split operands (attention_long.cuh: [hi box | lo box] per 32 KB slot), for wider
// heads a split operand tile would need two boxes per half -- 64 KB per operand, more shared memory than an SM has for
// Q, two K stages, two V stages and P.  The precise mode is the verification mode (north_star's <= 1e-3 bound), not the
// benchmark mode, so this kernel spends FFMAs instead: values are reconstructed as float(hi) + float(lo) (17
// significant bits, like the split GEMMs' operands) and everything downstream is fp32.
//
// Arithmetic: torch.nn.functional.multi_head_attention_forward, weights branch (torch/nn/functional.py:6630-6659):
// q scaled by 1 / sqrt(D), softmax(q k^T) v, optional head mean -- the same contract as attention.cuh.
//
// One CTA (256 threads) per (image, 16-query tile) looping over the heads, so that the head average accumulates in
// shared memory in a fixed order (no atomics: bit-reproducible).  Per head: S[16][N] into shared memory (keys in
// chunks of 64 through a staging tile), row softmax by one warp per two rows (expf, true division), maps written
// from the normalised probabilities, O = P V over the same key chunks.
#pragma once
#include <cuda_bf16.h>
#include "attention_long.cuh"

namespace vitb200 {

namespace attn_precise_cfg {
constexpr int kThreads = 256;
constexpr int QT = 16;     // query rows per CTA
constexpr int KC = 64;     // keys per staged chunk
__host__ __device__ constexpr int n_pad(int N) { return (N + 3) / 4 * 4 + 4; }
// q[QT][D] + kv[KC][D + 1] + s[QT][n_pad] + avg[QT][n_pad]
__host__ __device__ constexpr size_t smem_bytes(int N, int D) {
  return sizeof(float) * (static_cast<size_t>(QT) * D + static_cast<size_t>(KC) * (D + 1) + 2 * static_cast<size_t>(QT) * n_pad(N));
}
}  // namespace attn_precise_cfg

__global__ void __launch_bounds__(attn_precise_cfg::kThreads)
attention_precise_kernel(const __nv_bfloat16* __restrict__ qkv_hi, const __nv_bfloat16* __restrict__ qkv_lo, AttnLongParams p) {
  using namespace attn_precise_cfg;
  extern __shared__ __align__(16) float sm_f[];
  const int N = p.N, D = p.D, H = p.H, d = p.d, np = n_pad(N);
  float* q_s = sm_f;                    // [QT][D], already scaled by 1 / sqrt(D)
  float* kv_s = q_s + QT * D;           // [KC][D + 1]
  float* s_s = kv_s + KC * (D + 1);     // [QT][np] scores, then probabilities
  float* avg_s = s_s + QT * np;         // [QT][np] sum of the probabilities over the heads
  const int q_tiles = (N + QT - 1) / QT;
  const int b = blockIdx.x / q_tiles, q0 = (blockIdx.x - b * q_tiles) * QT;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long row_base = static_cast<long>(b) * N;
  const float scale = 1.0f / sqrtf(static_cast<float>(D));
  const bool want_avg = p.avg_map != nullptr;
  auto val = [&](long row, int col) {   // fp32 value of the split activation
    const long i = row * (3L * d) + col;
    return __bfloat162float(qkv_hi[i]) + __bfloat162float(qkv_lo[i]);
  };
  if (want_avg)
    for (int i = tid; i < QT * np; i += kThreads) avg_s[i] = 0.f;

  for (int h = 0; h < H; ++h) {
    __syncthreads();   // the previous head's P V has read s_s and kv_s
    for (int i = tid; i < QT * D; i += kThreads) {
      const int r = i / D, c = i - r * D;
      q_s[i] = (q0 + r < N) ? val(row_base + q0 + r, h * D + c) * scale : 0.f;
    }
    // ---- S = (q / sqrt(D)) K^T
    for (int kc = 0; kc < N; kc += KC) {
      __syncthreads();
      for (int i = tid; i < KC * D; i += kThreads) {
        const int j = i / D, c = i - j * D;
        kv_s[j * (D + 1) + c] = (kc + j < N) ? val(row_base + kc + j, d + h * D + c) : 0.f;
      }
      __syncthreads();
      for (int i = tid; i < QT * KC; i += kThreads) {
        const int r = i / KC, j = i - r * KC;   // consecutive threads: consecutive keys (kv rows D + 1 apart: no bank conflict)
        if (kc + j < N) {
          const float* qr = q_s + r * D;
          const float* kr = kv_s + j * (D + 1);
          float acc = 0.f;
          for (int c = 0; c < D; ++c) acc = fmaf(qr[c], kr[c], acc);
          s_s[r * np + kc + j] = acc;
        }
      }
    }
    __syncthreads();
    // ---- row softmax: warp w owns rows w and w + 8
    for (int r = warp; r < QT; r += kThreads / 32) {
      float* sr = s_s + r * np;
      float mx = -INFINITY;
      for (int j = lane; j < N; j += 32) mx = fmaxf(mx, sr[j]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float sum = 0.f;
      for (int j = lane; j < N; j += 32) {
        const float e = expf(sr[j] - mx);
        sr[j] = e;
        sum += e;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const int qrow = q0 + r;
      float* hm = (p.head_map != nullptr && qrow < N)
                      ? p.head_map + ((static_cast<size_t>(b) * H + h) * N + qrow) * p.ldmap : nullptr;
      float* cm = (p.cls_map != nullptr && qrow == 0) ? p.cls_map + (static_cast<size_t>(b) * H + h) * N : nullptr;
      for (int j = lane; j < N; j += 32) {
        const float pr = sr[j] / sum;
        sr[j] = pr;
        if (want_avg) avg_s[r * np + j] += pr;
        if (hm) hm[j] = pr;
        if (cm) cm[j] = pr;
      }
    }
    // ---- O = P V: thread (r, c) pairs, keys in the same chunks
    constexpr int kMaxOut = (QT * 128 + kThreads - 1) / kThreads;   // D <= 128
    float o[kMaxOut];
#pragma unroll
    for (int k = 0; k < kMaxOut; ++k) o[k] = 0.f;
    for (int kc = 0; kc < N; kc += KC) {
      __syncthreads();   // probabilities complete (first chunk) / previous chunk consumed
      for (int i = tid; i < KC * D; i += kThreads) {
        const int j = i / D, c = i - j * D;
        kv_s[j * (D + 1) + c] = (kc + j < N) ? val(row_base + kc + j, 2 * d + h * D + c) : 0.f;
      }
      __syncthreads();
      const int nk = min(KC, N - kc);
#pragma unroll
      for (int k = 0; k < kMaxOut; ++k) {
        const int i = tid + k * kThreads;
        if (i < QT * D) {
          const int r = i / D, c = i - r * D;
          const float* pr = s_s + r * np + kc;
          float acc = o[k];
          for (int j = 0; j < nk; ++j) acc = fmaf(pr[j], kv_s[j * (D + 1) + c], acc);
          o[k] = acc;
        }
      }
    }
#pragma unroll
    for (int k = 0; k < kMaxOut; ++k) {
      const int i = tid + k * kThreads;
      if (i < QT * D) {
        const int r = i / D, c = i - r * D;
        if (q0 + r < N) {
          const long oi = (row_base + q0 + r) * d + h * D + c;
          const __nv_bfloat16 hi = __float2bfloat16_rn(o[k]);
          p.ctx[oi] = hi;
          if (p.ctx_lo != nullptr) p.ctx_lo[oi] = __float2bfloat16_rn(o[k] - __bfloat162float(hi));
        }
      }
    }
  }
  if (want_avg) {
    __syncthreads();
    const float inv_h = 1.0f / static_cast<float>(H);
    for (int i = tid; i < QT * N; i += kThreads) {
      const int r = i / N, j = i - r * N;
      if (q0 + r < N) p.avg_map[(static_cast<size_t>(b) * N + q0 + r) * p.ldmap + j] = avg_s[r * np + j] * inv_h;
    }
  }
}

}  // namespace vitb200
