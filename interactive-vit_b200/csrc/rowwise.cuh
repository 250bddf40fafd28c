// HBM-bound kernels of the ViT forward path: LayerNorm, patch gather (im2col fused with the fp32 -> bf16
// ingest), class-token rows, fp32 -> bf16 weight packing, attention rollout.  All accesses are 128-bit and
// coalesced; reductions are warp-shuffle based.  Reference arithmetic:
//   LayerNorm(eps = 1e-6)               torchvision vision_transformer.py:96,105,134,175
//   _process_input / class token / pos  vision_transformer.py:268-287, 295-296, Encoder.forward 154-157
//   rollout                             not in the reference (north_star feature); defined by oracle/vit_oracle.py
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>

namespace vitb200 {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

// ---------------------------------------------------------------------------------------------------
// LayerNorm over the last dim of an fp32 row, bf16 output (the A operand of the following GEMM).
// One warp per row; the row lives in registers between the mean and the variance pass (two-pass, like
// ATen's CPU kernel).  kVec = d / 128 float4 loads per lane.  in_row_stride lets the final LayerNorm read
// only the class-token rows (stride N*d) while writing a dense [B, d] matrix.
template <int kVec>
__global__ void __launch_bounds__(256)
layernorm_f32_bf16_kernel(const float* __restrict__ x, long in_row_stride, const float* __restrict__ gamma,
                          const float* __restrict__ beta, __nv_bfloat16* __restrict__ y, int rows, float eps) {
  constexpr int d = kVec * 128;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float4* xr = reinterpret_cast<const float4*>(x + static_cast<long>(warp) * in_row_stride);
  float4 v[kVec];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    v[i] = xr[lane + 32 * i];
    s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  const float mean = warp_sum(s) * (1.0f / d);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, e = v[i].w - mean;
    q += (a * a + b * b) + (c * c + e * e);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.0f / d) + eps);
  uint2* yr = reinterpret_cast<uint2*>(y + static_cast<long>(warp) * d);
  const float4* g4 = reinterpret_cast<const float4*>(gamma);
  const float4* b4 = reinterpret_cast<const float4*>(beta);
#pragma unroll
  for (int i = 0; i < kVec; ++i) {
    const float4 g = g4[lane + 32 * i], bb = b4[lane + 32 * i];
    uint2 o;
    o.x = pack_bf16x2((v[i].x - mean) * rstd * g.x + bb.x, (v[i].y - mean) * rstd * g.y + bb.y);
    o.y = pack_bf16x2((v[i].z - mean) * rstd * g.z + bb.z, (v[i].w - mean) * rstd * g.w + bb.w);
    yr[lane + 32 * i] = o;
  }
}

// ---------------------------------------------------------------------------------------------------
// Patch gather: images fp32 [B, 3, S, S] -> patch matrix bf16 [B * n, 3 * p * p] with
// column = c * p * p + ky * p + kx (the flattening of conv_proj.weight [d, 3, p, p]), row = b * n + py * np + px.
// One thread moves 8 consecutive kx: two float4 reads, one 16-byte write.
__global__ void __launch_bounds__(256)
patchify_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int S, int p) {
  const int np = S / p;
  const int kx8 = p / 8;                                 // 16-byte groups per patch row
  const long total = static_cast<long>(B) * 3 * S * np * kx8;  // one item per (b, c, y, px, g)
  const long idx = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  // decode so that consecutive threads read consecutive image addresses: (b, c, y, px, g), g fastest
  long t = idx;
  const int g = t % kx8; t /= kx8;
  const int px = t % np; t /= np;
  const int y = t % S; t /= S;
  const int c = t % 3; t /= 3;
  const int b = static_cast<int>(t);
  const int py = y / p, ky = y - py * p;
  const float4* src = reinterpret_cast<const float4*>(img + ((static_cast<long>(b) * 3 + c) * S + y) * S + px * p + g * 8);
  const float4 a = src[0], bb = src[1];
  uint4 o;
  o.x = pack_bf16x2(a.x, a.y);
  o.y = pack_bf16x2(a.z, a.w);
  o.z = pack_bf16x2(bb.x, bb.y);
  o.w = pack_bf16x2(bb.z, bb.w);
  const long row = static_cast<long>(b) * np * np + py * np + px;
  const long col = static_cast<long>(c) * p * p + ky * p + g * 8;
  *reinterpret_cast<uint4*>(out + row * (3L * p * p) + col) = o;
}

// Class-token rows of the token stream: x[b, 0, :] = class_token + pos_embedding[0]  (fp32).
__global__ void __launch_bounds__(256)
cls_rows_kernel(const float* __restrict__ cls, const float* __restrict__ pos, float* __restrict__ x, int B, int N,
                int d) {
  const long idx = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const int d4 = d / 4;
  if (idx >= static_cast<long>(B) * d4) return;
  const int b = idx / d4, j = idx % d4;
  const float4 c = reinterpret_cast<const float4*>(cls)[j];
  const float4 q = reinterpret_cast<const float4*>(pos)[j];
  reinterpret_cast<float4*>(x + static_cast<long>(b) * N * d)[j] = make_float4(c.x + q.x, c.y + q.y, c.z + q.z, c.w + q.w);
}

// fp32 -> bf16 (weights at load time; also activations arriving from the wire).
__global__ void __launch_bounds__(256)
f32_to_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, long n) {
  const long i = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 a = *reinterpret_cast<const float4*>(in + i);
    uint2 o;
    o.x = pack_bf16x2(a.x, a.y);
    o.y = pack_bf16x2(a.z, a.w);
    *reinterpret_cast<uint2*>(out + i) = o;
  } else {
    for (long k = i; k < n; ++k) out[k] = __float2bfloat16_rn(in[k]);
  }
}

// ---------------------------------------------------------------------------------------------------
// Attention rollout, class-token row only.  For layer l with head-averaged map Abar_l [N, N]:
//     Ahat_l = rownorm(0.5 * Abar_l + 0.5 * I),      R = Ahat_L ... Ahat_1,      out = R[0, 1:]
// Only row 0 of R is needed, so r <- r * Ahat_l is evaluated right-to-left over layers L, L-1, ..., 1:
// L vector-matrix products of N x N instead of L matrix-matrix products.  With w_k = r_k / rowsum_k:
//     (r * Ahat)_j = 0.5 * sum_k w_k Abar[k, j] + 0.5 * w_j
// One CTA (1024 threads) per image: 32 warps take the row sums (coalesced 128-B row reads), then 4 thread
// groups split the k range of the vector-matrix product (coalesced across j) and are reduced through smem.
// maps: [L][B, N, ld] fp32 (layer stride given in floats); N <= 256.
constexpr int kRolloutThreads = 1024;
__global__ void __launch_bounds__(kRolloutThreads)
rollout_cls_kernel(const float* __restrict__ maps, long layer_stride, int L, int N, int ld, float* __restrict__ out /*[B, N-1]*/) {
  __shared__ float r[256];        // current row vector
  __shared__ float w[256];        // r_k / rowsum_k
  __shared__ float part[4][256];  // partial products per k group
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int j = tid & 255, g = tid >> 8;
  const int kper = (N + 3) / 4;
  for (int l = L - 1; l >= 0; --l) {
    const float* A = maps + l * layer_stride + static_cast<long>(b) * N * ld;
    // w_k = r_k / (0.5 * rowsum_k + 0.5); for the top layer r = e_0
    for (int k = warp; k < N; k += 32) {
      float s = 0.f;
#pragma unroll 8
      for (int c = lane; c < N; c += 32) s += A[static_cast<long>(k) * ld + c];
      s = warp_sum(s);
      if (lane == 0) {
        const float rk = (l == L - 1) ? (k == 0 ? 1.0f : 0.0f) : r[k];
        w[k] = rk / (0.5f * s + 0.5f);
      }
    }
    __syncthreads();
    float acc = 0.f;
    if (j < N) {
      if (l == L - 1) {
        if (g == 0) acc = w[0] * A[j];  // only k = 0 contributes
      } else {
        const int k0 = g * kper, k1 = min(N, k0 + kper);
#pragma unroll 8
        for (int k = k0; k < k1; ++k) acc = fmaf(w[k], A[static_cast<long>(k) * ld + j], acc);
      }
    }
    part[g][j] = acc;
    __syncthreads();
    if (g == 0 && j < N) r[j] = 0.5f * ((part[0][j] + part[1][j]) + (part[2][j] + part[3][j])) + 0.5f * w[j];
    __syncthreads();
  }
  if (g == 0 && j >= 1 && j < N) out[static_cast<long>(b) * (N - 1) + j - 1] = r[j];
}

}  // namespace vitb200
