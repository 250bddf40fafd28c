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
// L vector-matrix products of N x N instead of L matrix-matrix products.  One CTA per image.
// maps: [L][B, N, ld] fp32 (layer stride given in floats).
__global__ void __launch_bounds__(256)
rollout_cls_kernel(const float* __restrict__ maps, long layer_stride, int L, int N, int ld, float* __restrict__ out /*[B, N-1]*/) {
  extern __shared__ float sm[];
  float* r = sm;            // [ld] current row vector (scaled by 1/rowsum below)
  float* rn = sm + ld;      // [ld] next
  float* rs = sm + 2 * ld;  // [ld] 1 / rowsum of (0.5 A + 0.5 I)
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  for (int l = L - 1; l >= 0; --l) {
    const float* A = maps + l * layer_stride + static_cast<long>(b) * N * ld;
    // row sums of 0.5 A + 0.5 I
    for (int k = warp; k < N; k += nwarps) {
      float s = 0.f;
      for (int j = lane; j < N; j += 32) s += 0.5f * A[static_cast<long>(k) * ld + j] + (j == k ? 0.5f : 0.f);
      s = warp_sum(s);
      if (lane == 0) rs[k] = 1.0f / s;
    }
    __syncthreads();
    if (l == L - 1) {
      for (int j = tid; j < N; j += blockDim.x) rn[j] = (0.5f * A[j] + (j == 0 ? 0.5f : 0.f)) * rs[0];
    } else {
      for (int j = tid; j < N; j += blockDim.x) {
        float acc = 0.f;
        for (int k = 0; k < N; ++k) {
          const float a = 0.5f * A[static_cast<long>(k) * ld + j] + (j == k ? 0.5f : 0.f);
          acc = fmaf(r[k] * rs[k], a, acc);
        }
        rn[j] = acc;
      }
    }
    __syncthreads();
    for (int j = tid; j < N; j += blockDim.x) r[j] = rn[j];
    __syncthreads();
  }
  for (int j = tid + 1; j < N; j += blockDim.x) out[static_cast<long>(b) * (N - 1) + j - 1] = r[j];
}

}  // namespace vitb200
