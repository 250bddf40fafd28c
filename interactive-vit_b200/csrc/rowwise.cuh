// HBM-bound kernels of the ViT forward path: LayerNorm, patch gather (im2col fused with the fp32 -> bf16
// ingest), class-token rows, fp32 -> bf16 weight packing, attention rollout.  All accesses are 128-bit and
// coalesced; reductions are warp-shuffle based.  Reference arithmetic:
//   LayerNorm(eps = 1e-6)               torchvision vision_transformer.py:96,105,134,175
//   _process_input / class token / pos  vision_transformer.py:268-287, 295-296, Encoder.forward 154-157
//   rollout                             not in the reference (north_star feature); defined by oracle/vit_oracle.py
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdint>
#include "ptx.cuh"

namespace vitb200 {

// Width (in columns) of one LayerNorm partial-sum slot of a row of width d: the column group one epilogue warp of the
// producing GEMM owns (gemm.cuh: BN = 256 -> 128 columns, BN = 128 -> 64 columns).  Everything that writes or reads
// the partial sums (GEMM epilogues, cls_rows / rows_bf16_stats below, the consuming GEMM's role 3) agrees on it.
__host__ __device__ constexpr int ln_slot_width(int d) { return (d % 256 == 0) ? 128 : 64; }

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
                          const float* __restrict__ beta, __nv_bfloat16* __restrict__ y, int rows, float eps,
                          __nv_bfloat16* __restrict__ y_lo = nullptr /* fp32x3 mode: low halves */) {
  ptx::grid_dep_launch(), ptx::grid_dep_wait();   // PDL (ptx.cuh)
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
    const float r0 = (v[i].x - mean) * rstd * g.x + bb.x, r1 = (v[i].y - mean) * rstd * g.y + bb.y;
    const float r2 = (v[i].z - mean) * rstd * g.z + bb.z, r3 = (v[i].w - mean) * rstd * g.w + bb.w;
    const __nv_bfloat162 h01 = __floats2bfloat162_rn(r0, r1), h23 = __floats2bfloat162_rn(r2, r3);
    uint2 o;
    o.x = *reinterpret_cast<const uint32_t*>(&h01), o.y = *reinterpret_cast<const uint32_t*>(&h23);
    yr[lane + 32 * i] = o;
    if (y_lo != nullptr) {
      uint2 l;
      l.x = pack_bf16x2(r0 - __low2float(h01), r1 - __high2float(h01));
      l.y = pack_bf16x2(r2 - __low2float(h23), r3 - __high2float(h23));
      reinterpret_cast<uint2*>(y_lo + static_cast<long>(warp) * d)[lane + 32 * i] = l;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Preprocessing node (`<model>:transform`; the reference's VggModel applies weights.transforms() on the CPU,
// static/models/vgg16.py:40-42): torchvision's ImageClassification preset (transforms/_presets.py:58-65) =
// antialiased bilinear resize (shorter side -> `resize`), centre crop, normalise.  The resize follows ATen's
// _upsample_bilinear2d_aa (aten/src/ATen/native/cpu/UpSampleKernel.cpp, _compute_indices_min_size_weights_aa):
//     scale = in / out, support = max(scale, 1), centre = scale * (i + 0.5),
//     taps [int(centre - support + 0.5), int(centre + support + 0.5)) clipped to the image,
//     weight = max(0, 1 - |(j - centre + 0.5) / max(scale, 1)|), normalised to sum 1,
// horizontal taps accumulated first, then vertical (the order of ATen's two separable passes).
// One thread per output pixel (crop x crop x 3 x B); in: fp32 [B, 3, H, W] in [0, 1]; out: fp32 [B, 3, crop, crop].
struct PreprocessParams {
  int B, H, W;            // input
  int RH, RW;             // resized image
  int crop, top, left;    // centre crop window inside the resized image
  float mean[3], inv_std[3];
};

__device__ __forceinline__ void aa_taps(int i, float scale, int in_size, int& lo, int& n, float& centre, float& invscale,
                                        float& total) {
  const float support = scale >= 1.0f ? scale : 1.0f;
  centre = scale * (static_cast<float>(i) + 0.5f);
  invscale = scale >= 1.0f ? 1.0f / scale : 1.0f;
  lo = max(static_cast<int>(centre - support + 0.5f), 0);
  n = min(static_cast<int>(centre + support + 0.5f), in_size) - lo;
  total = 0.f;
  for (int j = 0; j < n; ++j) total += fmaxf(0.f, 1.0f - fabsf((static_cast<float>(j + lo) - centre + 0.5f) * invscale));
}

__global__ void __launch_bounds__(256)
preprocess_kernel(const float* __restrict__ in, float* __restrict__ out, PreprocessParams p) {
  const long idx = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long total_px = static_cast<long>(p.B) * 3 * p.crop * p.crop;
  if (idx >= total_px) return;
  const int x = idx % p.crop;
  const int y = (idx / p.crop) % p.crop;
  const int c = (idx / (static_cast<long>(p.crop) * p.crop)) % 3;
  const int b = idx / (3L * p.crop * p.crop);
  const float sx = static_cast<float>(p.W) / static_cast<float>(p.RW);
  const float sy = static_cast<float>(p.H) / static_cast<float>(p.RH);
  int x0, nx, y0, ny;
  float cx, ix, tx, cy, iy, ty;
  aa_taps(x + p.left, sx, p.W, x0, nx, cx, ix, tx);
  aa_taps(y + p.top, sy, p.H, y0, ny, cy, iy, ty);
  const float* plane = in + (static_cast<long>(b) * 3 + c) * p.H * p.W;
  float acc = 0.f;
  for (int j = 0; j < ny; ++j) {
    const float wy = fmaxf(0.f, 1.0f - fabsf((static_cast<float>(j + y0) - cy + 0.5f) * iy)) / ty;
    const float* row = plane + static_cast<long>(y0 + j) * p.W + x0;
    float h = 0.f;
    for (int i = 0; i < nx; ++i) {
      const float wx = fmaxf(0.f, 1.0f - fabsf((static_cast<float>(i + x0) - cx + 0.5f) * ix)) / tx;
      h = fmaf(row[i], wx, h);
    }
    acc = fmaf(h, wy, acc);
  }
  out[idx] = (acc - p.mean[c]) * p.inv_std[c];
}

// ---------------------------------------------------------------------------------------------------
// Patch gather: images fp32 [B, 3, S, S] -> patch matrix bf16 [B * n, 3 * p * p] with
// column = c * p * p + ky * p + kx (the flattening of conv_proj.weight [d, 3, p, p]), row = b * n + py * np + px.
// One thread moves 8 consecutive kx: two float4 reads, one 16-byte write.
__global__ void __launch_bounds__(256)
patchify_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out, int B, int S, int p,
                __nv_bfloat16* __restrict__ out_lo = nullptr /* fp32x3 mode: low halves */) {
  ptx::grid_dep_launch(), ptx::grid_dep_wait();   // PDL (ptx.cuh)
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
  if (out_lo != nullptr) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&o);
    uint4 l;
    l.x = pack_bf16x2(a.x - __low2float(h[0]), a.y - __high2float(h[0]));
    l.y = pack_bf16x2(a.z - __low2float(h[1]), a.w - __high2float(h[1]));
    l.z = pack_bf16x2(bb.x - __low2float(h[2]), bb.y - __high2float(h[2]));
    l.w = pack_bf16x2(bb.z - __low2float(h[3]), bb.w - __high2float(h[3]));
    *reinterpret_cast<uint4*>(out_lo + row * (3L * p * p) + col) = l;
  }
}

// ---- LayerNorm partial sums in ONE canonical order -----------------------------------------------------------------
// The folded LayerNorm (gemm.cuh) reads per-row partial sums (sum x, sum x^2) per 64- or 128-column slot.  Three
// producers write them: the residual / patch-embedding GEMM epilogues, cls_rows_kernel (class-token rows) and
// rows_bf16_stats_kernel (token streams that enter through the boundary, vitb200_set_tokens).  All three form a slot in
// the SAME order, so that a request whose tokens were uploaded again (interleaved requests, a client that edits the
// tokens) continues bit-identically to one whose tokens stayed resident:
//   quad   q_k = (x0 + x1) + (x2 + x3)  /  fma(x0, x0, x1 x1) + fma(x2, x2, x3 x3)  over columns 4k .. 4k+3 (explicit
//          intrinsics: no contraction choice left to the compiler),
//   chunk  f_c = ((q0 + q4) + (q2 + q6)) + ((q1 + q5) + (q3 + q7))  over the 8 quads of 32 columns (every level is ONE
//          commutative add of two partners: the result does not depend on which lane holds which),
//   slot   64 columns: f_0 + f_1;  128 columns: (f_0 + f_1) + (f_2 + f_3).
__device__ __forceinline__ float quad_sum(const float4& t) { return __fadd_rn(__fadd_rn(t.x, t.y), __fadd_rn(t.z, t.w)); }
__device__ __forceinline__ float quad_sumsq(const float4& t) {
  return __fadd_rn(__fmaf_rn(t.x, t.x, __fmul_rn(t.y, t.y)), __fmaf_rn(t.z, t.z, __fmul_rn(t.w, t.w)));
}
// Lane l of a warp holds quad (l & 7) of chunk (l >> 3) of 128 consecutive columns: returns the slot sum that covers this
// lane's chunk (sw = 64: chunks {0,1} / {2,3}; sw = 128: all four), identical on every lane of the slot.
__device__ __forceinline__ float slot_sum_128(float q, int sw) {
  q = __fadd_rn(q, __shfl_xor_sync(0xffffffffu, q, 4));
  q = __fadd_rn(q, __shfl_xor_sync(0xffffffffu, q, 2));
  q = __fadd_rn(q, __shfl_xor_sync(0xffffffffu, q, 1));    // f_c on the 8 lanes of chunk c
  q = __fadd_rn(q, __shfl_xor_sync(0xffffffffu, q, 8));    // f_0 + f_1 | f_2 + f_3
  const float o = __shfl_xor_sync(0xffffffffu, q, 16);
  return sw == 128 ? __fadd_rn(q, o) : q;
}
// One warp, 128 columns [c0, c0 + 128) of one row: fp32 value v (this lane's 4 columns) -> bf16 copy (+ low halves) and
// the slot statistics.  Columns at or beyond d contribute nothing.
__device__ __forceinline__ void row128_bf16_stats(const float4& v, long row, int c0, int d, int sw, int lane,
                                                  __nv_bfloat16* __restrict__ xb, __nv_bfloat16* __restrict__ xb_lo,
                                                  float2* __restrict__ stats) {
  const int col = c0 + 4 * lane;   // lane l: chunk l >> 3, quad l & 7 -> columns c0 + 32 (l >> 3) + 4 (l & 7)
  const bool ok = col < d;
  if (ok && xb != nullptr) {
    const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
    uint2 pk;
    pk.x = *reinterpret_cast<const uint32_t*>(&h01), pk.y = *reinterpret_cast<const uint32_t*>(&h23);
    *reinterpret_cast<uint2*>(xb + row * d + col) = pk;
    if (xb_lo != nullptr) {
      const __nv_bfloat162 l01 = __floats2bfloat162_rn(v.x - __low2float(h01), v.y - __high2float(h01));
      const __nv_bfloat162 l23 = __floats2bfloat162_rn(v.z - __low2float(h23), v.w - __high2float(h23));
      uint2 pl;
      pl.x = *reinterpret_cast<const uint32_t*>(&l01), pl.y = *reinterpret_cast<const uint32_t*>(&l23);
      *reinterpret_cast<uint2*>(xb_lo + row * d + col) = pl;
    }
  }
  const float s1 = slot_sum_128(ok ? quad_sum(v) : 0.f, sw), s2 = slot_sum_128(ok ? quad_sumsq(v) : 0.f, sw);
  // one writer per slot: the first lane of the slot's first chunk
  if (ok && stats != nullptr && (lane & (sw == 128 ? 31 : 15)) == 0) stats[row * (d / sw) + col / sw] = make_float2(s1, s2);
}

// Class-token rows of the token stream: x[b, 0, :] = class_token + pos_embedding[0]  (fp32), plus what the folded
// LayerNorm of the next GEMM needs for these rows: the bf16 copy and the per-slot partial sums.  One warp per
// (image, 128 columns).
__global__ void __launch_bounds__(256)
cls_rows_kernel(const float* __restrict__ cls, const float* __restrict__ pos, float* __restrict__ x,
                __nv_bfloat16* __restrict__ xb, float2* __restrict__ stats, int B, int N, int d, int sw,
                __nv_bfloat16* __restrict__ xb_lo = nullptr) {
  ptx::grid_dep_launch(), ptx::grid_dep_wait();   // PDL (ptx.cuh)
  const int groups = (d + 127) / 128;
  const long warp = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= static_cast<long>(B) * groups) return;
  const int b = static_cast<int>(warp / groups), c0 = static_cast<int>(warp % groups) * 128;
  const long row = static_cast<long>(b) * N;
  const int col = c0 + 4 * lane;
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (col < d) {
    const float4 a = *reinterpret_cast<const float4*>(cls + col), p = *reinterpret_cast<const float4*>(pos + col);
    v = make_float4(a.x + p.x, a.y + p.y, a.z + p.z, a.w + p.w);
    *reinterpret_cast<float4*>(x + row * d + col) = v;
  }
  row128_bf16_stats(v, row, c0, d, sw, lane, xb, xb_lo, xb != nullptr ? stats : nullptr);
}

// bf16 copy + partial LayerNorm sums of arbitrary fp32 rows (token streams that enter through the boundary,
// vitb200_set_tokens): one warp per row.
__global__ void __launch_bounds__(256)
rows_bf16_stats_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ xb, float2* __restrict__ stats, long rows,
                       int d, int sw, __nv_bfloat16* __restrict__ xb_lo = nullptr) {
  const long row = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  for (int c0 = 0; c0 < d; c0 += 128) {
    const int col = c0 + 4 * lane;
    const float4 v = col < d ? *reinterpret_cast<const float4*>(x + row * d + col) : make_float4(0.f, 0.f, 0.f, 0.f);
    row128_bf16_stats(v, row, c0, d, sw, lane, xb, xb_lo, stats);
  }
}

// LayerNorm folding, weight side (once, after the weights are loaded).  For a Linear that consumes LayerNorm(x):
//     W'[n, k] = bf16(gamma[k] * W[n, k]),   colsum[n] = sum_k float(W'[n, k]),   bias'[n] = bias[n] + sum_k beta[k] W[n, k]
// so that LN(x) W^T + bias = rstd * (x W'^T - mean * colsum) + bias'.  colsum is taken over the ROUNDED weights: the
// identity then holds exactly for what the tensor cores accumulate.  One warp per output row n.
__global__ void __launch_bounds__(256)
fold_ln_weight_kernel(const float* __restrict__ W, const float* __restrict__ gamma, const float* __restrict__ beta,
                      const float* __restrict__ bias, __nv_bfloat16* __restrict__ Wq, float* __restrict__ colsum,
                      float* __restrict__ bias_out, int N, int K,
                      __nv_bfloat16* __restrict__ Wq_lo = nullptr /* fp32x3 mode: low halves, included in colsum */) {
  const int n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (n >= N) return;
  float s = 0.f, bb = 0.f;
  for (int k = lane; k < K; k += 32) {
    const float w = W[static_cast<long>(n) * K + k];
    const float wg = w * gamma[k];
    const __nv_bfloat16 wq = __float2bfloat16_rn(wg);
    Wq[static_cast<long>(n) * K + k] = wq;
    s += __bfloat162float(wq);
    if (Wq_lo != nullptr) {
      const __nv_bfloat16 wl = __float2bfloat16_rn(wg - __bfloat162float(wq));
      Wq_lo[static_cast<long>(n) * K + k] = wl;
      s += __bfloat162float(wl);
    }
    bb = fmaf(beta[k], w, bb);
  }
  s = warp_sum(s), bb = warp_sum(bb);
  if (lane == 0) colsum[n] = s, bias_out[n] = bias[n] + bb;
}

// fp32 pair -> packed fp16 with saturation to the largest finite value (one F2FP.SATFINITE): the fused attention
// kernels take V in FP16, and an activation beyond 65504 must not become an infinity.
__device__ __forceinline__ uint32_t pack_f16x2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// Copy of a packed bf16 qkv activation [rows, 3d] with the V third converted to FP16 (what the qkv GEMM writes directly
// on the forward path; this kernel serves the single-kernel attention entry point, which takes bf16 inputs).
__global__ void __launch_bounds__(256)
qkv_v_to_f16_kernel(const __nv_bfloat16* __restrict__ in, uint16_t* __restrict__ out, long rows, int d) {
  const long i = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) * 8;   // 8 elements = 16 bytes
  if (i >= rows * 3 * d) return;
  uint4 v = *reinterpret_cast<const uint4*>(in + i);
  if (static_cast<int>(i % (3 * d)) >= 2 * d) {
    uint32_t* w = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&w[k]);
      w[k] = pack_f16x2_sat(__low2float(b), __high2float(b));
    }
  }
  *reinterpret_cast<uint4*>(out + i) = v;
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

// fp32 -> split bf16 (hi = bf16(x), lo = bf16(x - hi)): operands of the fp32x3 precision mode.
__global__ void __launch_bounds__(256)
f32_to_bf16_split_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                         long n) {
  const long i = static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = in[i];
  const __nv_bfloat16 h = __float2bfloat16_rn(x);
  hi[i] = h;
  lo[i] = __float2bfloat16_rn(x - __bfloat162float(h));
}

// ---------------------------------------------------------------------------------------------------
// Attention rollout, class-token row only.  For layer l with head-averaged map Abar_l [N, N]:
//     Ahat_l = rownorm(0.5 * Abar_l + 0.5 * I),      R = Ahat_L ... Ahat_1,      out = R[0, 1:]
// Only row 0 of R is needed, so r <- r * Ahat_l is evaluated right-to-left over layers L, L-1, ..., 1:
// L vector-matrix products of N x N instead of L matrix-matrix products.  With w_k = r_k / rowsum_k:
//     (r * Ahat)_j = 0.5 * sum_k w_k Abar[k, j] + 0.5 * w_j
//
// HBM-bound: L * N * ld * 4 bytes per image, each byte read ONCE.  One 256-thread CTA per image streams the
// maps through a ring of shared-memory stages with bulk async copies (cp.async.bulk + mbarrier), kRolloutRows
// rows per stage; the copies do not depend on r, so they run ahead across layer boundaries while the
// layer-to-layer dependency (the 197-float vector) stays on chip.  Per stage: every warp reduces the row sums of
// two rows (coalesced smem reads), then each thread accumulates its column(s) over the stage's rows.
// For the top layer r = e_0, so only row 0 is fetched.  maps: [L][B, N, ld] fp32; N <= kRolloutThreads * 3.
constexpr int kRolloutThreads = 256;
constexpr int kRolloutRows = 16;      // rows per stage
constexpr int kRolloutMaxCols = 3;    // columns per thread: N <= 768
__host__ __device__ inline int rollout_stage_bytes(int ld) { return kRolloutRows * ld * 4; }
__host__ __device__ inline int rollout_smem_bytes(int ld, int stages) {
  return stages * rollout_stage_bytes(ld) + 2 * kRolloutMaxCols * kRolloutThreads * 4 + kRolloutRows * 4 + 16 * 8 +
         4 * ld * 4 /* partial-product scratch */;
}

// one (layer, chunk) cursor of the streaming schedule; the top layer contributes only row 0 (r = e_0 there)
struct RolloutCursor {
  int layer, chunk, stage, parity;
  __device__ void advance(int L, int chunks_per_layer, int stages) {
    const int n = (layer == L - 1) ? 1 : chunks_per_layer;
    if (++chunk == n) chunk = 0, --layer;
    if (++stage == stages) stage = 0, parity ^= 1;
  }
};

__global__ void __launch_bounds__(kRolloutThreads)
rollout_cls_kernel(const float* __restrict__ maps, long layer_stride, int L, int N, int ld, int stages,
                   float* __restrict__ out /*[B, N-1]*/) {
  extern __shared__ __align__(128) uint8_t rsm[];
  const int stage_bytes = rollout_stage_bytes(ld);
  float* r = reinterpret_cast<float*>(rsm + stages * stage_bytes);   // current row vector [<= 768]
  float* w = r + kRolloutMaxCols * kRolloutThreads;                   // r_k / rowsum_k
  float* wchunk = w + kRolloutMaxCols * kRolloutThreads;              // the stage's w_k
  uint64_t* full = reinterpret_cast<uint64_t*>(wchunk + kRolloutRows);   // [<= 16] one mbarrier per stage; scratch follows
  float4* scratch = reinterpret_cast<float4*>(wchunk + kRolloutRows + 32);  // [4 row groups][ld] partial products
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int chunks_per_layer = (N + kRolloutRows - 1) / kRolloutRows;
  const int ld4 = ld >> 2;
  const float* img = maps + static_cast<long>(b) * N * ld;

  auto issue = [&](const RolloutCursor& c) {
    const int row0 = c.chunk * kRolloutRows;
    const int rows = (c.layer == L - 1) ? 1 : min(kRolloutRows, N - row0);
    const uint32_t bytes = static_cast<uint32_t>(rows) * ld * 4;
    const float* src = img + c.layer * layer_stride + static_cast<long>(row0) * ld;
    const uint32_t bar = static_cast<uint32_t>(__cvta_generic_to_shared(&full[c.stage]));
    const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(rsm + c.stage * stage_bytes));
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
  };
  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      const uint32_t bar = static_cast<uint32_t>(__cvta_generic_to_shared(&full[s]));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int j = tid; j < kRolloutMaxCols * kRolloutThreads; j += kRolloutThreads) r[j] = (j == 0) ? 1.0f : 0.0f, w[j] = 0.0f;
  __syncthreads();
  ptx::grid_dep_launch(), ptx::grid_dep_wait();   // PDL (ptx.cuh): the maps are the previous kernels' output
  RolloutCursor prod{L - 1, 0, 0, 0};   // next chunk to fetch (thread 0)
  if (tid == 0)
    for (int q = 0; q < stages && prod.layer >= 0; ++q) issue(prod), prod.advance(L, chunks_per_layer, stages);

  float4 acc4[kRolloutMaxCols];
#pragma unroll
  for (int i = 0; i < kRolloutMaxCols; ++i) acc4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const int g = tid >> 6, t = tid & 63;   // accumulate pass: row group, float4 column
  RolloutCursor cons{L - 1, 0, 0, 0};
  while (cons.layer >= 0) {
    const int row0 = cons.chunk * kRolloutRows;
    const int rows = (cons.layer == L - 1) ? 1 : min(kRolloutRows, N - row0);
    const bool layer_end = (cons.layer == L - 1) || (cons.chunk == chunks_per_layer - 1);
    const uint32_t bar = static_cast<uint32_t>(__cvta_generic_to_shared(&full[cons.stage]));
    uint32_t ok = 0;
    while (!ok) {
      asm volatile(
          "{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}\n"
          : "=r"(ok)
          : "r"(bar), "r"(cons.parity)
          : "memory");
    }
    const float* A = reinterpret_cast<const float*>(rsm + cons.stage * stage_bytes);
    // row sums -> w_k = r_k / (0.5 * rowsum_k + 0.5): 8 warps x 2 rows, both rows' 128-bit reads in flight together.
    // Only the float4 that straddles column N needs masking (pad columns may hold anything).
    {
      const int k0 = warp, k1 = warp + kRolloutThreads / 32;
      const float4* rowa = reinterpret_cast<const float4*>(A + k0 * ld);
      const float4* rowb = reinterpret_cast<const float4*>(A + k1 * ld);
      const bool has_b = k1 < rows;
      float sa = 0.f, sb = 0.f;
      if (k0 < rows) {
        for (int c = lane; c < ld4; c += 32) {
          float4 va = rowa[c];
          float4 vb = has_b ? rowb[c] : make_float4(0.f, 0.f, 0.f, 0.f);
          const int j = 4 * c;
          if (j + 3 >= N) {
            if (j >= N) va.x = 0.f, vb.x = 0.f;
            if (j + 1 >= N) va.y = 0.f, vb.y = 0.f;
            if (j + 2 >= N) va.z = 0.f, vb.z = 0.f;
            va.w = 0.f, vb.w = 0.f;
          }
          sa += (va.x + va.y) + (va.z + va.w);
          sb += (vb.x + vb.y) + (vb.z + vb.w);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          sa += __shfl_xor_sync(0xffffffffu, sa, o);
          sb += __shfl_xor_sync(0xffffffffu, sb, o);
        }
        if (lane == 0) {
          const float wk = __fdividef(r[row0 + k0], 0.5f * sa + 0.5f);
          wchunk[k0] = wk, w[row0 + k0] = wk;
          if (has_b) {
            const float wk1 = __fdividef(r[row0 + k1], 0.5f * sb + 0.5f);
            wchunk[k1] = wk1, w[row0 + k1] = wk1;
          }
        }
      }
    }
    __syncthreads();
    // columns: thread t of row-group g (4 groups of 64 threads) owns columns 4t .. 4t+3 (+256 i) and the stage's
    // rows k = g, g+4, ...; the four partial sums meet at the end of the layer
#pragma unroll
    for (int i = 0; i < kRolloutMaxCols; ++i) {
      const int c4 = t + 64 * i;
      if (c4 < ld4) {
        float4 a = acc4[i];
#pragma unroll
        for (int kk = 0; kk < kRolloutRows / 4; ++kk) {
          const int k = g + 4 * kk;
          if (k < rows) {
            const float wk = wchunk[k];
            const float4 v = reinterpret_cast<const float4*>(A + k * ld)[c4];
            a.x = fmaf(wk, v.x, a.x), a.y = fmaf(wk, v.y, a.y), a.z = fmaf(wk, v.z, a.z), a.w = fmaf(wk, v.w, a.w);
          }
        }
        acc4[i] = a;
      }
    }
    if (layer_end) {
#pragma unroll
      for (int i = 0; i < kRolloutMaxCols; ++i) {
        const int c4 = t + 64 * i;
        if (c4 < ld4) scratch[g * ld4 + c4] = acc4[i];
        acc4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    __syncthreads();  // the stage and wchunk are free again (and the partial products are staged)
    if (tid == 0 && prod.layer >= 0) issue(prod), prod.advance(L, chunks_per_layer, stages);
    if (layer_end) {
      // r <- 0.5 * (sum of the four partial products) + 0.5 * w
      const float* sc = reinterpret_cast<const float*>(scratch);
      for (int j = tid; j < N; j += kRolloutThreads) {
        const float a = (sc[j] + sc[ld + j]) + (sc[2 * ld + j] + sc[3 * ld + j]);
        r[j] = 0.5f * a + 0.5f * w[j];
      }
      __syncthreads();   // (w needs no reset: rows 1.. of the top layer keep their initial 0, later layers rewrite all)
    }
    cons.advance(L, chunks_per_layer, stages);
  }
  __syncthreads();
  for (int j = tid + 1; j < N; j += kRolloutThreads) out[static_cast<long>(b) * (N - 1) + j - 1] = r[j];
}


// Warp-autonomous variant of rollout_cls_kernel (the default for one-CTA-per-image launches; VITB200_ROLLOUT_WARP=0 keeps
// the block-synchronous kernel above).  That kernel meets at two __syncthreads per 16-row chunk (row sums -> w_k ->
// accumulate), ~2,000 cycles per 13 KB chunk where the arithmetic needs ~300: 160-260 us for 503 MB at batch 256 (HBM
// floor 78 us).  Here every consumer warp owns rows w and w + 8 of EVERY chunk outright: it sums them, forms w_k and adds
// w_k * row into accumulators of its own (lane l: float4 columns l, l + 32, ...), with no block-wide meeting inside a
// layer; a chunk goes back to the producer warp through an `empty` mbarrier (8 arrivals).  The eight partial vectors meet
// once per layer in a fixed tree: bit-reproducible.  9 warps: 8 consumers + 1 producer.
constexpr int kRolloutWarpThreads = 288;
constexpr int kRolloutLaneCols = 6;     // float4 columns per lane: ld / 4 <= 192
__host__ __device__ inline int rollout_warp_smem_bytes(int ld, int stages) {
  return stages * rollout_stage_bytes(ld) + 2 * kRolloutMaxCols * kRolloutThreads * 4 /* r, w */ + 2 * 16 * 8 /* barriers */ +
         8 * ld * 4 /* the warps' partial vectors */;
}

__global__ void __launch_bounds__(kRolloutWarpThreads, 2)
rollout_warp_kernel(const float* __restrict__ maps, long layer_stride, int L, int N, int ld, int stages,
                    float* __restrict__ out /*[B, N-1]*/) {
  extern __shared__ __align__(128) uint8_t rsm[];
  const int stage_bytes = rollout_stage_bytes(ld);
  constexpr int kVec = kRolloutMaxCols * kRolloutThreads;
  float* r = reinterpret_cast<float*>(rsm + stages * stage_bytes);
  float* w = r + kVec;
  uint64_t* full = reinterpret_cast<uint64_t*>(w + kVec);
  uint64_t* empty = full + 16;
  float* part = reinterpret_cast<float*>(empty + 16);    // [8 warps][ld]
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int chunks_per_layer = (N + kRolloutRows - 1) / kRolloutRows;
  const int ld4 = ld >> 2;
  const float* img = maps + static_cast<long>(b) * N * ld;

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(&full[s]))) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 8;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(&empty[s]))) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int j = tid; j < kVec; j += kRolloutWarpThreads) r[j] = (j == 0) ? 1.0f : 0.0f, w[j] = 0.0f;
  __syncthreads();
  ptx::grid_dep_launch(), ptx::grid_dep_wait();   // PDL (ptx.cuh): the maps are the previous kernels' output

  auto wait_bar = [](uint64_t* barp, uint32_t parity) {
    const uint32_t bar = static_cast<uint32_t>(__cvta_generic_to_shared(barp));
    uint32_t ok = 0;
    while (!ok) {
      asm volatile(
          "{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}\n"
          : "=r"(ok)
          : "r"(bar), "r"(parity)
          : "memory");
    }
  };

  if (warp == 8) {
    // ------------------------------------------------------------ producer: every chunk of every layer, top layer first
    if (lane == 0) {
      int stage = 0;
      uint32_t parity = 0;
      for (int layer = L - 1; layer >= 0; --layer) {
        const int nchunks = (layer == L - 1) ? 1 : chunks_per_layer;   // r = e_0 at the top: only row 0 matters
        for (int chunk = 0; chunk < nchunks; ++chunk) {
          wait_bar(&empty[stage], parity ^ 1);
          const int row0 = chunk * kRolloutRows;
          const int rows = (layer == L - 1) ? 1 : min(kRolloutRows, N - row0);
          const uint32_t bytes = static_cast<uint32_t>(rows) * ld * 4;
          const float* src = img + layer * layer_stride + static_cast<long>(row0) * ld;
          const uint32_t bar = static_cast<uint32_t>(__cvta_generic_to_shared(&full[stage]));
          const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(rsm + stage * stage_bytes));
          asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                       "l"(src), "r"(bytes), "r"(bar)
                       : "memory");
          if (++stage == stages) stage = 0, parity ^= 1;
        }
      }
    }
    return;
  }

  // ---------------------------------------------------------------- consumers (warps 0..7)
  int stage = 0;
  uint32_t parity = 0;
  for (int layer = L - 1; layer >= 0; --layer) {
    float4 acc[kRolloutLaneCols];
#pragma unroll
    for (int i = 0; i < kRolloutLaneCols; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int nchunks = (layer == L - 1) ? 1 : chunks_per_layer;
    for (int chunk = 0; chunk < nchunks; ++chunk) {
      const int row0 = chunk * kRolloutRows;
      const int rows = (layer == L - 1) ? 1 : min(kRolloutRows, N - row0);
      wait_bar(&full[stage], parity);
      const float* A = reinterpret_cast<const float*>(rsm + stage * stage_bytes);
      const int k0 = warp, k1 = warp + 8;
      if (k0 < rows) {
        const bool has_b = k1 < rows;
        const float4* rowa = reinterpret_cast<const float4*>(A + k0 * ld);
        const float4* rowb = reinterpret_cast<const float4*>(A + k1 * ld);
        float4 va[kRolloutLaneCols], vb[kRolloutLaneCols];
        float sa = 0.f, sb = 0.f;
#pragma unroll
        for (int i = 0; i < kRolloutLaneCols; ++i) {
          const int c = lane + 32 * i;
          va[i] = make_float4(0.f, 0.f, 0.f, 0.f), vb[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (c < ld4) {
            va[i] = rowa[c];
            if (has_b) vb[i] = rowb[c];
            const int j = 4 * c;
            if (j + 3 >= N) {   // the float4 that straddles column N and the pad columns behind it (may hold anything)
              if (j >= N) va[i].x = 0.f, vb[i].x = 0.f;
              if (j + 1 >= N) va[i].y = 0.f, vb[i].y = 0.f;
              if (j + 2 >= N) va[i].z = 0.f, vb[i].z = 0.f;
              va[i].w = 0.f, vb[i].w = 0.f;
            }
            sa += (va[i].x + va[i].y) + (va[i].z + va[i].w);
            sb += (vb[i].x + vb[i].y) + (vb[i].z + vb[i].w);
          }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          sa += __shfl_xor_sync(0xffffffffu, sa, o);
          sb += __shfl_xor_sync(0xffffffffu, sb, o);
        }
        const float wk0 = __fdividef(r[row0 + k0], 0.5f * sa + 0.5f);
        const float wk1 = has_b ? __fdividef(r[row0 + k1], 0.5f * sb + 0.5f) : 0.f;
        if (lane == 0) {
          w[row0 + k0] = wk0;
          if (has_b) w[row0 + k1] = wk1;
        }
#pragma unroll
        for (int i = 0; i < kRolloutLaneCols; ++i) {   // (masked pad columns contribute w_k * 0)
          acc[i].x = fmaf(wk0, va[i].x, acc[i].x), acc[i].y = fmaf(wk0, va[i].y, acc[i].y);
          acc[i].z = fmaf(wk0, va[i].z, acc[i].z), acc[i].w = fmaf(wk0, va[i].w, acc[i].w);
          acc[i].x = fmaf(wk1, vb[i].x, acc[i].x), acc[i].y = fmaf(wk1, vb[i].y, acc[i].y);
          acc[i].z = fmaf(wk1, vb[i].z, acc[i].z), acc[i].w = fmaf(wk1, vb[i].w, acc[i].w);
        }
      }
      __syncwarp();
      if (lane == 0)
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(&empty[stage]))) : "memory");
      if (++stage == stages) stage = 0, parity ^= 1;
    }
    // ---- end of the layer: the eight warps' partial vectors meet in a fixed tree
    float4* mine = reinterpret_cast<float4*>(part + warp * ld);
#pragma unroll
    for (int i = 0; i < kRolloutLaneCols; ++i) {
      const int c = lane + 32 * i;
      if (c < ld4) mine[c] = acc[i];
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    for (int j = tid; j < N; j += 256) {
      const float a = ((part[j] + part[ld + j]) + (part[2 * ld + j] + part[3 * ld + j])) +
                      ((part[4 * ld + j] + part[5 * ld + j]) + (part[6 * ld + j] + part[7 * ld + j]));
      r[j] = 0.5f * a + 0.5f * w[j];
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    // (w needs no reset: rows 1.. of the top layer keep their initial 0, later layers rewrite every row)
  }
  for (int j = tid + 1; j < N; j += 256) out[static_cast<long>(b) * (N - 1) + j - 1] = r[j];
}

// Class-token rollout for SMALL batches: the same recurrence as rollout_cls_kernel, one thread-block CLUSTER of C CTAs
// per image instead of one CTA (a single-image request ran 12 layers x 13 chunks serially on ONE of 148 SMs: 135 us of a
// 0.92 ms forward; ViT-H at batch 16: 1.3 ms of an 18 ms step on 16 SMs).  CTA `rank` streams the 16-row chunks
// rank, rank + C, ... of every layer; at the end of a layer each CTA pushes its partial product vector and the w_k of its
// rows into every peer's shared memory (DSMEM stores), one cluster barrier, and every CTA forms the same
//     r[j] = 0.5 * (part_0[j] + part_1[j] + ... + part_{C-1}[j]) + 0.5 * w[j]
// in the same order: bit-reproducible for a given C.  (C depends on the batch size, so the last bits of a rollout differ
// between batch sizes -- as they already do through the head-averaged maps' split CTAs; see test_bench_size_properties.)
// Exchange buffers alternate with the layer parity, so one barrier per layer suffices: a CTA can only push layer l - 2's
// values after barrier l - 1, which every peer enters after it has read layer l's.
constexpr int kRolloutMaxCluster = 8;
__host__ __device__ inline int rollout_cluster_smem_bytes(int ld, int stages, int C) {
  return stages * rollout_stage_bytes(ld) + (1 + 2) * kRolloutMaxCols * kRolloutThreads * 4 /* r, w x 2 */ + kRolloutRows * 4 + 16 * 8 +
         4 * ld * 4 /* group partials */ + 2 * C * ld * 4 /* peers' partial vectors x 2 */;
}

__device__ __forceinline__ void st_cluster_f32(float* local_ptr, uint32_t cta, float v) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\tst.shared::cluster.f32 [ra], %2;\n\t}" ::"r"(
          static_cast<uint32_t>(__cvta_generic_to_shared(local_ptr))),
      "r"(cta), "f"(v)
      : "memory");
}

__global__ void __launch_bounds__(kRolloutThreads)
rollout_cluster_kernel(const float* __restrict__ maps, long layer_stride, int L, int N, int ld, int stages, int C,
                       float* __restrict__ out /*[B, N-1]*/) {
  extern __shared__ __align__(128) uint8_t rsm[];
  const int stage_bytes = rollout_stage_bytes(ld);
  constexpr int kVec = kRolloutMaxCols * kRolloutThreads;
  float* r = reinterpret_cast<float*>(rsm + stages * stage_bytes);   // current row vector
  float* w2 = r + kVec;                                               // [2 parities][kVec] r_k / rowsum_k of the layer
  float* wchunk = w2 + 2 * kVec;                                      // the stage's w_k
  uint64_t* full = reinterpret_cast<uint64_t*>(wchunk + kRolloutRows);
  float4* scratch = reinterpret_cast<float4*>(wchunk + kRolloutRows + 32);   // [4 row groups][ld] partial products
  float* xpart = reinterpret_cast<float*>(scratch) + 4 * ld;                 // [2 parities][C][ld] every CTA's partial vector
  const uint32_t rank = ptx::cluster_ctarank();
  const int b = blockIdx.x / C;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int chunks_per_layer = (N + kRolloutRows - 1) / kRolloutRows;
  const int ld4 = ld >> 2;
  const float* img = maps + static_cast<long>(b) * N * ld;
  // chunks of a layer that belong to this CTA: rank, rank + C, ... (the top layer has one chunk -- row 0 -- owned by rank 0)
  auto owned = [&](int layer) {
    if (layer == L - 1) return rank == 0 ? 1 : 0;
    return chunks_per_layer > static_cast<int>(rank) ? (chunks_per_layer - static_cast<int>(rank) + C - 1) / C : 0;
  };

  if (tid == 0) {
    for (int s = 0; s < stages; ++s) {
      const uint32_t bar = static_cast<uint32_t>(__cvta_generic_to_shared(&full[s]));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int j = tid; j < kVec; j += kRolloutThreads) r[j] = (j == 0) ? 1.0f : 0.0f, w2[j] = 0.0f, w2[kVec + j] = 0.0f;
  __syncthreads();
  ptx::cluster_sync();   // every CTA's exchange buffers exist before the first remote store
  ptx::grid_dep_launch(), ptx::grid_dep_wait();   // PDL (ptx.cuh): the maps are the previous kernels' output

  // producer cursor (thread 0): (layer, i-th owned chunk)
  int p_layer = L - 1, p_i = 0, p_stage = 0;
  auto p_skip = [&]() { while (p_layer >= 0 && p_i >= owned(p_layer)) --p_layer, p_i = 0; };
  auto issue = [&]() {
    const int chunk = (p_layer == L - 1) ? 0 : static_cast<int>(rank) + p_i * C;
    const int row0 = chunk * kRolloutRows;
    const int rows = (p_layer == L - 1) ? 1 : min(kRolloutRows, N - row0);
    const uint32_t bytes = static_cast<uint32_t>(rows) * ld * 4;
    const float* src = img + p_layer * layer_stride + static_cast<long>(row0) * ld;
    const uint32_t bar = static_cast<uint32_t>(__cvta_generic_to_shared(&full[p_stage]));
    const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(rsm + p_stage * stage_bytes));
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
    ++p_i;
    if (++p_stage == stages) p_stage = 0;
  };
  if (tid == 0) {
    p_skip();
    for (int q = 0; q < stages && p_layer >= 0; ++q) issue(), p_skip();
  }

  float4 acc4[kRolloutMaxCols];
  const int g = tid >> 6, t = tid & 63;   // accumulate pass: row group, float4 column
  int c_stage = 0;
  uint32_t c_parity = 0;
  for (int layer = L - 1; layer >= 0; --layer) {
    const int par = layer & 1;
    float* w = w2 + par * kVec;
#pragma unroll
    for (int i = 0; i < kRolloutMaxCols; ++i) acc4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int mine = owned(layer);
    for (int ci = 0; ci < mine; ++ci) {
      const int chunk = (layer == L - 1) ? 0 : static_cast<int>(rank) + ci * C;
      const int row0 = chunk * kRolloutRows;
      const int rows = (layer == L - 1) ? 1 : min(kRolloutRows, N - row0);
      const uint32_t bar = static_cast<uint32_t>(__cvta_generic_to_shared(&full[c_stage]));
      uint32_t ok = 0;
      while (!ok) {
        asm volatile(
            "{\n\t.reg .pred P1;\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\tselp.u32 %0, 1, 0, P1;\n\t}\n"
            : "=r"(ok)
            : "r"(bar), "r"(c_parity)
            : "memory");
      }
      const float* A = reinterpret_cast<const float*>(rsm + c_stage * stage_bytes);
      // row sums -> w_k = r_k / (0.5 * rowsum_k + 0.5), as in rollout_cls_kernel; w_k also goes to every peer
      {
        const int k0 = warp, k1 = warp + kRolloutThreads / 32;
        const float4* rowa = reinterpret_cast<const float4*>(A + k0 * ld);
        const float4* rowb = reinterpret_cast<const float4*>(A + k1 * ld);
        const bool has_b = k1 < rows;
        float sa = 0.f, sb = 0.f;
        if (k0 < rows) {
          for (int c = lane; c < ld4; c += 32) {
            float4 va = rowa[c];
            float4 vb = has_b ? rowb[c] : make_float4(0.f, 0.f, 0.f, 0.f);
            const int j = 4 * c;
            if (j + 3 >= N) {
              if (j >= N) va.x = 0.f, vb.x = 0.f;
              if (j + 1 >= N) va.y = 0.f, vb.y = 0.f;
              if (j + 2 >= N) va.z = 0.f, vb.z = 0.f;
              va.w = 0.f, vb.w = 0.f;
            }
            sa += (va.x + va.y) + (va.z + va.w);
            sb += (vb.x + vb.y) + (vb.z + vb.w);
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            sa += __shfl_xor_sync(0xffffffffu, sa, o);
            sb += __shfl_xor_sync(0xffffffffu, sb, o);
          }
          const float wk0 = __fdividef(r[row0 + k0], 0.5f * sa + 0.5f);
          const float wk1 = has_b ? __fdividef(r[row0 + k1], 0.5f * sb + 0.5f) : 0.f;
          if (lane == 0) wchunk[k0] = wk0;
          if (lane == 0 && has_b) wchunk[k1] = wk1;
          if (lane < C) {   // lane c stores into CTA c (its own included)
            st_cluster_f32(&w[row0 + k0], static_cast<uint32_t>(lane), wk0);
            if (has_b) st_cluster_f32(&w[row0 + k1], static_cast<uint32_t>(lane), wk1);
          }
        }
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < kRolloutMaxCols; ++i) {
        const int c4 = t + 64 * i;
        if (c4 < ld4) {
          float4 a = acc4[i];
#pragma unroll
          for (int kk = 0; kk < kRolloutRows / 4; ++kk) {
            const int k = g + 4 * kk;
            if (k < rows) {
              const float wk = wchunk[k];
              const float4 v = reinterpret_cast<const float4*>(A + k * ld)[c4];
              a.x = fmaf(wk, v.x, a.x), a.y = fmaf(wk, v.y, a.y), a.z = fmaf(wk, v.z, a.z), a.w = fmaf(wk, v.w, a.w);
            }
          }
          acc4[i] = a;
        }
      }
      __syncthreads();  // the stage and wchunk are free again
      if (tid == 0 && p_layer >= 0) issue(), p_skip();
      if (++c_stage == stages) c_stage = 0, c_parity ^= 1;
    }
    // ---- end of the layer: this CTA's partial product (four row groups, fixed order) -> every CTA
#pragma unroll
    for (int i = 0; i < kRolloutMaxCols; ++i) {
      const int c4 = t + 64 * i;
      if (c4 < ld4) scratch[g * ld4 + c4] = acc4[i];
    }
    __syncthreads();
    {
      const float* sc = reinterpret_cast<const float*>(scratch);
      float* mine_part = xpart + (par * C + static_cast<int>(rank)) * ld;   // slot `rank` of parity `par`, in every CTA
      for (int j = tid; j < N; j += kRolloutThreads) {
        const float a = (sc[j] + sc[ld + j]) + (sc[2 * ld + j] + sc[3 * ld + j]);
        for (int c = 0; c < C; ++c) st_cluster_f32(&mine_part[j], static_cast<uint32_t>(c), a);
      }
    }
    ptx::cluster_sync();   // release / acquire: every CTA's partials and w_k of this layer are visible everywhere
    {
      const float* parts = xpart + par * C * ld;
      for (int j = tid; j < N; j += kRolloutThreads) {
        float a = parts[j];
        for (int c = 1; c < C; ++c) a += parts[c * ld + j];
        r[j] = 0.5f * a + 0.5f * w[j];
      }
    }
    __syncthreads();
    // (rows the top layer does not touch keep w = 0 in both parity buffers; every later layer rewrites all N rows)
  }
  if (rank == 0)
    for (int j = tid + 1; j < N; j += kRolloutThreads) out[static_cast<long>(b) * (N - 1) + j - 1] = r[j];
  ptx::cluster_sync();   // no CTA leaves while a peer may still address its shared memory
}


// Head-average maps of a small launch whose (image, query tile) items were split over S > 2 CTAs by heads
// (attention.cuh AttnParams::split): avg = ((part_0 + part_1) + part_2) + ..., always in index order -- bit-reproducible,
// which a reduce-add by more than two CTAs is not.  parts: [S][n4] float4, avg: [n4].
constexpr int kAvgPartsMax = 16;
__global__ void __launch_bounds__(256)
avg_parts_sum_kernel(const float4* __restrict__ parts, float4* __restrict__ avg, long n4, int S) {
  ptx::grid_dep_launch(), ptx::grid_dep_wait();   // PDL (ptx.cuh)
  const long i = static_cast<long>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n4) return;
  // every part's load in flight before the first add (a loop of dependent load + add pairs cost one L2 round trip per
  // part: 5.6 us for 12 parts of one image)
  float4 v[kAvgPartsMax];
#pragma unroll
  for (int s = 0; s < kAvgPartsMax; ++s)
    if (s < S) v[s] = parts[s * n4 + i];
  float4 a = v[0];
#pragma unroll
  for (int s = 1; s < kAvgPartsMax; ++s)
    if (s < S) a.x += v[s].x, a.y += v[s].y, a.z += v[s].z, a.w += v[s].w;
  avg[i] = a;
}

}  // namespace vitb200
