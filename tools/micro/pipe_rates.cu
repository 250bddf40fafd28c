// Issue cost of the candidate softmax instructions on sm_100a: cycles per warp-instruction and SM sub-partition with
// 4 resident warps per sub-partition (16 per SM) and 8 independent chains per thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>

template <int OP>
__global__ void __launch_bounds__(512, 1) k(uint32_t* out, long long* cyc, int iters, uint32_t seed) {
  uint32_t v[8];
  float f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = seed + threadIdx.x * 8 + i, f[i] = __uint_as_float(0x3c000000u + v[i]);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(f[i]));
      if (OP == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(v[i]));           // 2 MUFU.EX2.F16 + PRMT
      if (OP == 2) asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %1; add.rn.f32.f16 %0, lo, %0;}" : "+f"(f[i]) : "r"(v[i]));
      if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(f[i]));
      if (OP == 4) asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %1; cvt.f32.f16 %0, lo;}" : "=f"(f[i]) : "r"(v[i] + it));
      if (OP == 5) asm volatile("mul.rn.f16x2 %0, %0, %0;" : "+r"(v[i]));
      if (OP == 6) asm volatile("{.reg .b16 lo, hi; mov.b32 {lo, hi}, %0; ex2.approx.f16 lo, lo; mov.b32 %0, {lo, hi};}" : "+r"(v[i]));
      if (OP == 7) asm volatile("cvt.rn.f16x2.f32 %0, %1, %1;" : "=r"(v[i]) : "f"(f[i] + it));
      if (OP == 8) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(f[(i + 1) & 7]), "f"(f[(i + 2) & 7]));
    }
  }
  const long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) acc ^= v[i] ^ __float_as_uint(f[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  uint32_t* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4);
  cudaMalloc(&cyc, 148 * 8);
  const int iters = 4096;
  const char* names[] = {"ex2.approx.ftz.f32 (MUFU.EX2)", "ex2.approx.f16x2 (2 MUFU.EX2.F16 + PRMT)", "add.rn.f32.f16 (FHADD)",
                         "fma.rn.f32 (FFMA)", "cvt.f32.f16 (+IADD)", "mul.rn.f16x2 (HMUL2)", "ex2.approx.f16 scalar (1 MUFU.EX2.F16)",
                         "cvt.rn.f16x2.f32 (F2FP, +FADD)", "max.f32 3-input (FMNMX3)"};
  for (int op = 0; op < 9; ++op) {
    for (int rep = 0; rep < 2; ++rep) {
      switch (op) {
        case 0: k<0><<<148, 512>>>(out, cyc, iters, 1); break;
        case 1: k<1><<<148, 512>>>(out, cyc, iters, 1); break;
        case 2: k<2><<<148, 512>>>(out, cyc, iters, 1); break;
        case 3: k<3><<<148, 512>>>(out, cyc, iters, 1); break;
        case 4: k<4><<<148, 512>>>(out, cyc, iters, 1); break;
        case 5: k<5><<<148, 512>>>(out, cyc, iters, 1); break;
        case 6: k<6><<<148, 512>>>(out, cyc, iters, 1); break;
        case 7: k<7><<<148, 512>>>(out, cyc, iters, 1); break;
        case 8: k<8><<<148, 512>>>(out, cyc, iters, 1); break;
      }
      cudaDeviceSynchronize();
    }
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    // 4 warps per sub-partition x 8 asm statements x iters per warp
    const double per = (double)h[0] / (4.0 * 8 * iters);
    printf("%-45s %8.2f cycles per warp-statement and sub-partition (%lld cycles)\n", names[op], per, h[0]);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
