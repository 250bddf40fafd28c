// Issue-rate probe: scalar FFMA vs packed fma.rn.f32x2 on sm_100a (is the packed form 2 FMAs per issue slot?).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_ffma(float* out, int iters) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i;
  float m = 1.0001f, c = 0.5f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], m, c);
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float* out, int iters) {
  unsigned long long a[8];
  for (int i = 0; i < 8; ++i) {
    float lo = threadIdx.x * 0.001f + i, hi = lo + 0.5f;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a[i]) : "f"(lo), "f"(hi));
  }
  unsigned long long m, c;
  asm("mov.b64 %0, {%1, %1};" : "=l"(m) : "f"(1.0001f));
  asm("mov.b64 %0, {%1, %1};" : "=l"(c) : "f"(0.5f));
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(m), "l"(c));
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a[i]));
    s += lo + hi;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
// mixed: one MUFU.EX2 per 4 FMAs, scalar vs packed — does packing free issue slots for the MUFU pipe?
__global__ void k_mix(float* out, int iters) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i * 0.01f;
  float m = 0.999f, c = 0.001f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      a[i] = fmaf(a[i], m, c); a[i] = fmaf(a[i], m, c); a[i] = fmaf(a[i], m, c); a[i] = fmaf(a[i], m, c);
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
    }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mix2(float* out, int iters) {
  unsigned long long a[4];
  for (int i = 0; i < 4; ++i) {
    float lo = threadIdx.x * 0.001f + i * 0.01f, hi = lo + 0.005f;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a[i]) : "f"(lo), "f"(hi));
  }
  unsigned long long m, c;
  asm("mov.b64 %0, {%1, %1};" : "=l"(m) : "f"(0.999f));
  asm("mov.b64 %0, {%1, %1};" : "=l"(c) : "f"(0.001f));
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int r = 0; r < 4; ++r) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(m), "l"(c));
      float lo, hi;
      asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a[i]));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(lo));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(hi));
      asm volatile("mov.b64 %0, {%1, %2};" : "=l"(a[i]) : "f"(lo), "f"(hi));
    }
  }
  float s = 0;
  for (int i = 0; i < 4; ++i) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a[i]));
    s += lo + hi;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  float* out; cudaMalloc(&out, 148 * 1024 * 4);
  const int iters = 20000;
  for (int threads : {128, 256, 512, 1024}) {
    float t1 = timeit([&] { k_ffma<<<148, threads>>>(out, iters); });
    float t2 = timeit([&] { k_ffma2<<<148, threads>>>(out, iters); });
    float t3 = timeit([&] { k_mix<<<148, threads>>>(out, iters); });
    float t4 = timeit([&] { k_mix2<<<148, threads>>>(out, iters); });
    double fma1 = 148.0 * threads * iters * 8 / (t1 * 1e-3) / 1e12, fma2 = 148.0 * threads * iters * 16 / (t2 * 1e-3) / 1e12;
    printf("threads/SM %4d: FFMA %.2f T fma/s (%.3f ms)  FFMA2 %.2f T fma/s (%.3f ms)   mix(4 fma + 1 ex2 per elem, 8 elem): scalar %.3f ms packed %.3f ms\n",
           threads, fma1, t1, fma2, t2, t3, t4);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
