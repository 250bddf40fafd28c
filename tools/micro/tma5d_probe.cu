// Where does a 5-D tiled TMA load with the box (16, 2, g, pr, 1) over an image viewed as (kx, ky, px, py, b*3+c) put its
// elements in shared memory (SWIZZLE_128B)?   nvcc -gencode arch=compute_100a,code=sm_100a -o tma5d_probe tma5d_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
__global__ void probe(const __grid_constant__ CUtensorMap tm, float* out, int bytes, int words) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ uint64_t bar;
  float* f = reinterpret_cast<float*>(sm);
  for (int i = threadIdx.x; i < words; i += blockDim.x) f[i] = -1.0f;
  uint32_t b = (uint32_t)__cvta_generic_to_shared(&bar), d = (uint32_t)__cvta_generic_to_shared(sm);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  asm volatile("fence.proxy.async.shared::cta;");
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes));
    asm volatile("cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 ::"r"(d), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(b), "r"(0), "r"(2), "r"(0), "r"(0), "r"(1) : "memory");
  }
  uint32_t ok = 0;
  while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0,1,0,p;}" : "=r"(ok) : "r"(b));
  __syncthreads();
  for (int i = threadIdx.x; i < words; i += blockDim.x) out[i] = f[i];
}
int main() {
  const int S = 224, g = 14, pr = 9, B = 1;
  std::vector<float> img((size_t)B * 3 * S * S);
  for (size_t i = 0; i < img.size(); ++i) img[i] = (float)i;   // value = linear index (exact up to 2^24)
  float* d_img; cudaMalloc(&d_img, img.size() * 4); cudaMemcpy(d_img, img.data(), img.size() * 4, cudaMemcpyHostToDevice);
  CUtensorMap tm;
  cuuint64_t dims[5] = {16, 16, (cuuint64_t)g, (cuuint64_t)g, (cuuint64_t)B * 3};
  cuuint64_t str[4] = {(cuuint64_t)S * 4, 64, (cuuint64_t)16 * S * 4, (cuuint64_t)S * S * 4};
  cuuint32_t box[5] = {16, 2, (cuuint32_t)g, (cuuint32_t)pr, 1}, es[5] = {1, 1, 1, 1, 1};
  cuInit(0);
  CUresult r = cuTensorMapEncodeTiled(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, d_img, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode: %d\n", (int)r);
  const int words = 40 * 1024 / 4, bytes = 16 * 2 * g * pr * 4;
  float* d_out; cudaMalloc(&d_out, words * 4);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
  probe<<<1, 256, 40 * 1024>>>(tm, d_out, bytes, words);
  printf("kernel: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  std::vector<float> out(words);
  cudaMemcpy(out.data(), d_out, words * 4, cudaMemcpyDeviceToHost);
  int written = 0, last = -1;
  for (int i = 0; i < words; ++i) if (out[i] >= 0) ++written, last = i;
  printf("box bytes %d, words written %d (= %d bytes), last written byte offset %d\n", bytes, written, written * 4, last * 4 + 3);
  // loaded at coordinates (kx 0, ky 2, px 0, py 0, bc 1): element (kx, ky, px, py) has value 1*S*S + (py*16 + 2 + ky)*S + px*16 + kx
  auto where = [&](int kx, int ky, int px, int py) {
    const float v = (float)(1 * S * S + (py * 16 + 2 + ky) * S + px * 16 + kx);
    for (int i = 0; i < words; ++i) if (out[i] == v) return i * 4;
    return -1;
  };
  for (int py : {0, 1}) for (int px : {0, 1, 7}) for (int ky : {0, 1}) for (int kx : {0, 4, 5, 15}) {
    const int off = where(kx, ky, px, py);
    const int row = py * g + px, expect = row * 128 + ((((ky * 64 + kx * 4) >> 4) ^ (row & 7)) << 4) + ((kx * 4) & 15);
    printf("(kx %2d ky %d px %d py %d) at byte %6d   expected %6d %s\n", kx, ky, px, py, off, expect, off == expect ? "" : "  <-- differs");
  }
  return 0;
}
