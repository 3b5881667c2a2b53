// bring-up probe for the tensor-map TMA loads of the conv weight gradient: is the load legal, what lands where
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <vector>
#include <stdlib.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap m, float* out, int c0, int c1, int c2, int c3, int bytes) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(s32(sm)), "l"(reinterpret_cast<uint64_t>(&m)), "r"(s32(&bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
  }
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(&bar)) : "memory");
  }
  for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) out[i] = reinterpret_cast<float*>(sm)[i];
}
int main(int argc, char** argv) {
  const int W = 32, H = 32, C = 8, B = 2, boxC = 16;
  std::vector<float> h((size_t)W * H * C * B);
  for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
  float *d, *o;
  cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, 65536);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  CUtensorMap m;
  cuuint64_t dims[4] = {W, H, C, B};
  cuuint64_t strides[3] = {W * 4, W * H * 4, (cuuint64_t)W * H * C * 4};
  cuuint32_t box[4] = {32, 1, boxC, 1}, es[4] = {1, 1, 1, 1};
  CUresult r = cuTensorMapEncodeTiled(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc=%d\n", (int)r);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
  const int bytes = 32 * 4 * boxC;
  int tests[4][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
  for (int q = 0; q < 4; ++q) tests[0][q] = atoi(argv[1 + q]);
  for (int ti = 0; ti < 1; ++ti) { int* t = tests[ti];
    cudaMemset(o, 0xff, 65536);
    k<<<1, 128, 40000>>>(m, o, t[0], t[1], t[2], t[3], bytes);
    cudaError_t e = cudaDeviceSynchronize();
    printf("coords (%d,%d,%d,%d): %s\n", t[0], t[1], t[2], t[3], cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> ho(bytes / 4);
    cudaMemcpy(ho.data(), o, bytes, cudaMemcpyDeviceToHost);
    for (int row = 0; row < 10; ++row) {   // row = channel, 32 floats; print the first element of each 16-byte chunk
      printf(" row %2d:", row);
      for (int c = 0; c < 8; ++c) printf(" %8.0f", ho[row * 32 + c * 4]);
      printf("\n");
    }
  }
  return 0;
}
