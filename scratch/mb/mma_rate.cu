// microbenchmark: cycles per tcgen05.mma kind::tf32 (SS) for several N, with/without A collector reuse
#include <cstdio>
#include <cuda_runtime.h>
#include "../../localregneuralde.jl_b200/csrc/lrnde_umma.cuh"
using namespace umma;
template <int N, int MODE>  // MODE 0: 3 MMAs per kstep plain; 1: collector fill/lastuse; 2: single MMA per kstep (same A each)
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (64 * 1024) / 4; i += blockDim.x) ((float*)smem)[i] = 0.001f * (i % 97);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_slot;
  if (warp == 0) {
    constexpr uint32_t idesc = make_idesc(128, N);
    const uint32_t a_hi = desc_lo(smem_u32(smem)), a_lo = desc_lo(smem_u32(smem) + 16384);
    const uint32_t b_hi = desc_lo(smem_u32(smem) + 32768), b_lo = desc_lo(smem_u32(smem) + 32768 + 8192 * (N / 64));
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t ko = k * 2;
          if (MODE == 0) {
            mma_tf32_lo<0>(tm, a_lo + ko, b_hi + ko, idesc, 1u);
            mma_tf32_lo<0>(tm, a_hi + ko, b_lo + ko, idesc, 1u);
            mma_tf32_lo<0>(tm, a_hi + ko, b_hi + ko, idesc, 1u);
          } else if (MODE == 1) {
            mma_tf32_lo<0>(tm, a_lo + ko, b_hi + ko, idesc, 1u);
            mma_tf32_lo<1>(tm, a_hi + ko, b_lo + ko, idesc, 1u);
            mma_tf32_lo<2>(tm, a_hi + ko, b_hi + ko, idesc, 1u);
          } else {
            mma_tf32_lo<0>(tm, a_hi + ko, b_hi + ko, idesc, 1u);
            mma_tf32_lo<0>(tm + N, a_hi + ko, b_hi + ko, idesc, 1u);
            mma_tf32_lo<0>(tm, a_hi + ko, b_hi + ko, idesc, 1u);
          }
        }
      }
      __syncwarp();
    }
    if (elect_one_sync()) mma_commit(&bar);
    __syncwarp();
    long long t1 = clock64();
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory"); }
}
template <int MODE>  // MN-major operands (SW128_32B atoms), M = N = 128, as in wgrad_kernel
__global__ void __launch_bounds__(128, 1) rate_mn_kernel(long long* out, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (64 * 1024) / 4; i += blockDim.x) ((float*)smem)[i] = 0.001f * (i % 97);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  fence_proxy_async();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tmem_slot;
  if (warp == 0) {
    constexpr uint32_t idesc = make_idesc_mn(128, 128);
    const uint32_t base = smem_u32(smem);
    const uint32_t p_hi = desc_lo_mn(base), p_lo = desc_lo_mn(base + 16384);
    const uint32_t q_hi = desc_lo_mn(base + 32768), q_lo = desc_lo_mn(base + 49152);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (elect_one_sync()) {
#pragma unroll
        for (int kg = 0; kg < 4; ++kg) {
          const uint32_t ko = (uint32_t)kg * (4096u >> 4);
          if (MODE == 0) {
            mma_tf32_mn<0>(tm, p_lo + ko, q_hi + ko, idesc, 1u);
            mma_tf32_mn<0>(tm, p_hi + ko, q_lo + ko, idesc, 1u);
            mma_tf32_mn<0>(tm, p_hi + ko, q_hi + ko, idesc, 1u);
          } else {
            mma_tf32_mn<0>(tm, p_lo + ko, q_hi + ko, idesc, 1u);
            mma_tf32_mn<1>(tm, p_hi + ko, q_lo + ko, idesc, 1u);
            mma_tf32_mn<2>(tm, p_hi + ko, q_hi + ko, idesc, 1u);
          }
        }
      }
      __syncwarp();
    }
    if (elect_one_sync()) mma_commit(&bar);
    __syncwarp();
    long long t1 = clock64();
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512u) : "memory"); }
}
template <int MODE> void run_mn(const char* name) {
  long long* d; cudaMalloc(&d, 16); long long h[2];
  cudaFuncSetAttribute(rate_mn_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 200;
  rate_mn_kernel<MODE><<<148, 128, 100 * 1024>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%s MN-major M=N=128: issue %.1f cyc/MMA, complete %.1f cyc/MMA (%s)\n", name, h[0] / (12.0 * iters), h[1] / (12.0 * iters), cudaGetErrorString(e));
  cudaFree(d);
}
template <int N, int MODE> void run(const char* name) {
  long long* d; cudaMalloc(&d, 16); long long h[2];
  cudaFuncSetAttribute(rate_kernel<N, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int iters = 200;
  for (int grid : {1, 148}) {
    rate_kernel<N, MODE><<<grid, 128, 100 * 1024>>>(d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    printf("%s N=%d grid=%d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (%s)\n", name, N, grid, h[0] / (12.0 * iters), h[1] / (12.0 * iters), cudaGetErrorString(e));
  }
  cudaFree(d);
}
int main() {
  run<64, 0>("3x plain"); run<64, 1>("3x collector"); run<64, 2>("same-A 2 accum");
  run<128, 0>("3x plain"); run<128, 1>("3x collector");
  run<256, 0>("3x plain"); run<256, 1>("3x collector");
  run_mn<0>("3x plain"); run_mn<1>("3x collector");
  return 0;
}
