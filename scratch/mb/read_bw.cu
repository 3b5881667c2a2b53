// read-only / copy / tile-pattern bandwidth on B200 (roofline denominators for HBM-bound kernels)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void read_kernel(const float4* __restrict__ a, size_t n4, float* out) {
  float s = 0.f;
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    float4 v0 = a[i], v1 = a[i + stride], v2 = a[i + 2 * stride], v3 = a[i + 3 * stride];
    s += v0.x + v0.y + v0.z + v0.w + v1.x + v1.y + v1.z + v1.w + v2.x + v2.y + v2.z + v2.w + v3.x + v3.y + v3.z + v3.w;
  }
  for (; i < n4; i += stride) { float4 v = a[i]; s += v.x + v.y + v.z + v.w; }
  if (s == 123.456f) out[0] = s;
}
__global__ void copy_kernel(const float4* __restrict__ a, float4* __restrict__ b, size_t n4) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    float4 v0 = a[i], v1 = a[i + stride], v2 = a[i + 2 * stride], v3 = a[i + 3 * stride];
    b[i] = v0; b[i + stride] = v1; b[i + 2 * stride] = v2; b[i + 3 * stride] = v3;
  }
  for (; i < n4; i += stride) b[i] = a[i];
}
// 8 source arrays summed into one output (the lincomb / stage-combination shape)
__global__ void sum8_kernel(const float4* __restrict__ a, size_t n4, float4* __restrict__ out) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n4; i += stride) {
    float4 acc = a[i];
#pragma unroll
    for (int k = 1; k < 8; ++k) { float4 v = a[i + k * n4]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
    out[i] = acc;
  }
}
template <typename F> float timeit(F f, int reps) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); f();
  cudaEventRecord(e0);
  for (int i = 0; i < reps; ++i) f();
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / reps;
}
int main() {
  const size_t bytes = (size_t)1 << 30;  // 1 GiB per array
  float4 *a, *b; float* o;
  cudaMalloc(&a, 2 * bytes); cudaMalloc(&b, bytes); cudaMalloc(&o, 16);
  cudaMemset(a, 0, 2 * bytes); cudaMemset(b, 0, bytes);
  const size_t n4 = bytes / 16;
  for (int blocks : {148 * 4, 148 * 8, 148 * 16, 148 * 32}) {
    float t = timeit([&] { read_kernel<<<blocks, 256>>>(a, n4, o); }, 10);
    printf("read   1 GiB, %5d blocks x 256: %.1f GB/s\n", blocks, bytes / t / 1e6);
  }
  for (int blocks : {148 * 8, 148 * 32}) {
    float t = timeit([&] { copy_kernel<<<blocks, 256>>>(a, b, n4); }, 10);
    printf("copy   1 GiB, %5d blocks x 256: %.1f GB/s (read + write)\n", blocks, 2.0 * bytes / t / 1e6);
  }
  {
    const size_t m4 = (size_t)784 * 8192 / 4;  // one [784 x 8192] fp32 state array = 25.7 MB
    float t = timeit([&] { sum8_kernel<<<148 * 8, 256>>>(a, m4, b); }, 20);
    printf("sum8   8 x 25.7 MB -> 25.7 MB: %.2f us, %.1f GB/s (read + write)\n", t * 1e3, 9.0 * m4 * 16 / t / 1e6);
    float t2 = timeit([&] { read_kernel<<<148 * 8, 256>>>(a, 8 * m4, o); }, 20);
    printf("read   206 MB (L2 126 MB): %.2f us, %.1f GB/s\n", t2 * 1e3, 8.0 * m4 * 16 / t2 / 1e6);
  }
  cudaMemcpy(a, b, bytes, cudaMemcpyDeviceToDevice);
  float t = timeit([&] { cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice); }, 10);
  printf("cudaMemcpy D2D 1 GiB: %.1f GB/s (read + write)\n", 2.0 * bytes / t / 1e6);
  return 0;
}
