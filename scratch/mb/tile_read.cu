// What limits the operand-producer loads of dense_kernel?  128 CTAs x 512 threads, each CTA streams its
// 64-sample tile of NS [8192 x 784] arrays in 25 K-chunks of 64 rows x 128 B (the layer-1 prologue
// pattern), with different load schedules.  No MMA, no shared-memory stores.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int D = 784, B = 8192, NT = 64, KC = 24;  // 24 full chunks (768 floats)
template <int NS, int DEPTH, int MODE>  // MODE 0: rotating software pipeline; 1: burst of DEPTH chunks; 2: pipeline with ld.global.cg
__global__ void __launch_bounds__(512, 1) tile_kernel(const float* __restrict__ a, float* out, int reps) {
  const int tid = threadIdx.x, row = tid >> 3, cch = tid & 7;
  const size_t n0 = (size_t)blockIdx.x * NT;
  const float* base = a + (n0 + row) * D + cch * 4;
  const size_t arr = (size_t)D * B;
  float acc = 0.f;
  __shared__ float4 sm4[4096];
  float* sm = reinterpret_cast<float*>(sm4);
  auto ld = [&](int s, int kc) -> float4 {
    const float4* p = reinterpret_cast<const float4*>(base + s * arr + kc * 32);
    if (MODE == 2) { float4 v; asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p)); return v; }
    return *p;
  };
  for (int r = 0; r < reps; ++r) {
    if (MODE == 1) {
      for (int kc0 = 0; kc0 < KC; kc0 += DEPTH) {
        float4 buf[DEPTH][NS];
#pragma unroll
        for (int d = 0; d < DEPTH; ++d)
#pragma unroll
          for (int s = 0; s < NS; ++s) buf[d][s] = ld(s, kc0 + d);
#pragma unroll
        for (int d = 0; d < DEPTH; ++d)
#pragma unroll
          for (int s = 0; s < NS; ++s) acc += buf[d][s].x + buf[d][s].y + buf[d][s].z + buf[d][s].w;
      }
    } else {
      float4 buf[DEPTH][NS];
#pragma unroll
      for (int d = 0; d < DEPTH; ++d)
#pragma unroll
        for (int s = 0; s < NS; ++s) buf[d][s] = ld(s, d);
      for (int kc0 = 0; kc0 < KC; kc0 += DEPTH) {
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
          const int kc = kc0 + d;
#pragma unroll
          for (int s = 0; s < NS; ++s) acc += buf[d][s].x + buf[d][s].y + buf[d][s].z + buf[d][s].w;
          if (MODE >= 3) {  // what the real producer does after the combination
            float4 v = buf[d][0];
#pragma unroll
            for (int s = 1; s < NS; ++s) { v.x += buf[d][s].x; v.y += buf[d][s].y; v.z += buf[d][s].z; v.w += buf[d][s].w; }
            float4* sp = reinterpret_cast<float4*>(sm) + (kc & 3) * 1024 + tid;
            sp[0] = v; sp[512] = v;
            if (MODE == 3 || MODE == 5) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          }
          if (MODE != 5 && kc + DEPTH < KC) {
#pragma unroll
            for (int s = 0; s < NS; ++s) buf[d][s] = ld(s, kc + DEPTH);
          }
          __syncwarp();
          if (MODE == 5 && kc + DEPTH < KC) {   // loads issued after the fence
#pragma unroll
            for (int s = 0; s < NS; ++s) buf[d][s] = ld(s, kc + DEPTH);
          }
        }
      }
    }
  }
  if (acc == 123.456f) out[0] = acc;
}
template <int NS, int DEPTH, int MODE> void run(const float* a, float* o, const char* name) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  tile_kernel<NS, DEPTH, MODE><<<B / NT, 512>>>(a, o, 1);
  cudaEventRecord(e0);
  for (int i = 0; i < 20; ++i) tile_kernel<NS, DEPTH, MODE><<<B / NT, 512>>>(a, o, 1);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 20;
  const double bytes = (double)NS * B * KC * 32 * 4;
  printf("%-40s NS=%d DEPTH=%d: %7.2f us  %7.1f GB/s\n", name, NS, DEPTH, ms * 1e3, bytes / ms / 1e6);
}
int main() {
  float *a, *o;
  cudaMalloc(&a, (size_t)8 * D * B * 4); cudaMalloc(&o, 16);
  cudaMemset(a, 0, (size_t)8 * D * B * 4);
  run<7, 2, 0>(a, o, "pipeline");   run<7, 2, 1>(a, o, "burst");   run<7, 2, 2>(a, o, "pipeline ld.cg");
  run<7, 2, 3>(a, o, "pipe + sts + fence, loads after sts");  run<7, 2, 4>(a, o, "pipe + sts (no fence)"); run<7, 2, 5>(a, o, "pipe + sts + fence, loads after fence");
  run<2, 4, 3>(a, o, "pipe + sts + fence");  run<2, 4, 4>(a, o, "pipe + sts (no fence)"); run<2, 4, 5>(a, o, "loads after fence");
  run<7, 1, 0>(a, o, "pipeline");   run<7, 3, 1>(a, o, "burst");   run<7, 4, 1>(a, o, "burst");
  run<4, 2, 0>(a, o, "pipeline");   run<4, 4, 1>(a, o, "burst");   run<4, 4, 0>(a, o, "pipeline");
  run<2, 4, 0>(a, o, "pipeline");   run<2, 8, 1>(a, o, "burst");   run<2, 8, 0>(a, o, "pipeline");
  run<1, 4, 0>(a, o, "pipeline");   run<1, 8, 1>(a, o, "burst");   run<1, 8, 0>(a, o, "pipeline");
  return 0;
}
