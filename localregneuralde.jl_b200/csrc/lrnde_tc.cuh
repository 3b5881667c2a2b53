// tcgen05 / TMEM / mbarrier / bulk-copy PTX helpers shared by the tensor-core kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  const uint32_t addr = smem_u32(bar);
#ifdef LRNDE_WATCHDOG
  const long long t0 = clock64();
#endif
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
#ifdef LRNDE_WATCHDOG
    // debug builds (LRNDE_WATCHDOG=1 python build.py): a barrier that never completes becomes a trap with the
    // place it happened instead of a hung GPU
    if (!ok && clock64() - t0 > 4000000000LL) {
      printf("[lrnde watchdog] mbarrier wait timed out: block %d thread %d smem 0x%x parity %u\n", (int)blockIdx.x,
             (int)threadIdx.x, addr, parity);
      asm volatile("trap;");
    }
#endif
  } while (!ok);
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// multicast variants (thread-block cluster): the copy lands at the same shared-memory offset of
// every CTA in ctaMask and signals the mbarrier at the same offset there
__device__ __forceinline__ void bulk_g2s_mc(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar,
                                            uint16_t mask) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   start address >> 4 | LBO (ignored for swizzled K-major; 1) | SBO = 1024 B (8 rows x 128 B)
//   | version 1 (Blackwell) | layout type 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::tf32 instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32,
// both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// one lane of a converged warp (the pattern the compiler maps to ELECT + uniform-register operands;
// a divergent `if (lane == 0)` region makes every tcgen05 operand go through an R2UR waterfall loop)
__device__ __forceinline__ uint32_t elect_one_sync() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, %1;\n\t"
      "@px mov.s32 %0, 1;\n\t}"
      : "+r"(pred)
      : "r"(0xFFFFFFFFu));
  return pred;
}
// descriptor = {lo: start address >> 4 | LBO << 16, hi: SBO | version | layout}
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); }
// COLL: A-operand collector usage: 0 = discard, 1 = fill (keep A for the next MMA), 2 = lastuse
// (take A from the collector instead of shared memory).  In 3xTF32 the two MMAs that share
// A_hi are issued back to back as fill / lastuse: one 4 KB shared-memory read less per K-step
// (the kernels are shared-memory-bandwidth bound at N = 64).
template <int COLL = 0>
__device__ __forceinline__ void mma_tf32_lo(uint32_t d_tmem, uint32_t a_lo32, uint32_t b_lo32, uint32_t idesc,
                                            uint32_t accumulate) {
  if (COLL == 1) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32.collector::a::fill [%0], da, db, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo32), "r"(b_lo32), "r"(idesc), "r"(accumulate), "r"(kDescHi)
        : "memory");
  } else if (COLL == 2) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32.collector::a::lastuse [%0], da, db, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo32), "r"(b_lo32), "r"(idesc), "r"(accumulate), "r"(kDescHi)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "mov.b64 da, {%1, %5};\n\t"
        "mov.b64 db, {%2, %5};\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_lo32), "r"(b_lo32), "r"(idesc), "r"(accumulate), "r"(kDescHi)
        : "memory");
  }
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
      ::"r"(smem_u32(bar)), "h"(mask)
      : "memory");
}
// round to nearest tf32, ties away from zero -- the result of cvt.rna.tf32.f32 for every finite value, as two
// full-rate integer instructions: the cvt form issues at a few results per clock per SM and was the largest single
// cost of every kernel that splits operands into hi / lo parts on the fly (in-kernel clock64 timeline, profiles/r2)
__device__ __forceinline__ float tf32_rna(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
// ---- MN-major operands (both GEMM operands stored feature-contiguous per sample, K = samples): for tf32 the
// only MN-major shared-memory layout is SWIZZLE_128B_BASE32B (cute Layout_MN_SW128_32B_Atom): atoms of 4 samples x
// 128 bytes (32 rows), 32-byte chunks XOR-ed with the sample index; a_major = b_major = MN in the instruction descriptor
__host__ __device__ constexpr uint32_t make_idesc_mn(int M, int N) {
  return make_idesc(M, N) | (1u << 15) | (1u << 16);
}
// MN-major descriptor: atom (g, k4) of 512 B at ((k4 * 4 + g) * 512): LBO = 512 B between 32-row
// groups, SBO = 2048 B between 4-sample groups, layout type 1 = SWIZZLE_128B_BASE32B
constexpr uint32_t kDescHiMN = (2048u >> 4) | (1u << 14) | (1u << 29);
__device__ __forceinline__ uint32_t desc_lo_mn(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | ((512u >> 4) << 16); }
template <int COLL = 0>
__device__ __forceinline__ void mma_tf32_mn(uint32_t d_tmem, uint32_t a_lo32, uint32_t b_lo32, uint32_t idesc,
                                            uint32_t accumulate) {
#define LR_MMA_MN(QUAL)                                                                              \
  asm volatile(                                                                                      \
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"                                                \
      "setp.ne.b32 p, %4, 0;\n\t"                                                                    \
      "mov.b64 da, {%1, %5};\n\t"                                                                    \
      "mov.b64 db, {%2, %5};\n\t"                                                                    \
      "tcgen05.mma.cta_group::1.kind::tf32" QUAL " [%0], da, db, %3, p;\n\t}"                         \
      ::"r"(d_tmem), "r"(a_lo32), "r"(b_lo32), "r"(idesc), "r"(accumulate), "r"(kDescHiMN)           \
      : "memory")
  if (COLL == 1) LR_MMA_MN(".collector::a::fill");
  else if (COLL == 2) LR_MMA_MN(".collector::a::lastuse");
  else LR_MMA_MN("");
#undef LR_MMA_MN
}

// byte offset of (row, 16-byte chunk c) inside a K-major SWIZZLE_128B tile of 128-byte rows
__device__ __forceinline__ uint32_t swz_off(int row, int c) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((c ^ (row & 7)) << 4));
}

}  // namespace umma
