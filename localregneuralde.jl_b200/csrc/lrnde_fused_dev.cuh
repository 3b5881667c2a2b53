// Device helpers shared by the latent-space engines (lrnde_fused.cu: forward attempts, lrnde_adjoint.cu: adjoint
// attempts): tcgen05.mma wrappers (A from shared memory or from tensor memory), TMEM loads / stores, bulk stores,
// operand-image offsets.  sm_100a only.
#pragma once
#include "lrnde_act.cuh"
#include "lrnde_tc.cuh"

namespace fused {
using namespace umma;

constexpr int kNT = 64;                         // samples per tile
constexpr int kEpiWarps = 16;                   // warps 2..17: compute / epilogue (four per TMEM lane quarter)
constexpr int kThreads = 64 + 32 * kEpiWarps;   // warp 0: bulk copies, warp 1: tcgen05.mma issue
constexpr int kUR = 96;                         // rows of a kgemm operand unit: 6 stages x 16 samples
constexpr int kPieceBytes = 2 * kUR * 128;      // kgemm ring stage: [hi | lo] of a 96-row x 32-float chunk
constexpr int kTailBytes = 2 * kUR * 32;        // [hi | lo] of a 96-row x 8-float K-step

// shared-memory matrix descriptors (cute::UMMA::SmemDescriptor), K-major:
//   SWIZZLE_128B: rows of 128 B, 8-row atoms of 1024 B (SBO), 16-byte chunks XOR-ed with (row & 7)
//   SWIZZLE_32B : rows of  32 B, 8-row atoms of  256 B (SBO), 16-byte halves XOR-ed with ((row >> 2) & 1)
constexpr uint32_t kHi128 = (1024u >> 4) | (1u << 14) | (2u << 29);
constexpr uint32_t kHi32 = (256u >> 4) | (1u << 14) | (6u << 29);

template <int COLL>
__device__ __forceinline__ void mma(uint32_t d_tmem, uint32_t a_lo32, uint32_t b_lo32, uint32_t hi32, uint32_t idesc,
                                    uint32_t accumulate) {
#define LR_FMMA(QUAL)                                                                                \
  asm volatile(                                                                                      \
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"                                                \
      "setp.ne.b32 p, %4, 0;\n\t"                                                                    \
      "mov.b64 da, {%1, %5};\n\t"                                                                    \
      "mov.b64 db, {%2, %5};\n\t"                                                                    \
      "tcgen05.mma.cta_group::1.kind::tf32" QUAL " [%0], da, db, %3, p;\n\t}"                         \
      ::"r"(d_tmem), "r"(a_lo32), "r"(b_lo32), "r"(idesc), "r"(accumulate), "r"(hi32)                \
      : "memory")
  if (COLL == 1) LR_FMMA(".collector::a::fill");
  else if (COLL == 2) LR_FMMA(".collector::a::lastuse");
  else LR_FMMA("");
#undef LR_FMMA
}

// one K-step (8 tf32) of D += A B^T in 3xTF32 (A_lo B_hi + A_hi B_lo + A_hi B_hi) or plain TF32
__device__ __forceinline__ void mma_kstep(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                          uint32_t hi32, uint32_t idesc, uint32_t acc_first, int passes) {
  if (passes == 3) {
    mma<0>(d, a_lo, b_hi, hi32, idesc, acc_first);
    mma<1>(d, a_hi, b_lo, hi32, idesc, 1u);
    mma<2>(d, a_hi, b_hi, hi32, idesc, 1u);
  } else {
    mma<0>(d, a_hi, b_hi, hi32, idesc, acc_first);
  }
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, "
      "%13, %14, %15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, "
      "%14, %15, %16};"
      ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
        "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
        "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
        "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])),
        "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void bulk_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}

// byte offset of element (row r, column k) inside the hi image of an R-row operand: nfull SWIZZLE_128B chunks of
// 32 floats followed by SWIZZLE_32B K-steps of 8 floats
__host__ __device__ inline uint32_t img_off(int R, int nfull, int r, int k) {
  if (k < nfull * 32) {
    const int c = k >> 5, kk = k & 31;
    return (uint32_t)(c * R * 128 + (r >> 3) * 1024 + (r & 7) * 128 + (((kk >> 2) ^ (r & 7)) << 4) + (kk & 3) * 4);
  }
  const int s = (k - nfull * 32) >> 3, kk = k & 7;
  return (uint32_t)(nfull * R * 128 + s * R * 32 + r * 32 + ((((kk >> 2) & 1) ^ ((r >> 2) & 1)) << 4) + (kk & 3) * 4);
}

__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(taddr)
      : "memory");
#pragma unroll
  for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[j]);
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
               "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T : A = 128 lanes x 8 columns (tf32 in 32-bit cells)
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo32, uint32_t hi32, uint32_t idesc,
                                       uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], db, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo32), "r"(idesc), "r"(accumulate), "r"(hi32)
      : "memory");
}

constexpr int kDCol = 256;      // accumulator buffers at TMEM columns 256 and 384 (A operand in columns [0, 2 * 8 * KS))

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] = __uint_as_float(r[j]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

}  // namespace fused
