// tcgen05 / TMEM dense layer for sm_100a: the GEMM of one Lux Dense (or of its transposed
// data-gradient) in TF32 or 3xTF32, drop-in for dense_nn_kernel (same DenseP contract).
//
//   D[128 x NT] (TMEM, fp32)  +=  A[128 x 8] (smem, K-major, SWIZZLE_128B)  x  B[NT x 8]^T
//
//   A = weight tile (output features on the 128 TMEM lanes).  Weights are constant during a
//       solve, so they are split once per call into tf32 hi / lo parts and stored in global
//       memory as ready-made shared-memory images (swizzle applied); a single elected thread
//       streams them with cp.async.bulk (TMA engine) onto an mbarrier ring.
//   B = activations [NT samples x 32 features] per K-chunk.  The stage combination
//       uprev + dt * sum a_ij k_j (or the dense-output interpolant, or the TDChain time row /
//       bias row) is formed by the producer warps WHILE the tile is written to shared memory
//       -- the "linear combination in the operand prologue" of BASELINE.json's north star --
//       and split hi / lo on the fly.
//   3xTF32: D += A_lo*B_hi + A_hi*B_lo + A_hi*B_hi  (three tcgen05.mma per K-step).
//   Epilogue: tcgen05.ld -> activation / activation-derivative -> coalesced global stores
//       (lane = output feature => 32 consecutive floats per sample per warp).
//
// Two schedules:
//   RING      one 128-row weight tile per CTA (grid.y = M tiles); A and B chunks share the
//             mbarrier ring.  Used when K is long (layer 1: K = 786).
//   RESIDENT  all K-chunks of the activation tile stay in shared memory, the CTA loops over
//             every weight tile (up to 512 TMEM columns).  Used when K is short (layer 2:
//             K = 102, 7 weight tiles).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "lrnde_kernels.cuh"

#ifdef LRNDE_UMMA_TRACE
__device__ long long g_umma_trace[8192];
#define UMMA_TRACE(slot, it) do { if (blockIdx.x == 1 && blockIdx.y == 0 && (it) < 64) g_umma_trace[(RESIDENT ? 512 : 0) + (slot) * 64 + (it)] = clock64(); } while (0)
#define WG_TRACE(slot, it) do { if (blockIdx.x == 1 && blockIdx.y == 1 && (it) < 64) g_umma_trace[1024 + (slot) * 64 + (it)] = clock64(); } while (0)
#else
#define UMMA_TRACE(slot, it) do { } while (0)
#define WG_TRACE(slot, it) do { } while (0)
#endif

#include "lrnde_tc.cuh"

namespace umma {

constexpr int kChunkK = 32;                 // floats per K-chunk = one 128-byte swizzle row
constexpr int kAChunkFloats = 128 * kChunkK;  // 4096 floats = 16 KB (hi), same again for lo
constexpr int kAChunkBytes = 2 * kAChunkFloats * 4;  // 32 KB: [hi | lo]
constexpr int kProducerWarps = 16;          // many warps: the operand prologue is load-latency bound
constexpr int kThreads = 64 + kProducerWarps * 32;  // warp 0: bulk-copy producer, warp 1: MMA
                                                    // issuer, the rest: operand producers, then
                                                    // the epilogue


// ---- weights -> ready-made shared-memory images:  out[(mt*KC + kc)] = [hi 16 KB | lo 16 KB]
__global__ void pack_weights_kernel(const float* __restrict__ A, long lda, int M, int Kaug, int KC,
                                    float* __restrict__ out, int passes) {
  // blockIdx.y = replica: every CTA of a GEMM reads the same weight images at the same time, so
  // the images are replicated (and the chunk order rotated per CTA) to spread the L2 slices
  const int chunk = blockIdx.x;
  const int mt = chunk / KC, kc = chunk % KC;
  float* dst = out + ((size_t)blockIdx.y * gridDim.x + chunk) * (2 * kAChunkFloats);
  for (int e = threadIdx.x; e < kAChunkFloats; e += blockDim.x) {
    const int r = e & 127, kk = e >> 7;
    const int m = mt * 128 + r, k = kc * kChunkK + kk;
    float v = (m < M && k < Kaug) ? A[(size_t)k * lda + m] : 0.0f;
    float hi = (passes == 1) ? v : tf32_rna(v);
    float lo = (passes == 1) ? 0.0f : tf32_rna(v - hi);  // rounded, not left to the MMA's truncation
    const int c = kk >> 2;
    const int pos = (r >> 3) * 256 + (r & 7) * 32 + ((c ^ (r & 7)) << 2) + (kk & 3);
    dst[pos] = hi;
    dst[kAChunkFloats + pos] = lo;
  }
}

struct UmmaP {
  DenseP d;
  const float* Apack;  // packed weight images
  int n_mt;            // number of 128-row weight tiles
  int KC;              // number of 32-float K chunks (covers K + td + bias)
  int passes;          // 3 = 3xTF32, 1 = TF32
  int replicas;        // copies of the packed images (stride n_mt * KC chunks)
  int nacc_max;        // RING: upper bound on the number of split TMEM accumulators
  int res_group;       // RESIDENT: weight tiles per CTA (grid.y = groups of tiles)
};
constexpr int kReplicas = 4;
constexpr int kBulkParts = 2;
constexpr int kPrefetchAhead = 4;  // K-chunks of L2 prefetch distance for the activation operand

template <int NT>
struct Cfg {
  static constexpr int kBChunkBytes = NT * 128 * 2;  // [hi | lo]
  static constexpr int kRingStage = kAChunkBytes + kBChunkBytes;
  static constexpr int kRingStages = (NT == 64) ? 4 : 3;
  static constexpr int kResStages = (NT == 64) ? 5 : 3;
  static constexpr int kMaxResKC = 4;
  static constexpr int ring_smem() { return kRingStages * kRingStage + 1024; }
  static constexpr int res_smem(int KC) { return kResStages * kAChunkBytes + KC * kBChunkBytes + 1024; }
};

// One float4 group (row, 16-byte chunk c) of the activation operand: loads are issued by
// load_group() and consumed later by store_group(), so that several K-chunks of loads are in
// flight per thread (the prologue is load-latency bound: ~1 us per dependent global load).
struct ProdCtx {
  const DenseP* p;
  const LinComb* sd;
  int n0, nsrc, passes;
  bool vec;
  float tval;
  float* side;  // optional fp32 copy of the combined tile (only the blockIdx.y == 0 CTAs write it)
};

// GEN = false is the lean instantiation for the common case (K % 4 == 0, 16-byte aligned operands, no
// input activation): no scalar fallback in the hot loop -- the TDChain time row and the bias row are the
// first elements of the float4 group that starts at k == K.  The host picks it (MlpEval::dense).
template <int NS1, bool GEN>  // NS1 = 1 + number of lincomb sources held in registers
__device__ __forceinline__ void load_group(const ProdCtx& c, int row, int cch, int kc, float4 (&buf)[NS1]) {
  const DenseP& p = *c.p;
  const int n = c.n0 + row, k = kc * kChunkK + cch * 4;
  if (n < p.N && (GEN ? (c.vec && k + 3 < p.K) : (k < p.K))) {
    const size_t off = (size_t)n * p.ldx + k;
    // streaming loads (L2 only): every byte is used exactly once by this CTA, and the L1 data array is
    // the shared-memory array the MMA operands and the bulk copies already saturate
    buf[0] = c.sd->base ? __ldcg(reinterpret_cast<const float4*>(c.sd->base + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int s = 0; s < NS1 - 1; ++s)
      if (s < c.nsrc) buf[s + 1] = __ldcg(reinterpret_cast<const float4*>(c.sd->src[s] + off));
  }
}

template <int NS1, int NT, bool GEN>
__device__ __forceinline__ void store_group(const ProdCtx& c, int row, int cch, int kc, const float4 (&buf)[NS1],
                                            uint8_t* dst) {
  const DenseP& p = *c.p;
  const LinComb& sd = *c.sd;
  const int n = c.n0 + row, k = kc * kChunkK + cch * 4;
  float v[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  if (!GEN) {
    if (n < p.N) {
      if (k < p.K) {
        float4 x = buf[0];
        if (NS1 > 1 && c.nsrc) {
          float4 inner = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int s = 0; s < NS1 - 1; ++s)
            if (s < c.nsrc) {
              const float cf = sd.coef[s];
              inner.x = fmaf(cf, buf[s + 1].x, inner.x); inner.y = fmaf(cf, buf[s + 1].y, inner.y);
              inner.z = fmaf(cf, buf[s + 1].z, inner.z); inner.w = fmaf(cf, buf[s + 1].w, inner.w);
            }
          x.x = fmaf(sd.scale, inner.x, x.x); x.y = fmaf(sd.scale, inner.y, x.y);
          x.z = fmaf(sd.scale, inner.z, x.z); x.w = fmaf(sd.scale, inner.w, x.w);
        }
        if (c.side) *reinterpret_cast<float4*>(c.side + (size_t)n * p.ldx + k) = x;
        v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
      } else if (k == p.K) {
        if (p.td) { v[0] = c.tval; if (p.bias) v[1] = 1.0f; }
        else if (p.bias) v[0] = 1.0f;
      }
    }
  } else if (n < p.N) {
    const size_t off = (size_t)n * p.ldx + k;
    if (c.vec && k + 3 < p.K) {
      float4 inner = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int s = 0; s < NS1 - 1; ++s)
        if (s < c.nsrc) {
          const float cf = sd.coef[s];
          inner.x = fmaf(cf, buf[s + 1].x, inner.x); inner.y = fmaf(cf, buf[s + 1].y, inner.y);
          inner.z = fmaf(cf, buf[s + 1].z, inner.z); inner.w = fmaf(cf, buf[s + 1].w, inner.w);
        }
      if (c.nsrc) {
        v[0] = fmaf(sd.scale, inner.x, buf[0].x); v[1] = fmaf(sd.scale, inner.y, buf[0].y);
        v[2] = fmaf(sd.scale, inner.z, buf[0].z); v[3] = fmaf(sd.scale, inner.w, buf[0].w);
      } else { v[0] = buf[0].x; v[1] = buf[0].y; v[2] = buf[0].z; v[3] = buf[0].w; }
      if (c.side) *reinterpret_cast<float4*>(c.side + off) = make_float4(v[0], v[1], v[2], v[3]);
      if (p.in_act) {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = lr_act(p.in_act, v[e]);
      }
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int kk = k + e;
        if (kk < p.K) {
          float x = lr_lincomb_at(sd, off + e);
          if (c.side) c.side[off + e] = x;
          v[e] = p.in_act ? lr_act(p.in_act, x) : x;
        } else if (p.td && kk == p.K) v[e] = c.tval;
        else if (p.bias && kk == p.K + p.td) v[e] = 1.0f;
      }
    }
  }
  float4 hi, lo;
  if (c.passes == 3) {
    hi = make_float4(tf32_rna(v[0]), tf32_rna(v[1]), tf32_rna(v[2]), tf32_rna(v[3]));
    lo = make_float4(tf32_rna(v[0] - hi.x), tf32_rna(v[1] - hi.y), tf32_rna(v[2] - hi.z), tf32_rna(v[3] - hi.w));
  } else {
    hi = make_float4(v[0], v[1], v[2], v[3]);
    lo = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const uint32_t o = swz_off(row, cch);
  *reinterpret_cast<float4*>(dst + o) = hi;
  *reinterpret_cast<float4*>(dst + NT * 128 + o) = lo;
}

// CL = thread-block cluster size along the sample tiles (1 or 4): the CTAs of a cluster read the
// same weight chunks, so each loads a 1/CL slice and multicasts it to all of them (L2 -> SM
// traffic of the weight images / CL); stages are released cluster-wide by multicast commits.
template <int NT, bool RESIDENT, int CL, bool GEN>
__global__ void __launch_bounds__(kThreads, 1) dense_kernel(UmmaP q) {
  const DenseP& p = q.d;
  if (p.done && *p.done) return;
  const uint32_t crank = (CL > 1) ? cluster_ctarank() : 0u;
  constexpr uint16_t kMask = (uint16_t)((1u << CL) - 1u);
  using C = Cfg<NT>;
  constexpr int NS = RESIDENT ? C::kResStages : C::kRingStages;
  constexpr int PT = kProducerWarps * 32;  // producer / epilogue threads
  constexpr int GPT = NT * 8 / PT;         // float4 groups per producer thread per K-chunk
  static_assert(NT * 8 % PT == 0 && GPT >= 1, "producer mapping");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[8], empty_bar[8], bempty_bar[8], tile_bar[8], bres_bar;
  __shared__ uint32_t tmem_slot;
  __shared__ LinComb sdesc;
  __shared__ float s_t;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * NT;
  const int KC = q.KC;
  const int mt_begin = RESIDENT ? (int)blockIdx.y * q.res_group : (int)blockIdx.y;
  const int mt_count = RESIDENT ? min(q.res_group, q.n_mt - mt_begin) : 1;
  const int nchunks = mt_count * KC;
  // RING: the K loop is split over several TMEM accumulators that the epilogue adds in fp32
  // (the tensor core accumulates with truncation; shorter chains keep 3xTF32 at fp32 level)
  const int nacc = RESIDENT ? 1 : min(min(512 / NT, KC), q.nacc_max);
  const int ncols_used = RESIDENT ? mt_count * NT : nacc * NT;
  // chunk order rotated per CTA (accumulation order is free): de-synchronises the CTAs.
  // RESIDENT rotates whole weight tiles so that tiles finish one after the other and the
  // epilogue of a finished tile overlaps the MMAs of the next.
  const unsigned cid = blockIdx.x / CL;  // identical order inside a cluster
  const int rot = RESIDENT ? (int)((cid * 3u) % (unsigned)mt_count) * KC
                           : (int)((cid * 7u + blockIdx.y * 3u) % (unsigned)nchunks);

  auto a_stage = [&](int s) -> uint8_t* {
    return RESIDENT ? smem + (size_t)s * kAChunkBytes : smem + (size_t)s * C::kRingStage;
  };
  auto b_slot = [&](int s_or_kc) -> uint8_t* {
    return RESIDENT ? smem + (size_t)NS * kAChunkBytes + (size_t)s_or_kc * C::kBChunkBytes
                    : smem + (size_t)s_or_kc * C::kRingStage + kAChunkBytes;
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full_bar[s], RESIDENT ? 1u : (uint32_t)(kProducerWarps + 1));  // one arrival per warp
      mbar_init(&empty_bar[s], (uint32_t)CL);  // weight stage free in every CTA of the cluster
      mbar_init(&bempty_bar[s], 1u);           // activation stage (local)
    }
    for (int t = 0; t < 8; ++t) mbar_init(&tile_bar[t], 1u);
    mbar_init(&bres_bar, (uint32_t)kProducerWarps);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (p.xdesc) sdesc = *p.xdesc;
    else { sdesc.base = p.X; sdesc.n = 0; sdesc.scale = 0.0f; }
    s_t = p.tdesc ? p.tdesc->t : 0.0f;
  }
  uint32_t ncols = 32;
  while (ncols < (uint32_t)ncols_used) ncols <<= 1;
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // every CTA's barriers are initialised before any remote arrive
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (threadIdx.x == 0) UMMA_TRACE(7, 0);

  if (warp == 0) {
    // ------------------------------------------------ weight images: bulk copies onto the ring
    // (converged warp + one elected lane, same reason as the MMA warp)
    const float* src = q.Apack + (size_t)(cid % q.replicas) * q.n_mt * KC * (2 * kAChunkFloats);
    for (int it = 0; it < nchunks; ++it) {
      const int s = it % NS, ph = (it / NS) & 1;
      const int j = (it + rot) % nchunks;
      const int mt = mt_begin + j / KC, kc = j % KC;
      mbar_wait(&empty_bar[s], ph ^ 1);
      const uint8_t* g = (const uint8_t*)(src + (size_t)(mt * KC + kc) * (2 * kAChunkFloats));
      uint8_t* d = a_stage(s);
      if (elect_one_sync()) {
        UMMA_TRACE(0, it);
        mbar_arrive_expect_tx(&full_bar[s], kAChunkBytes);
        if (CL > 1) {
          constexpr uint32_t slice = kAChunkBytes / CL;
          bulk_g2s_mc(d + crank * slice, g + crank * slice, slice, &full_bar[s], kMask);
        } else {
#pragma unroll
          for (int part = 0; part < kBulkParts; ++part)
            bulk_g2s(d + part * (kAChunkBytes / kBulkParts), g + part * (kAChunkBytes / kBulkParts),
                     kAChunkBytes / kBulkParts, &full_bar[s]);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------ tcgen05.mma issue: the whole warp runs the
    // loop converged (uniform registers), one elected lane issues the MMAs and the commits
    constexpr uint32_t idesc = make_idesc(128, NT);
    if (RESIDENT) mbar_wait(&bres_bar, 0);
    for (int it = 0; it < nchunks; ++it) {
      const int s = it % NS, ph = (it / NS) & 1;
      const int j = (it + rot) % nchunks;
      const int mi = j / KC, kc = j % KC;
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      if (lane == 0) UMMA_TRACE(1, it);
      const uint32_t a_hi = desc_lo(smem_u32(a_stage(s)));
      const uint32_t a_lo = desc_lo(smem_u32(a_stage(s)) + kAChunkFloats * 4);
      const uint32_t b_hi = desc_lo(smem_u32(b_slot(RESIDENT ? kc : s)));
      const uint32_t b_lo = desc_lo(smem_u32(b_slot(RESIDENT ? kc : s)) + NT * 128);
      int acc = 0;
      bool group_start, group_end;
      if (!RESIDENT) {
        acc = (it * nacc) / KC;
        group_start = (it == 0) || (((it - 1) * nacc) / KC != acc);
        group_end = (it == nchunks - 1);
      } else {
        group_start = (kc == 0);      // whole tiles are rotated: a tile's chunks are contiguous
        group_end = (kc == KC - 1);
      }
      const uint32_t d = tmem_base + (uint32_t)((RESIDENT ? mi : acc) * NT);
      if (elect_one_sync()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t ko = (uint32_t)k * 2;  // 8 tf32 = 32 bytes = 2 x 16 B along K in the swizzle row
          const uint32_t first = (group_start && k == 0) ? 0u : 1u;
          if (q.passes == 3) {
            mma_tf32_lo<0>(d, a_lo + ko, b_hi + ko, idesc, first);
            mma_tf32_lo<1>(d, a_hi + ko, b_lo + ko, idesc, 1u);
            mma_tf32_lo<2>(d, a_hi + ko, b_hi + ko, idesc, 1u);
          } else {
            mma_tf32_lo(d, a_hi + ko, b_hi + ko, idesc, first);
          }
        }
        // frees the stage once these MMAs have read it: weights cluster-wide, activations locally
        if (CL > 1) mma_commit_mc(&empty_bar[s], kMask);
        else mma_commit(&empty_bar[s]);
        if (!RESIDENT) mma_commit(&bempty_bar[s]);
        if (group_end) mma_commit(&tile_bar[RESIDENT ? mi : 0]);  // accumulator(s) complete
        UMMA_TRACE(2, it);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------ warps 2..: operand producers
    const int tid = threadIdx.x - 64;
    ProdCtx pc;
    pc.p = &p; pc.sd = &sdesc; pc.n0 = n0; pc.nsrc = sdesc.n; pc.passes = q.passes; pc.tval = s_t;
    pc.side = (blockIdx.y == 0) ? (p.side_desc ? p.side_desc->dst : p.side) : nullptr;
    pc.vec = (p.ldx % 4 == 0) && ((((uintptr_t)sdesc.base) & 15) == 0) && ((((uintptr_t)pc.side) & 15) == 0);
    for (int k = 0; k < sdesc.n; ++k) pc.vec = pc.vec && ((((uintptr_t)sdesc.src[k]) & 15) == 0);
    if (!GEN && !pc.vec) asm volatile("trap;");   // the host selected the lean kernel for a misaligned operand

    // software-pipelined producer: DEPTH chunks of loads in flight per thread, about 8 float4
    // registers of load buffers whatever the number of lincomb sources
    auto run = [&](auto ns1_tag, auto depth_tag) {
      constexpr int NS1 = decltype(ns1_tag)::value;
      constexpr int DEPTH = decltype(depth_tag)::value;
      float4 buf[DEPTH][GPT][NS1];
      const int total = RESIDENT ? KC : nchunks;
      auto chunk_of = [&](int it) { return RESIDENT ? it : (it + rot) % KC; };
#pragma unroll
      for (int d = 0; d < DEPTH; ++d)
        if (d < total) {
#pragma unroll
          for (int g = 0; g < GPT; ++g) {
            const int gi = tid + g * PT;
            load_group<NS1, GEN>(pc, gi >> 3, gi & 7, chunk_of(d), buf[d][g]);
          }
        }
      for (int it0 = 0; it0 < total; it0 += DEPTH) {
#pragma unroll
        for (int d = 0; d < DEPTH; ++d) {
          const int it = it0 + d;
          if (it < total) {
            const int s = it % NS, ph = (it / NS) & 1;
            uint8_t* dst = RESIDENT ? b_slot(it) : b_slot(s);
            if (!RESIDENT) mbar_wait(&bempty_bar[s], ph ^ 1);
            if (tid == 0) UMMA_TRACE(3, it);
#pragma unroll
            for (int g = 0; g < GPT; ++g) {
              const int gi = tid + g * PT;
              store_group<NS1, NT, GEN>(pc, gi >> 3, gi & 7, chunk_of(it), buf[d][g], dst);
            }
            if (tid == 0) UMMA_TRACE(4, it);
            if (it + DEPTH < total) {
#pragma unroll
              for (int g = 0; g < GPT; ++g) {
                const int gi = tid + g * PT;
                load_group<NS1, GEN>(pc, gi >> 3, gi & 7, chunk_of(it + DEPTH), buf[d][g]);
              }
            }
            if (!RESIDENT) {
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) mbar_arrive(&full_bar[s]);  // 512 arrivals on one word would serialise
            }
          }
        }
      }
      if (RESIDENT) {
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bres_bar);
      }
    };
    using I1 = std::integral_constant<int, 1>;
    using I2 = std::integral_constant<int, 2>;
    using I3 = std::integral_constant<int, 3>;
    using I4 = std::integral_constant<int, 4>;
    using I8 = std::integral_constant<int, 8>;
    if (pc.nsrc == 0) run(I1{}, I4{});
    else if (pc.nsrc == 1) run(I2{}, I4{});
    else if (pc.nsrc <= 3) run(I4{}, I2{});
    else run(I8{}, I2{});

    // ------------------------------------------------ epilogue: TMEM -> registers -> global
    // warps sharing a TMEM lane quarter split the 16-column blocks; tiles are drained in the
    // order the MMA warp finishes them
    const int quarter = (warp & 3) * 32;     // the TMEM lanes this warp may read
    const int wgroup = (warp - 2) >> 2;
    constexpr int kGroups = kProducerWarps / 4;
    const int row = quarter + lane;
    float* Y = p.ydesc ? (p.ydesc->dst + p.y_off) : p.Y;
    for (int t = 0; t < mt_count; ++t) {
      const int mi = RESIDENT ? (t + rot / KC) % mt_count : 0;
      if (tid == 0) UMMA_TRACE(6, 0);
      mbar_wait(&tile_bar[mi], 0);
      tc_fence_after();
      if (tid == 0) UMMA_TRACE(6, 1);
      const int m = (mt_begin + mi) * 128 + row;
      for (int cb = wgroup; cb < NT / 16; cb += kGroups) {
        const int c0 = cb * 16;
        float accv[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) accv[j] = 0.0f;
        for (int a = 0; a < nacc; ++a) {
          uint32_t r[16];
          const uint32_t taddr =
              tmem_base + ((uint32_t)quarter << 16) + (uint32_t)((RESIDENT ? mi : a) * NT + c0);
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, "
              "%13, %14, %15}, [%16];\n\t"
              "tcgen05.wait::ld.sync.aligned;"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
                "=r"(r[15])
              : "r"(taddr)
              : "memory");
#pragma unroll
          for (int j = 0; j < 16; ++j) accv[j] += __uint_as_float(r[j]);
        }
        if (m < p.M) {
          if (GEN) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int n = n0 + c0 + j;
              if (n < p.N) {
                float v = accv[j];
                if (p.pre) p.pre[(size_t)n * p.ldpre + m] = v;
                if (p.dact >= 0) v = v * lr_dact(p.dact, p.dpre[(size_t)n * p.lddpre + m]);
                else v = lr_act(p.act, v);
                Y[(size_t)n * p.ldy + m] = v * p.out_scale;
              }
            }
          } else {
            // lean: act in {identity, tanh}, dact in {none, identity, tanh}; the mode is chosen once
            const int mode = (p.dact == ACT_TANH) ? 2 : ((p.dact < 0 && p.act == ACT_TANH) ? 1 : 0);
            const float os = p.out_scale;
            if (mode == 0) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int n = n0 + c0 + j;
                if (n < p.N) {
                  if (p.pre) p.pre[(size_t)n * p.ldpre + m] = accv[j];
                  Y[(size_t)n * p.ldy + m] = accv[j] * os;
                }
              }
            } else if (mode == 1) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int n = n0 + c0 + j;
                if (n < p.N) {
                  if (p.pre) p.pre[(size_t)n * p.ldpre + m] = accv[j];
                  Y[(size_t)n * p.ldy + m] = tanhf(accv[j]) * os;
                }
              }
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int n = n0 + c0 + j;
                if (n < p.N) {
                  if (p.pre) p.pre[(size_t)n * p.ldpre + m] = accv[j];
                  const float y = tanhf(p.dpre[(size_t)n * p.lddpre + m]);
                  Y[(size_t)n * p.ldy + m] = accv[j] * (1.0f - y * y) * os;
                }
              }
            }
          }
        }
      }
    }
    tc_fence_before();
    if (tid == 0) UMMA_TRACE(6, 2);
  }
  __syncthreads();
  if (CL > 1) cluster_sync_all();  // nobody exits while peers may still signal its barriers
  if (threadIdx.x == 0) UMMA_TRACE(6, 3);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// Weight gradient on the tensor cores:  part[z][..] = sum_{b in split z} P[:, b] * Q[:, b]^T
// Both operands are activations stored feature-contiguous per sample ([rows, B] column-major),
// i.e. MN-major for the MMA.  For tf32 the only MN-major shared-memory layout is
// SWIZZLE_128B_BASE32B (cute Layout_MN_SW128_32B_Atom): atoms of 4 samples x 128 bytes (32 rows),
// 32-byte chunks XOR-ed with the sample index (Swizzle<2,5,2>); LBO = stride between 32-row
// groups, SBO = stride between 4-sample groups; the instruction descriptor carries
// a_major = b_major = MN.  P = the operand with more rows
// (128-row tiles on the TMEM lanes), Q = the other one (<= 128 rows).  One operand may be the
// augmented layer input [x ; t ; 1] (TDChain time row, bias row) with an input activation.
// ------------------------------------------------------------------------------------------
struct WgOperand {
  const float* ptr; long ld; int rows;  // rows read from memory
  int td, bias, in_act;                 // extra generated rows (augmented input only)
};
struct WgradUP {
  WgOperand P, Q;
  int B, chunk;      // batch size, samples per split (multiple of 32)
  int transposed;    // 0: out[m + n * ldo] ; 1: out[n + m * ldo]
  int ldo;           // leading dimension of the [out x (in+2)] block
  float* part;       // [S][block] partial sums
  size_t block;      // floats per partial block
  int passes;
  const LinComb* tdesc;
  const int* done;
};
constexpr int kWgStageBytes = 4 * 16384;  // [P_hi | P_lo | Q_hi | Q_lo]
constexpr int kWgStages = 3;
constexpr int wgrad_smem() { return kWgStages * kWgStageBytes + 1024; }
// GEN = false: lean instantiation (rows % 4 == 0, aligned, no input activation): the time / bias rows are the
// first elements of the float4 group starting at row == rows (same idea as the lean dense kernel)
template <bool GEN>
__device__ __forceinline__ float4 wg_load(const WgOperand& o, int row, int b, bool valid, float tval, bool vec) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (!valid) return v;
  const float* p = o.ptr + (size_t)b * o.ld + row;
  if (!GEN) {
    if (row < o.rows) v = __ldcg(reinterpret_cast<const float4*>(p));
    else if (row == o.rows) {
      if (o.td) { v.x = tval; if (o.bias) v.y = 1.0f; }
      else if (o.bias) v.x = 1.0f;
    }
    return v;
  }
  if (vec && row + 3 < o.rows) {
    v = *reinterpret_cast<const float4*>(p);
    if (o.in_act) { v.x = lr_act(o.in_act, v.x); v.y = lr_act(o.in_act, v.y); v.z = lr_act(o.in_act, v.z); v.w = lr_act(o.in_act, v.w); }
  } else {
    float e[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = row + j;
      float x = 0.0f;
      if (r < o.rows) { x = p[j]; if (o.in_act) x = lr_act(o.in_act, x); }
      else if (o.td && r == o.rows) x = tval;
      else if (o.bias && r == o.rows + o.td) x = 1.0f;
      e[j] = x;
    }
    v = make_float4(e[0], e[1], e[2], e[3]);
  }
  return v;
}

template <bool GEN>
__global__ void __launch_bounds__(kThreads, 1) wgrad_kernel(WgradUP q) {
  if (q.done && *q.done) return;
  constexpr int NS = kWgStages;
  constexpr int PT = kProducerWarps * 32;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full_bar[NS], empty_bar[NS], done_bar;
  __shared__ uint32_t tmem_slot;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * 128;
  const int b_begin = blockIdx.y * q.chunk;
  const int b_end = min(q.B, b_begin + q.chunk);
  const int nchunks = (b_end - b_begin + 31) / 32;
  const float tval = q.tdesc ? q.tdesc->t : 0.0f;

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full_bar[s], (uint32_t)kProducerWarps);
      mbar_init(&empty_bar[s], 1u);
    }
    mbar_init(&done_bar, 1u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"(128u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (threadIdx.x == 0) WG_TRACE(7, 0);

  if (warp == 1) {
    constexpr uint32_t idesc = make_idesc_mn(128, 128);
    for (int it = 0; it < nchunks; ++it) {
      const int s = it % NS, ph = (it / NS) & 1;
      mbar_wait(&full_bar[s], ph);
      tc_fence_after();
      if (lane == 0) WG_TRACE(1, it);
      const uint32_t base = smem_u32(smem + (size_t)s * kWgStageBytes);
      const uint32_t p_hi = desc_lo_mn(base), p_lo = desc_lo_mn(base + 16384);
      const uint32_t q_hi = desc_lo_mn(base + 32768), q_lo = desc_lo_mn(base + 49152);
      if (elect_one_sync()) {
#pragma unroll
        for (int kg = 0; kg < 4; ++kg) {
          const uint32_t ko = (uint32_t)kg * (4096u >> 4);
          const uint32_t first = (it == 0 && kg == 0) ? 0u : 1u;
          if (q.passes == 3) {
            mma_tf32_mn<0>(tmem_base, p_lo + ko, q_hi + ko, idesc, first);
            mma_tf32_mn<1>(tmem_base, p_hi + ko, q_lo + ko, idesc, 1u);
            mma_tf32_mn<2>(tmem_base, p_hi + ko, q_hi + ko, idesc, 1u);
          } else {
            mma_tf32_mn(tmem_base, p_hi + ko, q_hi + ko, idesc, first);
          }
        }
        mma_commit(&empty_bar[s]);
        if (it == nchunks - 1) mma_commit(&done_bar);
        WG_TRACE(2, it);
      }
      __syncwarp();
    }
    if (nchunks == 0 && elect_one_sync()) mbar_arrive(&done_bar);
  } else if (warp >= 2) {
    const int tid = threadIdx.x - 64;
    const bool pvec = ((((uintptr_t)q.P.ptr) & 15) == 0) && (q.P.ld % 4 == 0);
    const bool qvec = ((((uintptr_t)q.Q.ptr) & 15) == 0) && (q.Q.ld % 4 == 0);
    if (!GEN && !(pvec && qvec)) asm volatile("trap;");
    // software-pipelined producer: kWgDepth chunks of loads in flight per thread (the loop is
    // load-latency bound: one dependent global load per chunk otherwise)
    constexpr int kWgDepth = 3;
    float4 v[kWgDepth][4];
    // 2048 float4 groups per chunk: [0,1024) operand P, [1024,2048) operand Q; group = (sample
    // bl, float4 index c4 along the rows) so that a warp reads 512 contiguous bytes of a sample
    auto load_chunk = [&](int it, float4 (&dst)[4]) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int gi = tid + g * PT;
        const bool isq = gi >= 1024;
        const int li = gi & 1023;
        const int bl = li >> 5, c4 = li & 31;
        const int b = b_begin + it * 32 + bl;
        if (!isq) dst[g] = wg_load<GEN>(q.P, m0 + c4 * 4, b, b < b_end, tval, pvec);
        else dst[g] = wg_load<GEN>(q.Q, c4 * 4, b, b < b_end, tval, qvec);
      }
    };
#pragma unroll
    for (int d = 0; d < kWgDepth; ++d)
      if (d < nchunks) load_chunk(d, v[d]);
    for (int it0 = 0; it0 < nchunks; it0 += kWgDepth) {
#pragma unroll
      for (int d = 0; d < kWgDepth; ++d) {
        const int it = it0 + d;
        if (it < nchunks) {
          const int s = it % NS, ph = (it / NS) & 1;
          uint8_t* stage = smem + (size_t)s * kWgStageBytes;
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (tid == 0) WG_TRACE(3, it);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int gi = tid + g * PT;
            const bool isq = gi >= 1024;
            const int li = gi & 1023;
            const int bl = li >> 5, c4 = li & 31;
            const int grp = c4 >> 3, c32 = (c4 & 7) >> 1, half = c4 & 1, k4 = bl >> 2, r = bl & 3;
            const uint32_t o = (uint32_t)((k4 * 4 + grp) * 512 + r * 128 + ((c32 ^ r) << 5) + half * 16);
            const float4 x = v[d][g];
            float4 hi, lo;
            if (q.passes == 3) {
              hi = make_float4(tf32_rna(x.x), tf32_rna(x.y), tf32_rna(x.z), tf32_rna(x.w));
              lo = make_float4(tf32_rna(x.x - hi.x), tf32_rna(x.y - hi.y), tf32_rna(x.z - hi.z),
                               tf32_rna(x.w - hi.w));
            } else { hi = x; lo = make_float4(0.f, 0.f, 0.f, 0.f); }
            uint8_t* img = stage + (isq ? 32768 : 0);
            *reinterpret_cast<float4*>(img + o) = hi;
            *reinterpret_cast<float4*>(img + 16384 + o) = lo;
          }
          if (tid == 0) WG_TRACE(4, it);
          if (it + kWgDepth < nchunks) load_chunk(it + kWgDepth, v[d]);
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&full_bar[s]);
        }
      }
    }
    // epilogue: lane = row of P's tile, 128 columns = rows of Q
    if (tid == 0) WG_TRACE(6, 0);
    mbar_wait(&done_bar, 0);
    tc_fence_after();
    if (tid == 0) WG_TRACE(6, 1);
    const int quarter = (warp & 3) * 32;
    const int wgroup = (warp - 2) >> 2;
    const int m = m0 + quarter + lane;
    const int mrows = q.P.rows + q.P.td + q.P.bias;
    const int nrows = q.Q.rows + q.Q.td + q.Q.bias;
    float* out = q.part + (size_t)blockIdx.y * q.block;
    for (int cb = wgroup; cb < 8; cb += kProducerWarps / 4) {
      const int c0 = cb * 16;
      uint32_t r[16];
      const uint32_t taddr = tmem_base + ((uint32_t)quarter << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, "
          "%13, %14, %15}, [%16];\n\t"
          "tcgen05.wait::ld.sync.aligned;"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
            "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(taddr)
          : "memory");
      if (m < mrows) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const int n = c0 + j;
          if (n < nrows) {
            const float val = (nchunks > 0) ? __uint_as_float(r[j]) : 0.0f;
            if (q.transposed) out[(size_t)m * q.ldo + n] = val;
            else out[(size_t)n * q.ldo + m] = val;
          }
        }
      }
    }
    tc_fence_before();
    if (tid == 0) WG_TRACE(6, 2);
  }
  __syncthreads();
  if (threadIdx.x == 0) WG_TRACE(6, 3);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u) : "memory");
  }
}

}  // namespace umma
