// Latent-space engine (see lrnde_fused.h): chain_kernel + kgemm_kernel, hand-written tcgen05 / TMEM / bulk-copy
// code for sm_100a.  Reference arithmetic: src/perform_step.jl:10-27 (stages, u, utilde), :34-38 and :208-212
// (residual, RMS), src/layers/common.jl:19-33 (TDChain time row), through the identities in the header.
#include "lrnde_fused.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "lrnde_fused_dev.cuh"

#ifdef LRNDE_UMMA_TRACE
__device__ long long g_fused_trace[4096];
// kernel 0 = chain, 1 = kgemm; CTA `blk` only; rows of 64 stamps
#define FTRACE(kern, blk, row, it) do { if ((int)blockIdx.x == (blk) && (it) < 64) g_fused_trace[(kern) * 2048 + (row) * 64 + (it)] = clock64(); } while (0)
extern "C" int lrnde_debug_trace_fused(long long* out, int n) {
  cudaMemcpyFromSymbol(out, g_fused_trace, sizeof(long long) * (size_t)n);
  return 0;
}
#else
#define FTRACE(kern, blk, row, it) do { } while (0)
#endif

namespace fused {
using namespace umma;

// ---------------------------------------------------------------------------------------------------------
// weight images (once per call: the parameters change every training iteration)
// ---------------------------------------------------------------------------------------------------------
// Mz = W1[:, :D] * W2a  ([H x Kaug], accumulated in double) as a plain zero-padded [128][128] matrix: one block per
// column k, thread = (row r, slice of the D-long dot product), fixed-order sum of the four slices
__global__ void __launch_bounds__(512) mz_plain_kernel(const float* __restrict__ W1, const float* __restrict__ W2a, int D, int H,
                                                       int Kaug, float* __restrict__ out) {
  __shared__ double part[4][128];
  const int r = threadIdx.x & 127, sl = threadIdx.x >> 7, k = blockIdx.x;
  double acc0 = 0.0, acc1 = 0.0;
  if (r < H && k < Kaug) {
    const float* w2 = W2a + (size_t)k * D;
    int d = sl;
    for (; d + 4 < D; d += 8) {
      acc0 += (double)W1[(size_t)d * H + r] * (double)w2[d];
      acc1 += (double)W1[(size_t)(d + 4) * H + r] * (double)w2[d + 4];
    }
    for (; d < D; d += 4) acc0 += (double)W1[(size_t)d * H + r] * (double)w2[d];
  }
  part[sl][r] = acc0 + acc1;
  __syncthreads();
  if (sl == 0) out[r * 128 + k] = (float)(((part[0][r] + part[1][r]) + part[2][r]) + part[3][r]);
}
// ... and as the 128-row tf32 hi / lo A-operand image of the chain kernel
__global__ void mz_image_kernel(const float* __restrict__ plain, int Kaug, int nfull, float* __restrict__ img, uint32_t img_bytes,
                                int passes) {
  const int r = threadIdx.x, k = blockIdx.x;
  const float v = (k < 128) ? plain[r * 128 + k] : 0.0f;
  const float hi = (passes == 1) ? v : tf32_rna(v);
  const float lo = (passes == 1) ? 0.0f : tf32_rna(v - hi);
  const uint32_t o = img_off(128, nfull, r, k);
  *reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(img) + o) = hi;
  *reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(img) + img_bytes + o) = lo;
}
// ---------------------------------------------------------------------------------------------------------
// (1) chain kernel: the stages of one attempt as a recurrence in latent space, one CTA per 64 samples.
// The 64 samples are two independent halves of 32 (column group cg): while the tensor core multiplies one
// half's hidden activations by Mz, the other half's warps compute their next activation.
// Operand images leave in the layout the kgemm kernel reads: units of 16 samples, 96 rows = 6 stages x 16.
// ---------------------------------------------------------------------------------------------------------
struct ChainP {
  SolveDev* S;
  const LinComb* single;      // nullptr: the six stages of S->st[0..6]; else: one evaluation of this descriptor
  const LinComb* single_out;  // single mode: descriptor whose dst receives the result (nullptr: single->dst)
  const int* done;
  const float* Mimg;
  uint32_t imgM;
  const float* w1t;  // layer-1 time column (nullptr without TDChain)
  const float* b1;
  float* hbuf;
  uint32_t unit_bytes;
  int B, H, td, KS, nfull, ntail, passes, nbuf, write_z;
  int combo;   // lean attempt: export only the operand images of sum_i a_7i H(k_i) and sum_i btilde_i H(k_i) (units of
               // 16 samples x 2 row groups) for kgemm_kernel<2, false>: u_{n+1} and the residual need nothing else
};

template <int ACT>
__global__ void __launch_bounds__(kThreads, 1) chain_kernel(ChainP p) {
  if (p.done && *p.done) return;
  if (!p.single && p.S->done) return;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t abar, bfull[2][2], bfree[2], zdone[2];
  __shared__ uint32_t tmem_slot;
  __shared__ LinComb sd[7];
  __shared__ float s_bt[7];
  __shared__ float* s_tape;
  __shared__ float* s_ztape;
  __shared__ float* s_htape;
  __shared__ size_t s_len, s_zlen;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * kNT;
  const int nst = p.single ? 1 : 6;
  const uint32_t tileB = (uint32_t)kNT * p.KS * 32 * 2;
  uint8_t* smA = smem;
  uint8_t* smB = smem + 2 * (size_t)p.imgM;

  if (threadIdx.x == 0) {
    mbar_init(&abar, 1u);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&bfull[0][b], (uint32_t)(kEpiWarps / 2));
      mbar_init(&bfull[1][b], (uint32_t)(kEpiWarps / 2));
      mbar_init(&bfree[b], 1u);
      mbar_init(&zdone[b], 1u);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (p.single) sd[0] = *p.single;
    else for (int j = 0; j < 7; ++j) sd[j] = p.S->st[j];
    if (p.single) sd[1] = p.single_out ? *p.single_out : *p.single;
    for (int j = 0; j < 7; ++j) s_bt[j] = p.S->err.coef[j];
    s_tape = p.S->tape; s_ztape = p.S->ztape; s_htape = p.S->htape; s_len = p.S->len; s_zlen = p.S->zlen;
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (threadIdx.x == 0) FTRACE(0, 1, 0, 0);
  auto zof = [&](const float* ptr) -> float* { return s_ztape + ((size_t)(ptr - s_tape) / s_len) * s_zlen; };
  auto hof = [&](const float* ptr) -> float* { return s_htape + ((size_t)(ptr - s_tape) / s_len) * s_zlen; };

  if (warp == 0) {
    // ---------------- Mz images in, operand images out (one elected lane; converged warp)
    if (elect_one_sync()) {
      mbar_arrive_expect_tx(&abar, 2 * p.imgM);
      bulk_g2s(smA, p.Mimg, 2 * p.imgM, &abar);
    }
    __syncwarp();
    for (int st = 0; st < nst; ++st) {
      const int b = st % p.nbuf;
      const uint32_t ph = (uint32_t)((st / p.nbuf) & 1);
      mbar_wait(&bfull[0][b], ph);
      mbar_wait(&bfull[1][b], ph);
      if (p.combo) {          // nothing to export per stage; the MMA of this stage has its own barrier (zdone)
        if (elect_one_sync()) mbar_arrive(&bfree[b]);
        __syncwarp();
        continue;
      }
      if (elect_one_sync()) {
        // the 64 rows of this stage = rows [16 j, 16 j + 16) of four 16-sample unit tiles
        const int j = p.single ? 0 : st;
        const uint8_t* s = smB + (size_t)b * tileB;
        for (int sg = 0; sg < 4; ++sg) {
          uint8_t* g = reinterpret_cast<uint8_t*>(p.hbuf) + ((size_t)blockIdx.x * 4 + sg) * p.unit_bytes;
          for (int c = 0; c < p.nfull; ++c)
            for (int lo = 0; lo < 2; ++lo)
              bulk_s2g(g + (size_t)c * kPieceBytes + lo * (kUR * 128) + j * 2048,
                       s + (size_t)c * (2 * kNT * 128) + lo * (kNT * 128) + sg * 2048, 2048);
          for (int t = 0; t < p.ntail; ++t)
            for (int lo = 0; lo < 2; ++lo)
              bulk_s2g(g + (size_t)p.nfull * kPieceBytes + (size_t)t * kTailBytes + lo * (kUR * 32) + j * 512,
                       s + (size_t)p.nfull * (2 * kNT * 128) + (size_t)t * (2 * kNT * 32) + lo * (kNT * 32) + sg * 512, 512);
        }
        bulk_commit();
        bulk_wait_read0();
        mbar_arrive(&bfree[b]);
      }
      __syncwarp();
    }
    if (p.combo) {
      // the two combination tiles (buffer 0: sum a_7i H(k_i), buffer 1: sum btilde_i H(k_i)) as row groups 0 / 1 of
      // units of 16 samples x 2 groups (the operand layout of kgemm_kernel<2, *>)
      const uint32_t ph = (uint32_t)((6 / p.nbuf) & 1);
      for (int b = 0; b < 2; ++b) { mbar_wait(&bfull[0][b], ph); mbar_wait(&bfull[1][b], ph); }
      if (elect_one_sync()) {
        for (int j = 0; j < 2; ++j) {
          const uint8_t* s = smB + (size_t)j * tileB;
          for (int sg = 0; sg < 4; ++sg) {
            uint8_t* g = reinterpret_cast<uint8_t*>(p.hbuf) + ((size_t)blockIdx.x * 4 + sg) * p.unit_bytes;
            for (int c = 0; c < p.nfull; ++c)
              for (int lo = 0; lo < 2; ++lo)
                bulk_s2g(g + (size_t)c * 8192 + lo * 4096 + j * 2048,
                         s + (size_t)c * (2 * kNT * 128) + lo * (kNT * 128) + sg * 2048, 2048);
            for (int t = 0; t < p.ntail; ++t)
              for (int lo = 0; lo < 2; ++lo)
                bulk_s2g(g + (size_t)p.nfull * 8192 + (size_t)t * 2048 + lo * 1024 + j * 512,
                         s + (size_t)p.nfull * (2 * kNT * 128) + (size_t)t * (2 * kNT * 32) + lo * (kNT * 32) + sg * 512, 512);
          }
        }
        bulk_commit();
      }
      __syncwarp();
    }
    if (elect_one_sync()) bulk_wait0();
    __syncwarp();
  } else if (warp == 1) {
    // ---------------- tcgen05.mma issue: Z(k_j) = Mz [h_j ; t_j ; 1]   (M = 128 latent rows x N = 32 samples per half)
    constexpr uint32_t idesc = make_idesc(128, kNT / 2);
    mbar_wait(&abar, 0);
    if (lane == 0) FTRACE(0, 1, 0, 1);
    const uint32_t a0 = smem_u32(smA);
    for (int st = 0; st < nst; ++st) {
      const int b = st % p.nbuf;
      for (int cg = 0; cg < 2; ++cg) {
        mbar_wait(&bfull[cg][b], (uint32_t)((st / p.nbuf) & 1));
        tc_fence_after();
        if (lane == 0) FTRACE(0, 1, 1, 2 * st + cg);
        const uint32_t b0 = smem_u32(smB + (size_t)b * tileB);
        const uint32_t d = tmem_base + (uint32_t)((p.single ? 7 : 2 + st) * kNT + cg * (kNT / 2));
        if (elect_one_sync()) {
          uint32_t first = 0u;
          for (int c = 0; c < p.nfull; ++c) {
            const uint32_t ah = desc_lo(a0 + c * (128 * 128)), al = desc_lo(a0 + p.imgM + c * (128 * 128));
            const uint32_t bb = b0 + c * (2 * kNT * 128) + cg * (kNT / 2) * 128;
            const uint32_t bh = desc_lo(bb), bl = desc_lo(bb + kNT * 128);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              mma_kstep(d, ah + 2 * k, al + 2 * k, bh + 2 * k, bl + 2 * k, kHi128, idesc, first, p.passes);
              first = 1u;
            }
          }
          for (int t = 0; t < p.ntail; ++t) {
            const uint32_t ao = p.nfull * (128 * 128) + t * (128 * 32);
            const uint32_t bb = b0 + p.nfull * (2 * kNT * 128) + t * (2 * kNT * 32) + cg * (kNT / 2) * 32;
            mma_kstep(d, desc_lo(a0 + ao), desc_lo(a0 + p.imgM + ao), desc_lo(bb), desc_lo(bb + kNT * 32), kHi32, idesc,
                      first, p.passes);
            first = 1u;
          }
          mma_commit(&zdone[cg]);
          FTRACE(0, 1, 2, 2 * st + cg);
        }
        __syncwarp();
      }
    }
  } else {
    // ---------------- stage arithmetic: lane = latent row, registers = 16 samples
    const int q = warp & 3, sub = (warp - 2) >> 2;
    const int cg = sub >> 1;
    const int hrow = q * 32 + lane;
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    const int col = sub * 16;
    const bool rowv = hrow < p.H;
    const float w1t = (p.td && rowv && p.w1t) ? p.w1t[hrow] : 0.0f;
    const float b1 = rowv ? p.b1[hrow] : 0.0f;
    const bool tr = (threadIdx.x == 64);
    // inputs computed earlier: Z(uprev) -> slot 0, Z(sources) -> slots 1..
    const int next = p.single ? 1 + sd[0].n : 2;
    if (p.single && sd[0].n > 6) asm volatile("trap;");
    for (int e = 0; e < next; ++e) {
      const float* src = (e == 0) ? sd[0].base : sd[0].src[e - 1];
      const float* zp = src ? zof(src) : nullptr;
      float v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int n = n0 + col + i;
        v[i] = (zp && rowv && n < p.B) ? __ldcg(zp + (size_t)n * LR_ZROW + hrow) : 0.0f;
      }
      tmem_st16(tlane + (uint32_t)(e * kNT + col), v);
    }
    tmem_st_wait();
    if (tr) FTRACE(0, 1, 0, 2);
    // operand-image position of (sample row, k = hrow): chunk q of the tile when hrow is inside a full chunk
    const bool in_full = hrow < p.nfull * 32;
    const bool in_img = hrow < p.KS * 8;
    uint32_t obase, ostep_lo;
    if (in_full) {
      obase = (uint32_t)(q * (2 * kNT * 128) + (lane & 3) * 4);
      ostep_lo = (uint32_t)(kNT * 128);
    } else {
      const int kt = hrow - p.nfull * 32, t = kt >> 3, kk = kt & 7;
      obase = (uint32_t)(p.nfull * (2 * kNT * 128) + t * (2 * kNT * 32) + (kk & 3) * 4) | (uint32_t)(((kk >> 2) & 1) << 30);
      ostep_lo = (uint32_t)(kNT * 32);
    }
    // cumulative hidden image of the state: u_{n+1} = x + W2a C_{n+1},  C_{n+1} = C_n + dt sum_i a_7i H(k_i)
    // (hidden tape of the latent-space adjoint); H(k_1) comes from the tape, H(k_2..6) from the stages below
    const bool write_c = !p.single && s_htape && hrow < p.H + p.td + 1;
    float csum[16], bsum[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { csum[i] = 0.0f; bsum[i] = 0.0f; }
    if (write_c) {
      const float* h1 = hof(sd[5].src[0]);
      const float cf = sd[5].coef[0], cb = s_bt[0];
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (n0 + col + i < p.B) {
          const float h1v = __ldcg(h1 + (size_t)(n0 + col + i) * LR_ZROW + hrow);
          csum[i] = cf * h1v;
          if (p.combo) bsum[i] = cb * h1v;
        }
    }
    auto tile_put16 = [&](uint8_t* tile, const float (&val)[16]) {
      if (!in_img) return;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int row = col + i;
        float hi, lo;
        if (p.passes == 3) { hi = tf32_rna(val[i]); lo = tf32_rna(val[i] - hi); }
        else { hi = val[i]; lo = 0.0f; }
        uint32_t o;
        if (in_full) o = obase + (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + (((lane >> 2) ^ (row & 7)) << 4));
        else o = (obase & 0x3FFFFFFFu) + (uint32_t)(row * 32 + ((((obase >> 30) & 1) ^ ((row >> 2) & 1)) << 4));
        *reinterpret_cast<float*>(tile + o) = hi;
        *reinterpret_cast<float*>(tile + o + ostep_lo) = lo;
      }
    };
    for (int st = 0; st < nst; ++st) {
      const int b = st % p.nbuf;
      const LinComb& d = sd[st];
      const bool last_full = (!p.single && st == 5);
      const float tstage = last_full ? sd[6].t : d.t;
      const int nsrc = d.n;
      float* zlin = last_full ? zof(d.dst) : nullptr;   // Z(u_{n+1}) = Z(uprev) + dt * sum a_7i Z(k_i)
      if (tr) FTRACE(0, 1, 3, st);
      if (st >= p.nbuf) mbar_wait(&bfree[b], (uint32_t)(((st / p.nbuf) - 1) & 1));
      if (tr) FTRACE(0, 1, 4, st);
      uint8_t* tile = smB + (size_t)b * tileB;
      float inner[16], v[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) inner[i] = 0.0f;
      for (int s = 0; s < nsrc; ++s) {
        tmem_ld16(tlane + (uint32_t)((1 + s) * kNT + col), v);
        const float cf = d.coef[s];
#pragma unroll
        for (int i = 0; i < 16; ++i) inner[i] = fmaf(cf, v[i], inner[i]);
      }
      tmem_ld16(tlane + (uint32_t)col, v);
      const float scale = d.scale;
      if (nsrc) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = fmaf(scale, inner[i], v[i]);
      }
      if (zlin && rowv) {
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (n0 + col + i < p.B) zlin[(size_t)(n0 + col + i) * LR_ZROW + hrow] = v[i];
      }
      // activation (branch-free over the 16 samples: the compiler interleaves the dependent chains)
      const float hconst = (p.td && hrow == p.H) ? tstage : ((hrow == p.H + p.td) ? 1.0f : 0.0f);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float x = p.td ? fmaf(w1t, tstage, v[i]) : v[i];
        x += b1;
        const float h = lr_act(ACT, x);
        v[i] = rowv ? h : hconst;
      }
      if (write_c) {
        if (p.combo) {
          const float cb = s_bt[st + 1];
#pragma unroll
          for (int i = 0; i < 16; ++i) bsum[i] = fmaf(cb, v[i], bsum[i]);
        }
        if (!last_full) {
          const float cf = sd[5].coef[st + 1];
#pragma unroll
          for (int i = 0; i < 16; ++i) csum[i] = fmaf(cf, v[i], csum[i]);
        } else {
          const float* c0 = hof(sd[5].base);
          float* c1 = hof(sd[5].dst);
          const float sc = sd[5].scale;
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (n0 + col + i < p.B) {
              const size_t e = (size_t)(n0 + col + i) * LR_ZROW + hrow;
              c1[e] = fmaf(sc, csum[i], __ldcg(c0 + e));
            }
        }
      }
      const float* kdst = p.single ? sd[1].dst : (last_full ? sd[6].dst : d.dst);
      if (s_htape && hrow < p.H + p.td + 1) {
        // hidden tape for the latent-space adjoint (lrnde_adjoint.cu): H(k_j) = [h_j ; t_j ; 1]
        float* ho = hof(kdst);
#pragma unroll
        for (int i = 0; i < 16; ++i)
          if (n0 + col + i < p.B) ho[(size_t)(n0 + col + i) * LR_ZROW + hrow] = v[i];
      }
      tile_put16(tile, v);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bfull[cg][b]);
      if (tr) FTRACE(0, 1, 5, st);
      mbar_wait(&zdone[cg], (uint32_t)(st & 1));
      tc_fence_after();
      if (tr) FTRACE(0, 1, 6, st);
      // Z(k_j) to the latent tape
      if ((p.single && p.write_z) || last_full) {   // intermediate Z(k_j) are not read back by anything
        float* zo = zof(kdst);
        const int slot = p.single ? 7 : 2 + st;
        tmem_ld16(tlane + (uint32_t)(slot * kNT + col), v);
        if (rowv) {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (n0 + col + i < p.B) zo[(size_t)(n0 + col + i) * LR_ZROW + hrow] = v[i];
        }
      }
    }
    if (p.combo) {
      // every MMA has read its tile (zdone of the last stage): both buffers are free for the two combination tiles
#pragma unroll
      for (int i = 0; i < 16; ++i) { if (n0 + col + i >= p.B) { csum[i] = 0.0f; bsum[i] = 0.0f; } }
      tile_put16(smB, csum);
      tile_put16(smB + (size_t)tileB, bsum);
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) { mbar_arrive(&bfull[cg][0]); mbar_arrive(&bfull[cg][1]); }
    }
    tc_fence_before();
    if (tr) FTRACE(0, 1, 0, 3);
  }
  __syncthreads();
  if (threadIdx.x == 0) FTRACE(0, 1, 0, 4);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------
// (2) kgemm kernel: k_j = W2a [h_j ; t_j ; 1] for all stages of the attempt, epilogue = tape + u_{n+1} + residual.
// One CTA = one tile of MT output features; W2a (tf32 hi / lo) sits in TENSOR MEMORY as the A operand for the whole
// launch (the MMAs read only the activation operand from shared memory: N = 96 runs at the tensor floor), the
// CTAs of a cluster are the feature tiles of the same samples, so the operand pieces are fetched once per
// cluster and multicast.  A unit = 16 samples x 6 stages = 96 accumulator columns, double-buffered in TMEM.
// ---------------------------------------------------------------------------------------------------------
struct KgemmP {
  SolveDev* S;
  const LinComb* single;
  const LinComb* single_out;
  const int* done;
  const float* W2a;   // [D x Kaug] column-major block of the parameter vector (weight | time column | bias)
  const float* hbuf;
  uint32_t unit_bytes;
  int B, D, MT, n_mt, Kaug, KS, nfull, ntail, passes, nunits, nclusters, ring;
  int dbg;   // LRNDE_KG_DBG experiments (profiling only): 1 = no k stores, 2 = no loads, 4 = no residual, 8 = no u store
  int lean;  // lean tape: k_2..k_6 are not stored (the latent-space adjoint and dense output read the hidden tape)
  const float* add_base;   // single mode: dst = add_base + W2a [h ; t ; 1]  (dense output y(t) = x + W2a c(t))
};

// NSTG = stage columns per sample (6: k_2..k_7 of a forward attempt; 2: the b / btilde combinations of an adjoint attempt);
// ADJ: A = W1[:, :D]^T (features x hidden), epilogue = lambda_{n+1} and its residual (lrnde_adjoint.cu)
template <int NSTG, bool ADJ>
__global__ void __launch_bounds__(kThreads, 1) kgemm_kernel(KgemmP p) {
  constexpr int UR = NSTG * 16;                   // rows of an operand unit
  constexpr int PIECE = 2 * UR * 128, TAIL = 2 * UR * 32;
  if (p.done && *p.done) return;
  if (!p.single && p.S->done) return;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smR = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t a_ready, full_bar[2], empty_bar[2], acc_full[2], tmem_free[2];
  __shared__ uint32_t tmem_slot;
  __shared__ LinComb s_err, s_un;
  __shared__ double s_red[kEpiWarps];

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t csize = cluster_nctarank();
  const uint32_t crank = (csize > 1) ? cluster_ctarank() : 0u;
  const uint16_t cmask = (uint16_t)((1u << csize) - 1u);
  const int mt = (int)(blockIdx.x % (unsigned)p.n_mt);
  const int cid = (int)(blockIdx.x / (unsigned)p.n_mt);
  const int lo_col = p.KS * 8;
  const uint32_t stage_bytes = (p.unit_bytes + 1023u) & ~1023u;   // one ring stage = the operand images of one unit

  if (threadIdx.x == 0) {
    mbar_init(&a_ready, (uint32_t)kEpiWarps);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&full_bar[s], 1u);
      mbar_init(&empty_bar[s], csize);
      mbar_init(&acc_full[s], 1u);
      mbar_init(&tmem_free[s], (uint32_t)kEpiWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (p.single) { s_err = *p.single; s_un = p.single_out ? *p.single_out : *p.single; }
    else { s_err = p.S->err; s_un = p.S->st[5]; }
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (csize > 1) cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  if (threadIdx.x == 0) FTRACE(1, 0, 0, 0);

  if (warp == 0) {
    // ---------------- operand images of every unit: each CTA of the cluster fetches a slice and multicasts it
    int u = 0;
    for (int g = cid; g < p.nunits; g += p.nclusters, ++u) {
      const uint8_t* src = reinterpret_cast<const uint8_t*>(p.hbuf) + (size_t)g * p.unit_bytes;
      const int s = u & 1, ph = (u >> 1) & 1;
      mbar_wait(&empty_bar[s], (uint32_t)(ph ^ 1));
      uint8_t* dst = smR + (size_t)s * stage_bytes;
      if (elect_one_sync()) {
        FTRACE(1, 0, 1, u);
        mbar_arrive_expect_tx(&full_bar[s], p.unit_bytes);
        if (csize > 1) {
          const uint32_t sl = (((p.unit_bytes + csize - 1) / csize) + 15u) & ~15u;
          const uint32_t o = crank * sl;
          if (o < p.unit_bytes) bulk_g2s_mc(dst + o, src + o, min(sl, p.unit_bytes - o), &full_bar[s], cmask);
        } else {
          bulk_g2s(dst, src, p.unit_bytes, &full_bar[s]);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ---------------- tcgen05.mma issue: D[128 features x (6 stages x 16 samples)] = A (TMEM) x operand images
    const uint32_t idesc = p.single ? make_idesc(128, 16) : make_idesc(128, UR);
    mbar_wait(&a_ready, 0);
    tc_fence_after();
    if (lane == 0) FTRACE(1, 0, 0, 1);
    int u = 0;
    for (int g = cid; g < p.nunits; g += p.nclusters, ++u) {
      const int db = u & 1;
      if (u >= 2) { mbar_wait(&tmem_free[db], (uint32_t)(((u >> 1) - 1) & 1)); tc_fence_after(); }
      mbar_wait(&full_bar[db], (uint32_t)((u >> 1) & 1));
      tc_fence_after();
      if (lane == 0) FTRACE(1, 0, 2, u);
      const uint32_t d = tmem_base + (uint32_t)(kDCol + db * 128);
      const uint32_t b0 = smem_u32(smR + (size_t)db * stage_bytes);
      if (elect_one_sync()) {
        uint32_t first = 0u;
        for (int pc = 0; pc < p.nfull; ++pc) {
          const uint32_t bh = desc_lo(b0 + pc * PIECE), bl = desc_lo(b0 + pc * PIECE + UR * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t ah = tmem_base + (uint32_t)((pc * 4 + k) * 8), al = ah + (uint32_t)lo_col;
            if (p.passes == 3) {
              mma_ts(d, al, bh + 2 * k, kHi128, idesc, first);
              mma_ts(d, ah, bl + 2 * k, kHi128, idesc, 1u);
              mma_ts(d, ah, bh + 2 * k, kHi128, idesc, 1u);
            } else mma_ts(d, ah, bh + 2 * k, kHi128, idesc, first);
            first = 1u;
          }
        }
        for (int t = 0; t < p.ntail; ++t) {
          const uint32_t bb = b0 + p.nfull * PIECE + t * TAIL;
          const uint32_t bh = desc_lo(bb), bl = desc_lo(bb + UR * 32);
          const uint32_t ah = tmem_base + (uint32_t)((p.nfull * 4 + t) * 8), al = ah + (uint32_t)lo_col;
          if (p.passes == 3) {
            mma_ts(d, al, bh, kHi32, idesc, first);
            mma_ts(d, ah, bl, kHi32, idesc, 1u);
            mma_ts(d, ah, bh, kHi32, idesc, 1u);
          } else mma_ts(d, ah, bh, kHi32, idesc, first);
          first = 1u;
        }
        if (csize > 1) mma_commit_mc(&empty_bar[db], cmask);
        else mma_commit(&empty_bar[db]);
        mma_commit(&acc_full[db]);
        FTRACE(1, 0, 3, u);
      }
      __syncwarp();
    }
  } else {
    // ---------------- warps 2..17: lane = output feature (coalesced global access), registers = 4 samples
    const int q = warp & 3, sub = (warp - 2) >> 2;
    const int mloc = q * 32 + lane;
    const int m = mt * p.MT + mloc;
    const bool mv = (mloc < p.MT) && (m < p.D);
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    const bool tr = (threadIdx.x == 64);
    // W2a tile -> tensor memory, tf32 hi / lo (warp `sub` of a lane quarter writes the K-steps ks = sub, sub + 4, ...)
    {
      float w[4][8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ks = sub + 4 * j;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const int k = ks * 8 + e;
          if (ADJ) w[j][e] = (mv && k < p.Kaug) ? __ldg(p.W2a + (size_t)m * p.Kaug + k) : 0.0f;   // W1[k, m]: [H x D] column-major, Kaug = H here
          else w[j][e] = (mv && k < p.Kaug) ? __ldg(p.W2a + (size_t)k * p.D + m) : 0.0f;
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ks = sub + 4 * j;
        if (ks < p.KS) {
          float hi[8], lo[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            if (p.passes == 3) { hi[e] = tf32_rna(w[j][e]); lo[e] = tf32_rna(w[j][e] - hi[e]); }
            else { hi[e] = w[j][e]; lo[e] = 0.0f; }
          }
          tmem_st8(tlane + (uint32_t)(ks * 8), hi);
          tmem_st8(tlane + (uint32_t)(lo_col + ks * 8), lo);
        }
      }
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(&a_ready);

    const float abstol = p.S->abstol, reltol = p.S->reltol;
    const unsigned D = (unsigned)p.D;
    double acc = 0.0;
    // array base pointers at this thread's feature (element index inside an array = sample * D: 32 bits)
    const float* pu = s_err.base + m;
    const float* pk1 = s_err.src[0] + m;
    float* po[7];
#pragma unroll
    for (int jj = 0; jj < 6; ++jj) po[jj] = (ADJ || NSTG == 2) ? nullptr : const_cast<float*>(s_err.src[jj + 1]) + m;
    po[6] = s_err.dst + m;
    float* psingle = (p.single ? s_un.dst : s_err.dst) + m;
    float ca[6], cb[7];
#pragma unroll
    for (int jj = 0; jj < 6; ++jj) ca[jj] = s_un.coef[jj];
#pragma unroll
    for (int jj = 0; jj < 7; ++jj) cb[jj] = s_err.coef[jj];
    const float sdt = s_un.scale, edt = s_err.scale;
    // uprev and k1 = fsalfirst are the only arrays an attempt reads: fetched one unit ahead
    float up[4], k1[4];
    auto fetch = [&](int g, float (&a)[4], float (&b)[4]) {
      const int nb = g * 16 + sub * 4;
      if (mv && nb + 4 <= p.B && !(p.dbg & 2)) {
        const unsigned e = (unsigned)nb * D;
#pragma unroll
        for (int i = 0; i < 4; ++i) { a[i] = __ldcg(pu + (e + (unsigned)i * D)); b[i] = (ADJ || NSTG == 2) ? 0.0f : __ldcg(pk1 + (e + (unsigned)i * D)); }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const bool ok = mv && (nb + i < p.B) && !(p.dbg & 2);
          a[i] = ok ? __ldcg(pu + (size_t)(nb + i) * D) : 0.0f;
          b[i] = (ok && !ADJ && NSTG != 2) ? __ldcg(pk1 + (size_t)(nb + i) * D) : 0.0f;
        }
      }
    };
    // one unit: accumulators -> registers, TMEM buffer released, next unit's inputs requested, outputs stored
    auto unit = [&](int g, int u, float (&upc)[4], float (&k1c)[4], float (&upn)[4], float (&k1n)[4]) {
      const int db = u & 1;
      const int nb = g * 16 + sub * 4;                  // first sample of this thread's 4
      const bool full = mv && (nb + 4 <= p.B);
      const unsigned e0 = (unsigned)nb * D;             // element index of (sample nb, this feature)
      const uint32_t dcol = tlane + (uint32_t)(kDCol + db * 128 + sub * 4);
      if (tr) FTRACE(1, 0, 4, 2 * u);
      mbar_wait(&acc_full[db], (uint32_t)((u >> 1) & 1));
      tc_fence_after();
      if (tr) FTRACE(1, 0, 4, 2 * u + 1);
      if (p.single) {
        float v[4];
        tmem_ld4(dcol, v);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tmem_free[db]);
        if (mv) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (nb + i < p.B) psingle[e0 + (unsigned)i * D] = p.add_base ? v[i] + __ldcg(p.add_base + m + (e0 + (unsigned)i * D)) : v[i];
        }
        return;
      }
      float k[NSTG][4];
#pragma unroll
      for (int jj = 0; jj < NSTG; ++jj) tmem_ld4(dcol + (uint32_t)(jj * 16), k[jj]);
      tmem_ld_wait();
      if (tr) FTRACE(1, 0, 6, u);
      // every accumulator of this thread is in registers: the MMAs of unit u + 2 may overwrite them
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_free[db]);
      if (g + p.nclusters < p.nunits) fetch(g + p.nclusters, upn, k1n);
      if (tr) FTRACE(1, 0, 7, u);
      float un[4], ut[4], unew[4];
      if constexpr (ADJ) {
        // accumulators = W1^T Delta_b, W1^T Delta_btilde: lambda' = -W1^T delta  =>  lambda_{n+1} = lambda_n - dt (W1^T Delta_b),
        // utilde = -dt (W1^T Delta_btilde)   (perform_step.jl:18-27 on the lambda block of the adjoint state)
#pragma unroll
        for (int i = 0; i < 4; ++i) { unew[i] = fmaf(-edt, k[0][i], upc[i]); ut[i] = -k[NSTG - 1][i]; }
      } else if constexpr (NSTG == 2) {
        // lean forward attempt: accumulators = W2a sum_i a_7i H(k_i) and W2a sum_i btilde_i H(k_i) (k_1 included), i.e.
        // sum_i a_7i k_i and sum_i btilde_i k_i of perform_step.jl:18-27 by linearity of the output layer
#pragma unroll
        for (int i = 0; i < 4; ++i) { unew[i] = fmaf(sdt, k[0][i], upc[i]); ut[i] = k[1][i]; }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) { un[i] = ca[0] * k1c[i]; ut[i] = cb[0] * k1c[i]; }
#pragma unroll
        for (int jj = 0; jj < NSTG; ++jj) {   // stage jj + 2
          if (jj < NSTG - 1) {
#pragma unroll
            for (int i = 0; i < 4; ++i) un[i] = fmaf(ca[jj + 1], k[jj][i], un[i]);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) ut[i] = fmaf(cb[jj + 1], k[jj][i], ut[i]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) unew[i] = fmaf(sdt, un[i], upc[i]);
      }
      if (tr) FTRACE(1, 0, 8, u);
      if (full) {
        if (!ADJ && NSTG != 2 && !(p.dbg & 1)) {
#pragma unroll
          for (int jj = 0; jj < NSTG; ++jj)
            if (jj == NSTG - 1 || !p.lean) {   // k_7 = fsalfirst of the next attempt is always needed
#pragma unroll
              for (int i = 0; i < 4; ++i) __stcg(po[jj] + (e0 + (unsigned)i * D), k[jj][i]);
            }
        }
        if (!(p.dbg & 8)) {
#pragma unroll
          for (int i = 0; i < 4; ++i) __stcg(po[6] + (e0 + (unsigned)i * D), unew[i]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (!(p.dbg & 4)) {
            const float r = (edt * ut[i]) / (abstol + fmaxf(fabsf(upc[i]), fabsf(unew[i])) * reltol);
            acc += (double)(r * r);
          } else acc += (double)unew[i];
        }
      } else if (mv) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (nb + i < p.B) {
            const size_t e = (size_t)(nb + i) * D;
            if (!ADJ && NSTG != 2) {
#pragma unroll
              for (int jj = 0; jj < NSTG; ++jj)
                if (jj == NSTG - 1 || !p.lean) po[jj][e] = k[jj][i];
            }
            po[6][e] = unew[i];
            const float r = (edt * ut[i]) / (abstol + fmaxf(fabsf(upc[i]), fabsf(unew[i])) * reltol);
            acc += (double)(r * r);
          }
        }
      }
      if (tr) FTRACE(1, 0, 5, u);
    };
    // two register sets for the fetched inputs: the unit loop is unrolled by two so that the values requested
    // during unit u are first touched in unit u + 1 (no register moves that would wait for the loads)
    float upB[4], k1B[4];
    if (!p.single) fetch(cid, up, k1);
    int u = 0;
    for (int g = cid; g < p.nunits; g += 2 * p.nclusters, u += 2) {
      unit(g, u, up, k1, upB, k1B);
      if (g + p.nclusters < p.nunits) unit(g + p.nclusters, u + 1, upB, k1B, up, k1);
    }
    if (!p.single) {
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
      if (lane == 0) s_red[warp - 2] = acc;
      asm volatile("bar.sync 1, %0;" ::"r"(kEpiWarps * 32) : "memory");
      if (warp == 2 && lane == 0) {
        double s = 0.0;
        for (int w = 0; w < kEpiWarps; ++w) s += s_red[w];
        p.S->partials[blockIdx.x] = s;
        for (unsigned i = blockIdx.x + gridDim.x; i < (unsigned)LR_ERR_BLOCKS; i += gridDim.x) p.S->partials[i] = 0.0;
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (csize > 1) cluster_sync_all();   // peers may still multicast into this CTA / signal its barriers
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// Dense output in hidden space: the operand image (stage-0 rows of every 16-sample unit) of
//   c(t) = C_n + dt_n sum_i b_i(theta) H(k_i)     so that   y(t) = x + W2a c(t)   (single-mode kgemm with add_base)
struct InterpP {
  const float* src[8];   // C_n, H(k_1) ... H(k_7) of the interval
  float coef[7];
  float scale;
  const SolveDev* A;     // non-null: interval / weights from the device descriptor yd of this (adjoint) solve instead
  const LinComb* yd;
  LinComb* ydesc_out;    // with A: receives {base = ybase, n = 0, t = yd->t} (the materialised y(t) as a descriptor)
  float* ybase;
  float* hbuf;
  uint32_t unit_bytes;
  int B, Kaug, KS, nfull, passes;
};
__global__ void __launch_bounds__(256) interp_image_kernel(InterpP p) {
  const int kk = threadIdx.x & 127;
  const int n = blockIdx.x * 2 + (threadIdx.x >> 7);
  if (p.A && p.A->failed) return;
  if (p.A && p.ydesc_out && blockIdx.x == 0 && threadIdx.x == 0) {
    LinComb d;
    d.base = p.ybase; d.dst = nullptr; d.n = 0; d.scale = 0.0f; d.t = p.yd->t; d.pad_ = 0;
    for (int k = 0; k < LR_MAXSRC; ++k) { d.src[k] = nullptr; d.coef[k] = 0.0f; }
    *p.ydesc_out = d;
  }
  if (n >= p.B || kk >= p.KS * 8) return;
  float v = 0.0f;
  if (kk < p.Kaug) {
    const size_t e = (size_t)n * LR_ZROW + kk;
    float acc = 0.0f;
    if (p.A) {
      const size_t slotn = (size_t)(p.yd->base - p.A->ftape) / ((size_t)7 * p.A->flen);
      const float* hb = p.A->fhtape + slotn * 7 * p.A->fzlen;
#pragma unroll
      for (int i = 0; i < 7; ++i) acc = fmaf(p.yd->coef[i], hb[(size_t)(i < 6 ? i + 1 : 8) * p.A->fzlen + e], acc);
      v = fmaf(p.yd->scale, acc, hb[e]);
    } else {
#pragma unroll
      for (int i = 0; i < 7; ++i) acc = fmaf(p.coef[i], p.src[i + 1][e], acc);
      v = fmaf(p.scale, acc, p.src[0][e]);
    }
  }
  float hi, lo;
  if (p.passes == 3) { hi = tf32_rna(v); lo = tf32_rna(v - hi); }
  else { hi = v; lo = 0.0f; }
  uint8_t* g = reinterpret_cast<uint8_t*>(p.hbuf) + (size_t)(n >> 4) * p.unit_bytes;
  const int r = n & 15;
  uint32_t o, lo_step;
  if (kk < p.nfull * 32) {
    const int c = kk >> 5, k5 = kk & 31;
    o = (uint32_t)(c * kPieceBytes + (r >> 3) * 1024 + (r & 7) * 128 + (((k5 >> 2) ^ (r & 7)) << 4) + (k5 & 3) * 4);
    lo_step = (uint32_t)(kUR * 128);
  } else {
    const int kt = kk - p.nfull * 32, t = kt >> 3, k3 = kt & 7;
    o = (uint32_t)(p.nfull * kPieceBytes + t * kTailBytes + r * 32 + ((((k3 >> 2) & 1) ^ ((r >> 2) & 1)) << 4) + (k3 & 3) * 4);
    lo_step = (uint32_t)(kUR * 32);
  }
  *reinterpret_cast<float*>(g + o) = hi;
  *reinterpret_cast<float*>(g + o + lo_step) = lo;
}

}  // namespace fused

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
#define LRF_COUNT(ctx) do { if ((ctx)->capturing) (ctx)->captured++; else (ctx)->launches++; } while (0)

static size_t lrf_round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

bool lrf_shape(const lrnde_model* m, FusedShape* out) {
  if (!m->conv.empty() || m->layers.size() != 2) return false;
  if (m->input_act != ACT_IDENTITY) return false;
  const LayerInfo& L1 = m->layers[0];
  const LayerInfo& L2 = m->layers[1];
  if (L2.act != ACT_IDENTITY) return false;
  FusedShape s;
  s.D = m->D; s.H = L1.out; s.td = m->td; s.act = L1.act;
  s.Kaug = s.H + s.td + 1;
  if (s.Kaug > 128) return false;
  s.KS = (s.Kaug + 7) / 8;
  s.nfull = s.KS / 4;
  s.ntail = s.KS % 4;
  s.n_mt = (s.D + 127) / 128;
  s.MT = (int)lrf_round_up((size_t)(s.D + s.n_mt - 1) / s.n_mt, 16);
  if (out) *out = s;
  return true;
}

static void lrf_set_attrs() {
  static bool attr_set = false;
  if (!attr_set) {
    LR_CUDA(cudaFuncSetAttribute(fused::chain_kernel<ACT_IDENTITY>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    LR_CUDA(cudaFuncSetAttribute(fused::chain_kernel<ACT_TANH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    LR_CUDA(cudaFuncSetAttribute(fused::chain_kernel<ACT_GELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    LR_CUDA(cudaFuncSetAttribute(fused::chain_kernel<ACT_SIGMOID>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    LR_CUDA(cudaFuncSetAttribute(fused::chain_kernel<ACT_RELU>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    LR_CUDA(cudaFuncSetAttribute(fused::kgemm_kernel<6, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    LR_CUDA(cudaFuncSetAttribute(fused::kgemm_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    LR_CUDA(cudaFuncSetAttribute(fused::kgemm_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    attr_set = true;
  }
}

bool FusedEngine::eligible(const lrnde_model* m) { return lrf_shape(m, nullptr); }

FusedEngine::FusedEngine(lrnde_ctx* c, const lrnde_model* mm, const float* p, int64_t b, int npasses)
    : ctx(c), m(mm), ps(p), B(b), passes(npasses) {
  lrf_shape(m, &sh);
  ntiles = (int)((B + fused::kNT - 1) / fused::kNT);
  nunits = (int)((B + 15) / 16);
  imgM = lrf_round_up((size_t)128 * sh.KS * 32, 1024);
  unit_bytes = (size_t)sh.nfull * fused::kPieceBytes + (size_t)sh.ntail * fused::kTailBytes;
  unit_bytes2 = (size_t)sh.nfull * 8192 + (size_t)sh.ntail * 2048;   // units of 16 samples x 2 row groups (combo attempts)
  nbuf = (2 * imgM + 2 * (size_t)fused::kNT * sh.KS * 64 + 2048 <= 226 * 1024) ? 2 : 1;
  Mimg = (float*)ctx->alloc(2 * imgM);
  Mplain = (float*)ctx->alloc(sizeof(float) * 128 * 128);
  hbuf = (float*)ctx->alloc((size_t)ntiles * 4 * unit_bytes);
  // kgemm launch geometry: one cluster = the n_mt feature tiles of the same samples (operand pieces multicast)
  // (measured at 8192 samples: 7-CTA clusters fit 15 at a time = 105 SMs, 80 us per attempt; 147 independent CTAs
  // re-reading the operand images from L2: 64 us -- so the multicast path is opt-in until its ring is deeper)
  cluster = (sh.n_mt <= 8 && getenv("LRNDE_FUSED_CLUSTER")) ? sh.n_mt : 1;
  ring = 8;
  lrf_set_attrs();
  int maxc = 0;
  if (cluster > 1) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(cluster * 64);
    cfg.blockDim = dim3(fused::kThreads);
    cfg.dynamicSmemBytes = 2 * lrf_round_up(unit_bytes, 1024) + 1024;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&maxc, fused::kgemm_kernel<6, false>, &cfg);
    if (e != cudaSuccess || maxc < 1) { cudaGetLastError(); cluster = 1; }
  }
  if (cluster == 1) maxc = 148 / sh.n_mt > 0 ? 148 / sh.n_mt : 1;
  if (const char* e = getenv("LRNDE_FUSED_MAXC")) maxc = std::max(1, atoi(e));
  const int rounds = (nunits + maxc - 1) / maxc;
  nclusters = (nunits + rounds - 1) / rounds;
  if (getenv("LRNDE_FUSED_VERBOSE")) fprintf(stderr, "[lrnde fused] B=%lld units=%d cluster=%d max clusters=%d -> %d clusters, nbuf=%d\n", (long long)B, nunits, cluster, maxc, nclusters, nbuf);
}

FusedEngine::~FusedEngine() {
  ctx->release(Mimg);
  ctx->release(Mplain);
  ctx->release(hbuf);
}

void lrf_mz_plain(lrnde_ctx* ctx, const lrnde_model* m, const float* ps, float* out) {
  FusedShape sh;
  lrf_shape(m, &sh);
  const LayerInfo& L1 = m->layers[0];
  const LayerInfo& L2 = m->layers[1];   // [D x (H + td)] followed by the bias: one [D x Kaug] column-major block
  fused::mz_plain_kernel<<<128, 512, 0, ctx->stream>>>(ps + L1.w_off, ps + L2.w_off, sh.D, sh.H, sh.Kaug, out);
  LRF_COUNT(ctx);
  LR_CUDA(cudaGetLastError());
}

void FusedEngine::prepare(float* mz_plain_out) {
  float* plain = mz_plain_out ? mz_plain_out : Mplain;
  lrf_mz_plain(ctx, m, ps, plain);
  fused::mz_image_kernel<<<sh.KS * 8, 128, 0, ctx->stream>>>(plain, sh.Kaug, sh.nfull, Mimg, (uint32_t)imgM, passes);
  LRF_COUNT(ctx);
  LR_CUDA(cudaGetLastError());
}

static void lrf_launch_chain(FusedEngine& E, SolveDev* S, const LinComb* single, const LinComb* single_out, const int* done,
                             int write_z) {
  const FusedShape& sh = E.sh;
  const LayerInfo& L1 = E.m->layers[0];
  fused::ChainP cp;
  memset(&cp, 0, sizeof(cp));
  cp.S = S; cp.single = single; cp.single_out = single_out; cp.done = done;
  cp.Mimg = E.Mimg; cp.imgM = (uint32_t)E.imgM;
  cp.w1t = sh.td ? E.ps + L1.w_off + (size_t)sh.D * sh.H : nullptr;
  cp.b1 = E.ps + L1.b_off;
  cp.hbuf = E.hbuf; cp.unit_bytes = (uint32_t)E.unit_bytes;
  cp.B = (int)E.B; cp.H = sh.H; cp.td = sh.td; cp.KS = sh.KS; cp.nfull = sh.nfull; cp.ntail = sh.ntail;
  cp.passes = E.passes; cp.nbuf = E.nbuf; cp.write_z = write_z;
  cp.combo = (!single && E.combo()) ? 1 : 0;
  if (cp.combo) cp.unit_bytes = (uint32_t)E.unit_bytes2;
  const size_t smem_c = 2 * E.imgM + (size_t)E.nbuf * fused::kNT * sh.KS * 64 + 1024;
  cudaStream_t st = E.ctx->stream;
  switch (sh.act) {
    case ACT_TANH: fused::chain_kernel<ACT_TANH><<<E.ntiles, fused::kThreads, smem_c, st>>>(cp); break;
    case ACT_GELU: fused::chain_kernel<ACT_GELU><<<E.ntiles, fused::kThreads, smem_c, st>>>(cp); break;
    case ACT_SIGMOID: fused::chain_kernel<ACT_SIGMOID><<<E.ntiles, fused::kThreads, smem_c, st>>>(cp); break;
    case ACT_RELU: fused::chain_kernel<ACT_RELU><<<E.ntiles, fused::kThreads, smem_c, st>>>(cp); break;
    default: fused::chain_kernel<ACT_IDENTITY><<<E.ntiles, fused::kThreads, smem_c, st>>>(cp); break;
  }
  LRF_COUNT(E.ctx);
  LR_CUDA(cudaGetLastError());
}

static void lrf_launch_kgemm(FusedEngine& E, SolveDev* S, const LinComb* single, const LinComb* single_out, const int* done) {
  const FusedShape& sh = E.sh;
  const LayerInfo& L2 = E.m->layers[1];
  fused::KgemmP kp;
  memset(&kp, 0, sizeof(kp));
  kp.S = S; kp.single = single; kp.single_out = single_out; kp.done = done;
  kp.W2a = E.ps + L2.w_off; kp.hbuf = E.hbuf; kp.unit_bytes = (uint32_t)E.unit_bytes;
  kp.B = (int)E.B; kp.D = sh.D; kp.MT = sh.MT; kp.n_mt = sh.n_mt; kp.Kaug = sh.Kaug; kp.KS = sh.KS; kp.nfull = sh.nfull;
  kp.dbg = getenv("LRNDE_KG_DBG") ? atoi(getenv("LRNDE_KG_DBG")) : 0;
  kp.lean = (!single && E.lean) ? 1 : 0;
  kp.add_base = single ? E.add_base : nullptr;
  const bool combo = !single && E.combo();
  if (combo) kp.unit_bytes = (uint32_t)E.unit_bytes2;
  kp.ntail = sh.ntail; kp.passes = E.passes; kp.nunits = E.nunits; kp.nclusters = E.nclusters; kp.ring = E.ring;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(sh.n_mt * E.nclusters);
  cfg.blockDim = dim3(fused::kThreads);
  cfg.dynamicSmemBytes = 2 * lrf_round_up(combo ? E.unit_bytes2 : E.unit_bytes, 1024) + 1024;
  cfg.stream = E.ctx->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = E.cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  if (combo) LR_CUDA(cudaLaunchKernelEx(&cfg, fused::kgemm_kernel<2, false>, kp));
  else LR_CUDA(cudaLaunchKernelEx(&cfg, fused::kgemm_kernel<6, false>, kp));
  LRF_COUNT(E.ctx);
  LR_CUDA(cudaGetLastError());
}

void lrf_launch_kgemm_adj(lrnde_ctx* ctx, SolveDev* S, const FusedShape& sh, const float* W1, const float* hbuf,
                          size_t unit_bytes, int64_t B, int passes, int nunits, int nclusters) {
  lrf_set_attrs();
  fused::KgemmP kp;
  memset(&kp, 0, sizeof(kp));
  kp.S = S;
  kp.W2a = W1; kp.hbuf = hbuf; kp.unit_bytes = (uint32_t)unit_bytes;
  kp.B = (int)B; kp.D = sh.D; kp.MT = sh.MT; kp.n_mt = sh.n_mt; kp.Kaug = sh.H; kp.KS = sh.KS; kp.nfull = sh.nfull;
  kp.ntail = sh.ntail; kp.passes = passes; kp.nunits = nunits; kp.nclusters = nclusters; kp.ring = 2;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(sh.n_mt * nclusters);
  cfg.blockDim = dim3(fused::kThreads);
  cfg.dynamicSmemBytes = 2 * lrf_round_up(unit_bytes, 1024) + 1024;
  cfg.stream = ctx->stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  LR_CUDA(cudaLaunchKernelEx(&cfg, fused::kgemm_kernel<2, true>, kp));
  LRF_COUNT(ctx);
  LR_CUDA(cudaGetLastError());
}

static void lrf_launch(FusedEngine& E, SolveDev* S, const LinComb* single, const LinComb* single_out, const int* done,
                       int write_z) {
  lrf_launch_chain(E, S, single, single_out, done, write_z);
  lrf_launch_kgemm(E, S, single, single_out, done);
}

void FusedEngine::dense_output(SolveDev* S, const float* const* harr, const float* coef, float scale, const float* xbase,
                               const LinComb* out_desc) {
  fused::InterpP ip;
  memset(&ip, 0, sizeof(ip));
  for (int i = 0; i < 8; ++i) ip.src[i] = harr[i];
  for (int i = 0; i < 7; ++i) ip.coef[i] = coef[i];
  ip.scale = scale; ip.hbuf = hbuf; ip.unit_bytes = (uint32_t)unit_bytes;
  ip.B = (int)B; ip.Kaug = sh.Kaug; ip.KS = sh.KS; ip.nfull = sh.nfull; ip.passes = passes;
  fused::interp_image_kernel<<<(int)((B + 1) / 2), 256, 0, ctx->stream>>>(ip);
  LRF_COUNT(ctx);
  add_base = xbase;
  lrf_launch_kgemm(*this, S, out_desc, out_desc, nullptr);
  add_base = nullptr;
}

void FusedEngine::dense_output_dev(SolveDev* A, const LinComb* yd, const float* xbase, float* ybuf, LinComb* ydesc_out,
                                   const LinComb* out_desc) {
  fused::InterpP ip;
  memset(&ip, 0, sizeof(ip));
  ip.A = A; ip.yd = yd; ip.ydesc_out = ydesc_out; ip.ybase = ybuf;
  ip.hbuf = hbuf; ip.unit_bytes = (uint32_t)unit_bytes;
  ip.B = (int)B; ip.Kaug = sh.Kaug; ip.KS = sh.KS; ip.nfull = sh.nfull; ip.passes = passes;
  fused::interp_image_kernel<<<(int)((B + 1) / 2), 256, 0, ctx->stream>>>(ip);
  LRF_COUNT(ctx);
  add_base = xbase;
  lrf_launch_kgemm(*this, A, out_desc, out_desc, &A->failed);
  add_base = nullptr;
}

void FusedEngine::step_chain(SolveDev* S, int write_z) { lrf_launch_chain(*this, S, nullptr, nullptr, nullptr, write_z); }
void FusedEngine::step_kgemm(SolveDev* S) { lrf_launch_kgemm(*this, S, nullptr, nullptr, nullptr); }

void FusedEngine::step(SolveDev* S, int write_z) { lrf_launch(*this, S, nullptr, nullptr, nullptr, write_z); }

void FusedEngine::eval(SolveDev* S, const LinComb* in, const LinComb* out, const int* done, int write_z) {
  lrf_launch(*this, S, in, out, done, write_z);
}
