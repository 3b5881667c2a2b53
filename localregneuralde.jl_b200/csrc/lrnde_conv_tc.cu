// Tensor-core engine of the conv dynamics (see lrnde_conv_tc.h): hand-written tcgen05 / TMEM / bulk-copy code for
// sm_100a.  Reference arithmetic: experiments/src/construct.jl:212-218 (the Chain), src/layers/common.jl:19-33
// (TDChain time channel), NNlib conv = true convolution (flipped kernel), Lux BatchNorm in training mode.
#include "lrnde_conv_tc.h"

#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "lrnde_fused_dev.cuh"

#define LCT_COUNT(ctx) do { if ((ctx)->capturing) (ctx)->captured++; else (ctx)->launches++; } while (0)

// in-kernel timeline of CTA 0 (LRNDE_CT_DBG & 32): rows of 16 clock64 stamps
__device__ long long g_ct_trace[8 * 16];
#define CTRACE(row, it) do { if (ctrace_on && (it) < 16) g_ct_trace[(row) * 16 + (it)] = clock64(); } while (0)
extern "C" int lrnde_debug_trace_convtc(long long* out, int n) {
  return (int)cudaMemcpyFromSymbol(out, g_ct_trace, sizeof(long long) * (size_t)n);
}

namespace convtc {
using namespace umma;
using fused::mma;
using fused::tmem_ld16;

__device__ __forceinline__ float lr_lincomb_at_tc(const LinComb& d, size_t i) {   // lr_lincomb_at of lrnde_kernels.cuh
  float inner = 0.0f;
  for (int k = 0; k < d.n; ++k) inner = fmaf(d.coef[k], d.src[k][i], inner);
  const float b = d.base ? d.base[i] : 0.0f;
  return d.n ? fmaf(d.scale, inner, b) : b;
}

constexpr int kTiles = 4;                                  // M tiles per group
constexpr int kRun = kCtGroup + 2 * kCtGuard;              // positions of a patch run (584)
constexpr int kRunB = kRun * 16;                           // bytes of one channel chunk of a run
constexpr int kStages = 3;
constexpr int kThreads = 192;                              // warp 0: bulk copies, warp 1: tcgen05.mma, warps 2..5: epilogue
// K-major, no swizzle ("interleave"): core matrix = 8 rows x 16 bytes, rows 16 bytes apart, 8-row groups SBO = 128 B
// apart (= uniform 16-byte row stride), the two 16-byte K chunks of an MMA LBO apart (LBO travels in the low word)
constexpr uint32_t kHiNone = (128u >> 4) | (1u << 14);
__device__ __forceinline__ uint32_t dlo(uint32_t saddr, uint32_t lbo_bytes) {
  return ((saddr >> 4) & 0x3FFFu) | ((lbo_bytes >> 4) << 16);
}
__host__ __device__ constexpr int wstage_bytes(int NOUT) { return 2 * 9 * 2 * NOUT * 16; }
__host__ __device__ constexpr int stage_bytes(int NOUT) { return 4 * kRunB + wstage_bytes(NOUT); }

// ---------------------------------------------------------------------------------------------------------
// pack: one pass over an activation, everything element-wise between two convolutions, hi / lo split
// ---------------------------------------------------------------------------------------------------------
// derivative of the activation in the pullback epilogue: ONE out-of-line body (sixteen call sites per loop: inlining the
// five-way switch with its tanhf / expf sixty-four times made the kernel instruction-fetch bound, 100 -> 750 us)
// gelu (tanh form, the Lux default of the cifar10 config) inline: sixteen independent chains per loop body
__device__ __forceinline__ float gelu_d(float x) {   // tanh through exp: absolute error ~2e-7
  const float inner = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  const float th = 1.0f - __fdividef(2.0f, __expf(2.0f * inner) + 1.0f);
  const float dinner = 0.7978845608028654f * (1.0f + 3.0f * 0.044715f * x * x);
  return 0.5f * (1.0f + th) + 0.5f * x * (1.0f - th * th) * dinner;
}
__device__ __forceinline__ float gelu_f(float x) {
  const float inner = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  return x * (1.0f - __fdividef(1.0f, __expf(2.0f * inner) + 1.0f));   // 0.5 x (1 + tanh(inner)) = x (1 - 1 / (e^{2 inner} + 1))
}
__device__ __noinline__ float dact_call(int act, float x) { return act == ACT_GELU ? gelu_d(x) : lr_dact(act, x); }

__device__ __noinline__ float act_call(int act, float x) { return act == ACT_GELU ? gelu_f(x) : lr_act(act, x); }

struct PackK {
  ConvTcPackP p;
  int Wd, Ht, PW, PH, IMG;
  long NPA;
};

// one block = one image row (b, y) and min(8, C / 4) warps; warp w owns the channel chunks c4 = w, w + nw, ...; lane = x
// (two passes when Wd > 32).
// The halo positions of (F) are zero from the allocation on and no pack ever writes them.
__global__ void __launch_bounds__(256) pack_kernel(PackK k) {
  const ConvTcPackP& p = k.p;
  if (p.done && *p.done) return;
  __shared__ LinComb xd;
  if (p.xdesc) {
    if (threadIdx.x == 0) xd = *p.xdesc;
    __syncthreads();
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int b = blockIdx.x / k.Ht, y = blockIdx.x % k.Ht;
  float* side = p.side_to_desc_dst ? xd.dst : p.side;
  const int C4 = p.C >> 2;
  const size_t HW = (size_t)k.Wd * k.Ht;
  for (int x = lane; x < k.Wd; x += 32) {
    const size_t pos = (size_t)kCtGuard + (size_t)b * k.IMG + (size_t)(y + 1) * k.PW + (x + 1);
    const size_t pix = x + (size_t)k.Wd * y;
#pragma unroll 1
    for (int c4a = warp; c4a < C4; c4a += 2 * nw) {
      // two chunks per pass (c4a, c4a + nw): every load of the pass is issued before the first out-of-line activation call
      float v[2][4], g[2][4];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int c4 = c4a + nw * u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const size_t idx = pix + HW * ((size_t)c4 * 4 + j + (size_t)p.C * b);
          const bool ok = c4 < C4;
          v[u][j] = !ok ? 0.0f : (p.xdesc ? lr_lincomb_at_tc(xd, idx) : __ldcg(p.X + idx));
          g[u][j] = (ok && p.bwd_g) ? __ldcg(p.bwd_g + idx) : 0.0f;
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int c4 = c4a + nw * u;
        if (c4 >= C4) break;
        float hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = c4 * 4 + j;
          const size_t idx = pix + HW * ((size_t)c + (size_t)p.C * b);
          float r = v[u][j];
          if (p.xdesc && side) side[idx] = r;
          if (p.bwd_g) {
            if (p.bwd_coef) {
              const float a = p.in_ab[c], bb = p.in_ab[p.C + c];
              const float pre = fmaf(a, r, bb);
              const float gh = g[u][j] * (p.bwd_act == ACT_GELU ? gelu_d(pre) : dact_call(p.bwd_act, pre));
              const float xhat = (r - p.bwd_stat[c]) * p.bwd_stat[p.C + c];
              r = a * (gh - p.bwd_coef[c] - xhat * p.bwd_coef[p.C + c]);
            } else r = g[u][j] * dact_call(p.bwd_act, r);
          } else {
            if (p.in_ab) r = fmaf(p.in_ab[c], r, p.in_ab[p.C + c]);
            if (p.in_act == ACT_GELU) r = gelu_f(r);
            else if (p.in_act != ACT_IDENTITY) r = act_call(p.in_act, r);
          }
          hi[j] = tf32_rna(r);
          lo[j] = tf32_rna(r - hi[j]);
          if (p.Pv) p.Pv[idx] = r;
          if (p.rowsum) {   // (Wd == 32) sum over the image row, its first and last pixel: the time channel's weight gradient
            float sum = r;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            const float last = __shfl_sync(0xffffffffu, r, 31);
            if (lane == 0) {
              float* dst = p.rowsum + ((size_t)blockIdx.x * p.C + c) * 3;
              dst[0] = sum; dst[1] = r; dst[2] = last;
            }
          }
        }
        if (p.Fhi) {
          const size_t o = ((size_t)c4 * (size_t)k.NPA + pos) * 4;
          *reinterpret_cast<float4*>(p.Fhi + o) = make_float4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<float4*>(p.Flo + o) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// weight image: logical convolution out[o, n] = sum_tap sum_k Wl[tap][k][n] in[o + d_tap, k], d_tap = (tap / 3 - 1,
// tap % 3 - 1) = (dy, dx).  NNlib's flipped kernel: Wl[tap][ci][co] = w[2 - tap % 3, 2 - tap / 3, ci, co]; the
// data-gradient convolution: Wl[tap][k = co][n = ci] = w[tap % 3, tap / 3, ci, co].
// image element (s, part, tap, k4, n, j): channel k = 8 s + 4 k4 + j, hi (part 0) / lo (part 1)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float wl_at(const float* w, int CinTot, int Cout, int transposed, int tap, int k, int n) {
  if (!transposed) return w[(2 - tap % 3) + 3 * ((2 - tap / 3) + 3 * (k + (size_t)CinTot * n))];
  return w[(tap % 3) + 3 * ((tap / 3) + 3 * (n + (size_t)CinTot * k))];
}
__global__ void wpack_kernel(const float* w, int CinTot, int Cout, int transposed, int K, int Nreal, int NOUT, int td,
                             float* img, float* tsum) {
  const int nst = K / 8;
  const int per_stage = 2 * 9 * 2 * NOUT * 4;
  const int n_img = nst * per_stage;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_img; e += gridDim.x * blockDim.x) {
    int r = e;
    const int j = r & 3; r >>= 2;
    int n, k4, tap, part, s;
    if (NOUT == 16) {   // [stage][tap][k4][hi rows 0..15 | lo rows 16..31][4]: [B_hi; B_lo] is ONE B operand of N = 32
      n = r & 15; r >>= 4;
      part = r & 1; r >>= 1;
      k4 = r & 1; r >>= 1;
      tap = r % 9; r /= 9;
      s = r;
    } else {            // [stage][hi | lo][tap][k4][NOUT rows][4]
      n = r % NOUT; r /= NOUT;
      k4 = r & 1; r >>= 1;
      tap = r % 9; r /= 9;
      part = r & 1; r >>= 1;
      s = r;
    }
    const int k = 8 * s + 4 * k4 + j;
    float v = 0.0f;
    if (n < Nreal) v = wl_at(w, CinTot, Cout, transposed, tap, k, n);
    const float hi = tf32_rna(v);
    img[e] = part ? tf32_rna(v - hi) : hi;
  }
  if (tsum && td && !transposed) {
    // T[ym][xm][n] = sum over the taps whose input pixel lies inside the image (bit dy + 1 of ym, dx + 1 of xm)
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < 64 * NOUT; e += gridDim.x * blockDim.x) {
      const int n = e % NOUT, xm = (e / NOUT) & 7, ym = e / (8 * NOUT);
      float acc = 0.0f;
      if (n < Nreal)
        for (int tap = 0; tap < 9; ++tap)
          if (((ym >> (tap / 3)) & 1) && ((xm >> (tap % 3)) & 1)) acc += wl_at(w, CinTot, Cout, 0, tap, CinTot - 1, n);
      tsum[e] = acc;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// the convolution
// ---------------------------------------------------------------------------------------------------------
struct ConvK {
  ConvTcP p;
  int Wd, Ht, PW, IMG, ngroups;
  long NP, NPA;
  int dbg;   // LRNDE_CT_DBG (profiling only): 1 = no output stores, 2 = no time-channel term
};

// sum over the warp of 32 per-lane values; lane L returns the total of value L (31 shuffles)
__device__ __forceinline__ float warp_transpose_sum(float (&v)[32], int lane) {
#pragma unroll
  for (int off = 16, n = 32; off >= 1; off >>= 1, n >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < n / 2; ++i) {
      const float send = up ? v[i] : v[i + n / 2];
      const float keep = up ? v[i + n / 2] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

constexpr int kConvThreads = 64 + 256;   // warp 0: bulk copies, warp 1: tcgen05.mma, warps 2..9: epilogue (two per TMEM lane quarter)
constexpr int kConvProd = 256;           // SRC = 1: warps 10..17 form the A operand in shared memory
// MODE 0: plain epilogue; 1: + (sum, sum of squares) of the raw output; 2: + the two sums of the BatchNorm pullback
// SRC 0: the A operand is the packed (F) image (bulk copies); 1: it is formed here from a stored [W,H,K,B] array: four
// producer warps read the patch run of a stage (584 positions x 8 channels), apply BatchNorm scale / shift + activation,
// split into tf32 hi / lo and write the (F) layout straight into the stage -- the pack kernel's work without its 151 MB
// round trip through HBM (the weights still arrive by bulk copy)
template <int NOUT, int MODE, int SRC>
__global__ void __launch_bounds__(kConvThreads + SRC * kConvProd, 1) conv_kernel(ConvK k) {
  const ConvTcP& p = k.p;
  if (p.done && *p.done) return;
  constexpr int STB = stage_bytes(NOUT), WSB = wstage_bytes(NOUT);
  // NOUT = 16 (bound by MMA issue / the A read, not by N): D[:, 0:16] += A_hi B_hi + A_lo B_hi, D[:, 16:32] += A_hi B_lo with
  // B' = [B_hi; B_lo] as one N = 32 operand: two MMAs per K-step instead of three; the epilogue adds the two halves
  constexpr int DCOL = (NOUT == 16) ? 32 : NOUT;
  constexpr int TCOLS = 2 * kTiles * DCOL;   // 512 (NOUT = 64) or 256 (NOUT = 16)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
  __shared__ uint64_t full_bar[kStages], empty_bar[kStages], acc_full[2], tmem_free[2];
  __shared__ uint32_t tmem_slot;
  // per-channel constants of the BatchNorm pullback (a, b, mean, invstd): with 222 KB of shared memory there is next to no
  // L1 left, every __ldg of them went to L2 (long-scoreboard stalls on the first FFMA were 90 % of the epilogue's samples)
  __shared__ float s_bw[MODE == 2 ? 4 : 1][MODE == 2 ? 64 : 1];
  __shared__ float s_in[SRC == 1 ? 2 : 1][SRC == 1 ? 64 : 1];   // SRC = 1: scale / shift of the input channels
  if (SRC == 1 && threadIdx.x < 64) {
    const int c = threadIdx.x;
    const bool ok = p.src_ab != nullptr && c < p.K;
    s_in[0][c] = ok ? p.src_ab[c] : 1.0f; s_in[1][c] = ok ? p.src_ab[p.K + c] : 0.0f;
  }
  if (MODE == 2 && threadIdx.x < 64) {
    const int c = threadIdx.x;
    const bool ok = c < p.Cout;
    s_bw[0][c] = ok ? p.bwd_ab[c] : 0.0f; s_bw[1][c] = ok ? p.bwd_ab[p.Cout + c] : 0.0f;
    s_bw[2][c] = ok ? p.bwd_stat[c] : 0.0f; s_bw[3][c] = ok ? p.bwd_stat[p.Cout + c] : 0.0f;
  }

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int nst = p.K >> 3;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], SRC == 1 ? 1u + kConvProd / 32 : 1u); mbar_init(&empty_bar[s], 1u); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1u); mbar_init(&tmem_free[s], 8u); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)),
                 "r"((uint32_t)TCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  // LRNDE_CT_DBG: 32 = trace CTA 0; + 64: only the K = 8 convolution into 64 channels; + 128: only K = 64 into 64
  const bool ctrace_on = (k.dbg & 32) && blockIdx.x == 0 && (!(k.dbg & 64) || (p.K == 8 && NOUT == 64)) &&
                         (!(k.dbg & 128) || (p.K == 64 && NOUT == 64));
  if (threadIdx.x == 0) CTRACE(0, 0);

  if (warp == 0) {
    // ---------------- bulk copies: per stage of 8 input channels, [hi k4=0 | hi k4=1 | lo k4=0 | lo k4=1] runs + weights
    uint32_t ks = 0;
    for (int g = blockIdx.x; g < k.ngroups; g += gridDim.x) {
      for (int s = 0; s < nst; ++s, ++ks) {
        const uint32_t slot = ks % kStages, ph = (ks / kStages) & 1u;
        mbar_wait(&empty_bar[slot], ph ^ 1u);
        if (elect_one_sync()) {
          uint8_t* dst = sm + (size_t)slot * STB;
          if (s == 0) CTRACE(1, (int)(ks / nst));
          mbar_arrive_expect_tx(&full_bar[slot], (uint32_t)(SRC == 1 ? WSB : STB));
          if (SRC == 0) {
#pragma unroll
            for (int part = 0; part < 2; ++part)
#pragma unroll
              for (int k4 = 0; k4 < 2; ++k4) {
                const float* src = (part ? p.Flo : p.Fhi) + ((size_t)(2 * s + k4) * (size_t)k.NPA + (size_t)g * kCtGroup) * 4;
                bulk_g2s(dst + (part * 2 + k4) * kRunB, src, (uint32_t)kRunB, &full_bar[slot]);
              }
          }
          bulk_g2s(dst + 4 * kRunB, p.Wimg + (size_t)s * WSB, (uint32_t)WSB, &full_bar[slot]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ---------------- tcgen05.mma issue: D[128 positions x NOUT] += A(patch shifted by the tap) x B(weights of the tap)
    const uint32_t idesc = make_idesc(128, NOUT), idesc32 = make_idesc(128, 32);
    uint32_t ks = 0;
    int it = 0;
    for (int g = blockIdx.x; g < k.ngroups; g += gridDim.x, ++it) {
      const int db = it & 1;
      if (it >= 2) { mbar_wait(&tmem_free[db], (uint32_t)(((it >> 1) - 1) & 1)); tc_fence_after(); }
      for (int s = 0; s < nst; ++s, ++ks) {
        const uint32_t slot = ks % kStages, ph = (ks / kStages) & 1u;
        mbar_wait(&full_bar[slot], ph);
        tc_fence_after();
        if (elect_one_sync()) {
          if (s == 0) CTRACE(2, it);
          const uint32_t base = smem_u32(sm + (size_t)slot * STB);
          const uint32_t wb = base + 4 * kRunB;
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap) {
            const int sh = (tap / 3 - 1) * k.PW + (tap % 3 - 1);
            const uint32_t bh = (NOUT == 16) ? dlo(wb + tap * (2 * 32 * 16), 32 * 16) : dlo(wb + tap * (2 * NOUT * 16), NOUT * 16);
            const uint32_t bl = dlo(wb + 9 * 2 * NOUT * 16 + tap * (2 * NOUT * 16), NOUT * 16);
            const uint32_t acc = (s > 0 || tap > 0) ? 1u : 0u;
#pragma unroll
            for (int i = 0; i < kTiles; ++i) {
              const uint32_t aoff = (uint32_t)((kCtGuard + i * 128 + sh) * 16);
              const uint32_t ah = dlo(base + aoff, kRunB), al = dlo(base + 2 * kRunB + aoff, kRunB);
              const uint32_t d = tmem_base + (uint32_t)(db * kTiles * DCOL + i * DCOL);
              if (NOUT == 16) {
                mma<0>(d, ah, bh, kHiNone, idesc32, acc);     // [A_hi B_hi | A_hi B_lo]
                mma<0>(d, al, bh, kHiNone, idesc, 1u);        // A_lo B_hi into the first half
              } else {
                mma<0>(d, al, bh, kHiNone, idesc, acc);
                mma<1>(d, ah, bl, kHiNone, idesc, 1u);
                mma<2>(d, ah, bh, kHiNone, idesc, 1u);
              }
            }
          }
          mma_commit(&empty_bar[slot]);
          if (s == nst - 1) { mma_commit(&acc_full[db]); CTRACE(3, it); }
        }
        __syncwarp();
      }
    }
  } else if (SRC == 1 && warp >= kConvThreads / 32) {
    // ---------------- A-operand producers: thread pt owns the run positions r = pt + 128 m (decoded once per group)
    const int pt = (int)threadIdx.x - kConvThreads;
    const uint32_t uIMG = (uint32_t)k.IMG, uPW = (uint32_t)k.PW;
    const size_t HW = (size_t)k.Wd * k.Ht;
    constexpr int NPOS = (kRun + kConvProd - 1) / kConvProd;   // 5
    uint32_t ks = 0;
    for (int g = blockIdx.x; g < k.ngroups; g += gridDim.x) {
      uint32_t poff[NPOS];
#pragma unroll
      for (int m = 0; m < NPOS; ++m) {
        const int r = pt + kConvProd * m;
        const long pp = (long)g * kCtGroup + r - kCtGuard;
        poff[m] = 0xFFFFFFFFu;
        if (r < kRun && pp >= 0 && pp < k.NP) {
          const uint32_t up = (uint32_t)pp, b = up / uIMG, r0 = up - b * uIMG;
          const uint32_t yh = r0 / uPW, xh = r0 - yh * uPW;
          if (xh >= 1 && xh <= (uint32_t)k.Wd && yh >= 1 && yh <= (uint32_t)k.Ht)
            poff[m] = (xh - 1) + (uint32_t)k.Wd * ((yh - 1) + (uint32_t)k.Ht * ((uint32_t)p.K * b));
        }
      }
      // software pipeline over the half stages (4 channels each): the loads of the next half are in flight while the
      // current one is transformed and stored (the kernel has no L1: every load is an L2 / HBM round trip)
      auto issue = [&](int c0, float (&v)[NPOS][4]) {
#pragma unroll
        for (int m = 0; m < NPOS; ++m)
#pragma unroll
          for (int j = 0; j < 4; ++j)
            v[m][j] = (poff[m] != 0xFFFFFFFFu) ? __ldcg(p.srcZ + poff[m] + (size_t)(c0 + j) * HW) : 0.0f;
      };
      auto store = [&](uint8_t* dst, int k4, int c0, const float (&v)[NPOS][4]) {
#pragma unroll
        for (int m = 0; m < NPOS; ++m) {
          const int r = pt + kConvProd * m;
          if (r < kRun) {
            float4 h4 = make_float4(0.0f, 0.0f, 0.0f, 0.0f), l4 = h4;
            if (poff[m] != 0xFFFFFFFFu) {
              float t[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                t[j] = fmaf(s_in[0][c0 + j], v[m][j], s_in[1][c0 + j]);
                if (p.src_act == ACT_GELU) t[j] = gelu_f(t[j]);
                else if (p.src_act != ACT_IDENTITY) t[j] = act_call(p.src_act, t[j]);
              }
              h4 = make_float4(tf32_rna(t[0]), tf32_rna(t[1]), tf32_rna(t[2]), tf32_rna(t[3]));
              l4 = make_float4(tf32_rna(t[0] - h4.x), tf32_rna(t[1] - h4.y), tf32_rna(t[2] - h4.z), tf32_rna(t[3] - h4.w));
            }
            *reinterpret_cast<float4*>(dst + k4 * kRunB + r * 16) = h4;
            *reinterpret_cast<float4*>(dst + (2 + k4) * kRunB + r * 16) = l4;
          }
        }
      };
      // four half stages in flight (2 x 9.3 KB each way would bound the stream at ~0.8 TB/s): rounds r = 2 s + k4, register
      // sets v0..v3 in rotation, two stages per trip (nst is even: the host uses this path only for K % 16 == 0)
      float v0[NPOS][4], v1[NPOS][4], v2[NPOS][4], v3[NPOS][4];
      const int nr = 2 * nst;
      auto c_of = [](int r) { return 4 * r; };
      issue(c_of(0), v0);
      issue(c_of(1), v1);
      issue(c_of(2), v2);
      for (int s = 0; s < nst; s += 2, ks += 2) {
        const int r = 2 * s;
        const uint32_t slot0 = ks % kStages, ph0 = (ks / kStages) & 1u;
        const uint32_t slot1 = (ks + 1) % kStages, ph1 = ((ks + 1) / kStages) & 1u;
        uint8_t* dst0 = sm + (size_t)slot0 * STB;
        uint8_t* dst1 = sm + (size_t)slot1 * STB;
        issue(c_of(r + 3), v3);
        mbar_wait(&empty_bar[slot0], ph0 ^ 1u);
        store(dst0, 0, c_of(r), v0);
        if (r + 4 < nr) issue(c_of(r + 4), v0);
        store(dst0, 1, c_of(r + 1), v1);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[slot0]);
        if (r + 5 < nr) issue(c_of(r + 5), v1);
        mbar_wait(&empty_bar[slot1], ph1 ^ 1u);
        store(dst1, 0, c_of(r + 2), v2);
        if (r + 6 < nr) issue(c_of(r + 6), v2);
        store(dst1, 1, c_of(r + 3), v3);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[slot1]);
      }
    }
  } else {
    // ---------------- epilogue: lane = position (coalesced along x), registers = 16 output channels at a time;
    // the two warps of a lane quarter take alternate 16-channel chunks (NOUT = 16: alternate tiles).  The position is
    // decoded once per tile in 32-bit arithmetic (a 64-bit division per chunk was half of the epilogue's latency).
    const int q = warp & 3, half = (warp - 2) >> 2;
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    float* out = p.ydesc ? p.ydesc->dst : p.Y;
    const float tval = p.tdesc ? p.tdesc->t : 0.0f;
    const size_t HW = (size_t)k.Wd * k.Ht;
    constexpr int NCC = NOUT / 16;
    constexpr int NCW = (NCC > 1) ? NCC / 2 : 1;     // chunks per warp
    const bool gelu_bwd = (MODE == 2) && p.bwd_act == ACT_GELU;
    const float oscale = p.out_scale;
    const bool no_store = (k.dbg & 1) != 0, no_ts = (k.dbg & 2) != 0;
    const uint32_t uIMG = (uint32_t)k.IMG, uPW = (uint32_t)k.PW, uNP = (uint32_t)k.NP;
    int it = 0;
    for (int g = blockIdx.x; g < k.ngroups; g += gridDim.x, ++it) {
      const int db = it & 1;
      float st[NCW][32];
#pragma unroll
      for (int h2 = 0; h2 < NCW; ++h2)
#pragma unroll
        for (int j = 0; j < 32; ++j) st[h2][j] = 0.0f;
      if (threadIdx.x == 64) CTRACE(4, it);
      mbar_wait(&acc_full[db], (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      if (threadIdx.x == 64) CTRACE(5, it);
#pragma unroll 1
      for (int i = (NCC > 1 ? 0 : half); i < ((k.dbg & 4) ? 0 : kTiles); i += (NCC > 1 ? 1 : 2)) {
        const uint32_t pp = (uint32_t)g * kCtGroup + (uint32_t)(i * 128 + q * 32 + lane);
        const uint32_t b = pp / uIMG, r0 = pp - b * uIMG;
        const uint32_t yh = r0 / uPW, xh = r0 - yh * uPW;
        const bool valid = pp < uNP && xh >= 1 && xh <= (uint32_t)k.Wd && yh >= 1 && yh <= (uint32_t)k.Ht;
        const int x = (int)xh - 1, y = (int)yh - 1;
        const size_t obase0 = (size_t)x + (size_t)k.Wd * ((size_t)y + (size_t)k.Ht * ((size_t)p.Cout * b));
        const int ym = (y > 0 ? 1 : 0) | 2 | (y < k.Ht - 1 ? 4 : 0), xm = (x > 0 ? 1 : 0) | 2 | (x < k.Wd - 1 ? 4 : 0);
#pragma unroll
        for (int h2 = 0; h2 < NCW; ++h2) {
          const int cc = (NCC > 1) ? half + 2 * h2 : 0;
          const size_t obase = obase0 + (size_t)(cc * 16) * HW;
          const float* ts = (p.tsum && !no_ts) ? p.tsum + (ym * 8 + xm) * NOUT + cc * 16 : nullptr;
          float v[16];
          tmem_ld16(tlane + (uint32_t)(db * kTiles * DCOL + i * DCOL + cc * 16), v);
          if (NOUT == 16) {
            float v2[16];
            tmem_ld16(tlane + (uint32_t)(db * kTiles * DCOL + i * DCOL + 16), v2);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] += v2[j];
          }
          if (valid && !(k.dbg & 16)) {
            // every load of the chunk is issued before the arithmetic (one basic block: no per-element branches; the
            // channels >= Cout of a partial chunk have zero weights, so their accumulators and table entries are zero)
            float* o = out + obase;
            float tsv[16], zz[16];
            if (MODE != 2) {
#pragma unroll
              for (int j = 0; j < 16; ++j) tsv[j] = ts ? __ldg(ts + j) : 0.0f;
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) zz[j] = __ldcg(p.bwd_z + obase0 + (size_t)min(cc * 16 + j, p.Cout - 1) * HW);
            }
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float r = v[j];
              if (MODE != 2) r = fmaf(tval, tsv[j], r);
              if (MODE == 1) {
                st[h2][j] += r;
                st[h2][16 + j] = fmaf(r, r, st[h2][16 + j]);
              } else if (MODE == 2) {
                const int c = cc * 16 + j;
                const float pre = fmaf(s_bw[0][c], zz[j], s_bw[1][c]);
                const float gh = r * (gelu_bwd ? gelu_d(pre) : dact_call(p.bwd_act, pre));
                st[h2][j] += gh;
                st[h2][16 + j] = fmaf(gh, (zz[j] - s_bw[2][c]) * s_bw[3][c], st[h2][16 + j]);
              }
              if (cc * 16 + j < p.Cout && !no_store) o[(size_t)j * HW] = r * oscale;
            }
          }
        }
      }
      // every accumulator of this warp has been read: the MMAs of the group after next may overwrite the buffer
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_free[db]);
      if (threadIdx.x == 64) CTRACE(6, it);
      if (MODE != 0 && !(k.dbg & 8)) {
#pragma unroll
        for (int h2 = 0; h2 < NCW; ++h2) {
          const int cc = (NCC > 1) ? half + 2 * h2 : 0;
          const float tot = warp_transpose_sum(st[h2], lane);
          const int co = cc * 16 + (lane & 15);
          // NOUT = 16: the two warps of a quarter saw different tiles of the same channels -> separate partial rows
          const size_t row = (NCC > 1) ? ((size_t)g * 4 + q) : (((size_t)g * 4 + q) * 2 + half);
          if (co < p.Cout) reinterpret_cast<float*>(p.stat_part + row * p.Cout + co)[lane >> 4] = tot;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TCOLS) : "memory");
  }
}


// ---------------------------------------------------------------------------------------------------------
// weight gradient (Wd == 32): see lrnde_conv_tc.h
//   step i of a CTA = P row t = t0 + i (flattened (image, row) index): tensor-map TMA delivers the fp32 tile [64 channels]
//   [32 pixels] (K-major SWIZZLE_128B) of the STORED array; warps 2 and 3 (a thread per channel row) apply BatchNorm +
//   activation, split into tf32 hi / lo and write the tile and its two dx-shifted, zero-padded copies (a TMA box cannot
//   start at a pixel that is not a multiple of 16 bytes: measured, scratch/mb/tma_test.cu); warps 4 and 5 do the same
//   (without shifts) for the Q rows; Q rows t - 1, t, t + 1 of the same image are the B operands of the taps dy = +1, 0, -1.
//   A (M = 128) = two adjacent P tiles: (dx = -1, dx = 0) and (dx = +1, whatever follows: rows 64..127 are ignored)
//   TMEM: accumulator (dy, mt) at columns (dy * 2 + mt) * NB, zeroed by the epilogue warps before the first MMA.
//   The three Q rows of a step are ONE B operand of N = 3 NB -- their tiles are adjacent in a ring of 4 slots whose first
//   two are mirrored behind the last -- and the accumulators of an M tile are adjacent, (mt * 3 + n) * NB with n = 1 - dy:
//   a third of the MMAs and of the A-operand reads (the separate N = NB MMAs were bound by shared-memory bandwidth).
// ---------------------------------------------------------------------------------------------------------
constexpr int kWgQS = 4;                            // Q ring slots (+ 2 mirrors)
constexpr int kWgRS = 4;                            // raw P tiles in flight (TMA staging ring: the TMA latency of ~1-2 us is not
                                                    // part of the MMA -> operand preparation -> MMA chain of a P slot)
__host__ __device__ constexpr int wg_xs(int NB) { return NB == 64 ? 2 : 3; }   // P ring slots (the mirrored Q ring of 64-row tiles takes 96 KB)
constexpr int kWgPTile = 64 * 128;                  // one P tile (hi or lo)
constexpr int kWgXSlot = 6 * kWgPTile;              // [hi dx-1 | hi dx0 | hi dx+1 | lo dx-1 | lo dx0 | lo dx+1]

struct WgK {
  int Ht, nrows_total, rows_per_cta;
  int Pc, Qc, swapped, CinTot, Cd;
  float* part; size_t block;
  const int* done;
  int dbg;   // LRNDE_WG_DBG (bring-up): 1 = no TMA / MMA, 2 = TMA but no MMA
  const float* p_ab; const float* q_ab; int p_act, q_act;   // v <- act(a[c] v + b[c]) applied to the rows of P / Q
};

// one operand row of 32 pixels: BatchNorm scale / shift + activation (rows >= C of the box are zero fill: left alone)
// (a, b: the row's scale / shift, fetched ONCE per thread -- the kernel has next to no L1 and a __ldg per step went to L2)
__device__ __forceinline__ void wg_transform(float (&f)[34], bool affine, float a, float b, int act, int c, int C) {
  if (c >= C) return;
  if (affine) {
#pragma unroll
    for (int x = 1; x <= 32; ++x) f[x] = fmaf(a, f[x], b);
  }
  if (act == ACT_GELU) {
#pragma unroll
    for (int x = 1; x <= 32; ++x) f[x] = gelu_f(f[x]);
  } else if (act != ACT_IDENTITY) {
#pragma unroll
    for (int x = 1; x <= 32; ++x) f[x] = act_call(act, f[x]);
  }
}

__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

template <int NB>
__global__ void __launch_bounds__(kThreads, 1) wgrad_kernel(const __grid_constant__ CUtensorMap mP, const __grid_constant__ CUtensorMap mQ, WgK k) {
  if (k.done && *k.done) return;
  constexpr int QT = NB * 128;           // one Q tile (hi or lo)
  constexpr bool STACK = true;
  constexpr int kWgXS = wg_xs(NB);
  constexpr int QSLOT = 2 * QT;
  // Q ring: STACK: [hi of slots 0..5 | lo of slots 0..5] (slots 4, 5 mirror 0, 1); otherwise [hi | lo] per slot
  auto q_hi = [&](uint8_t* smQ_, int slot) { return STACK ? smQ_ + (size_t)slot * QT : smQ_ + (size_t)slot * QSLOT; };
  auto q_lo = [&](uint8_t* smQ_, int slot) { return STACK ? smQ_ + (size_t)(6 + slot) * QT : smQ_ + (size_t)slot * QSLOT + QT; };
  auto acc_col = [&](int dyi, int mt) { return STACK ? (mt * 3 + (2 - dyi)) * NB : (dyi * 2 + mt) * NB; };
  constexpr uint32_t TCOLS = (6 * NB <= 128) ? 128u : 512u;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* smQ = sm + kWgXS * kWgXSlot;
  uint8_t* smR = smQ + 12 * QT;          // raw P tiles (after the Q ring: the ignored rows of the last A tile read into the Q ring)
  __shared__ uint64_t r_full[kWgRS], r_empty[kWgRS], s_full[kWgXS], x_empty[kWgXS], q_full[kWgQS], qs_full[kWgQS], q_empty[kWgQS], acc_done,
      acc_zero;
  __shared__ uint32_t tmem_slot;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int t0 = blockIdx.x * k.rows_per_cta;
  const int n = min(k.rows_per_cta, k.nrows_total - t0);   // >= 1 by construction of the grid

  if (threadIdx.x == 0) {
    for (int s = 0; s < kWgXS; ++s) { mbar_init(&s_full[s], 2u); mbar_init(&x_empty[s], 1u); }
    for (int s = 0; s < kWgRS; ++s) { mbar_init(&r_full[s], 1u); mbar_init(&r_empty[s], 2u); }
    for (int s = 0; s < kWgQS; ++s) { mbar_init(&q_full[s], 1u); mbar_init(&qs_full[s], 2u); mbar_init(&q_empty[s], 1u); }
    mbar_init(&acc_done, 1u);
    mbar_init(&acc_zero, 4u);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(TCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_slot;
  const bool ctrace_on = (k.dbg & 32) && blockIdx.x == 0 && (!(k.dbg & 64) || NB == 64);   // LRNDE_WG_DBG=32 (+64: only NB = 64)
  if (threadIdx.x == 0) CTRACE(0, 0);

  if (k.dbg & 1) {
    if (warp >= 2) {
      float* out = k.part + (size_t)blockIdx.x * k.block;
      for (size_t e = threadIdx.x - 64; e < k.block; e += 128) out[e] = 0.0f;
    }
  } else if (warp == 0) {
    // ---------------- TMA: Q row j <-> flattened row u = t0 - 1 + j; P row i <-> t0 + i
    auto load_q = [&](int j) {
      const int slot = j % kWgQS;
      if (j >= kWgQS) mbar_wait(&q_empty[slot], (uint32_t)(((j / kWgQS) - 1) & 1));
      if (elect_one_sync()) {
        const int u = t0 - 1 + j;
        // rows in front of / behind the batch: coordinates outside the tensor (zero fill); they are never multiplied
        const int b = (u < 0) ? -1 : u / k.Ht, y = (u < 0) ? 0 : u % k.Ht;
        uint8_t* dst = q_hi(smQ, slot);
        mbar_arrive_expect_tx(&q_full[slot], (uint32_t)QT);
        tma_load_4d(dst, &mQ, 0, y, 0, b, &q_full[slot]);
      }
      __syncwarp();
    };
    load_q(0);
    load_q(1);
    for (int i = 0; i < n; ++i) {
      const int slot = i % kWgRS;
      if (i >= kWgRS) mbar_wait(&r_empty[slot], (uint32_t)(((i / kWgRS) - 1) & 1));
      if (elect_one_sync()) {
        const int t = t0 + i, b = t / k.Ht, y = t % k.Ht;
        CTRACE(1, i);
        mbar_arrive_expect_tx(&r_full[slot], (uint32_t)kWgPTile);
        tma_load_4d(smR + (size_t)slot * kWgPTile, &mP, 0, y, 0, b, &r_full[slot]);
      }
      __syncwarp();
      load_q(i + 2);
    }
  } else if (warp == 1) {
    // ---------------- tcgen05.mma issue
    const uint32_t idesc1 = make_idesc(128, NB), idesc2 = make_idesc(128, 2 * NB), idesc3 = make_idesc(128, 3 * NB);
    int q_arrived = 0;
    mbar_wait(&acc_zero, 0u);
    tc_fence_after();
    for (int i = 0; i < n; ++i) {
      const int xs = i % kWgXS;
      mbar_wait(&s_full[xs], (uint32_t)((i / kWgXS) & 1));
      while (q_arrived <= i + 2) {
        mbar_wait(&qs_full[q_arrived % kWgQS], (uint32_t)((q_arrived / kWgQS) & 1));
        ++q_arrived;
      }
      tc_fence_after();
      const int yy = (t0 + i) % k.Ht;
      if (elect_one_sync() && !(k.dbg & 2)) {
        CTRACE(4, i);
        const uint32_t xb = smem_u32(sm + (size_t)xs * kWgXSlot);
        if (STACK) {
          // Q rows i + n (n = 0, 1, 2 <-> dy = +1, 0, -1) of the same image, adjacent in the (mirrored) ring: one B operand
          const int n0 = (yy == 0) ? 1 : 0, n1 = (yy == k.Ht - 1) ? 1 : 2, cnt = n1 - n0 + 1;
          const int s0 = (i + n0) % kWgQS;
          const uint32_t qh = desc_lo(smem_u32(q_hi(smQ, s0))), ql = desc_lo(smem_u32(q_lo(smQ, s0)));
          const uint32_t idesc = cnt == 3 ? idesc3 : (cnt == 2 ? idesc2 : idesc1);
#pragma unroll
          for (int mt = 0; mt < 2; ++mt) {
            const uint32_t ah = desc_lo(xb + mt * 2 * kWgPTile), al = desc_lo(xb + 3 * kWgPTile + mt * 2 * kWgPTile);
            const uint32_t d = tmem_base + (uint32_t)((mt * 3 + n0) * NB);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              mma<0>(d, al + 2 * ks, qh + 2 * ks, fused::kHi128, idesc, 1u);
              mma<1>(d, ah + 2 * ks, ql + 2 * ks, fused::kHi128, idesc, 1u);
              mma<2>(d, ah + 2 * ks, qh + 2 * ks, fused::kHi128, idesc, 1u);
            }
          }
        } else {
#pragma unroll 1
          for (int dyi = 0; dyi < 3; ++dyi) {
            const int dy = dyi - 1, y = yy - dy;
            if (y < 0 || y >= k.Ht) continue;
            const int j = i + 1 - dy;
            const uint32_t qh = desc_lo(smem_u32(q_hi(smQ, j % kWgQS))), ql = desc_lo(smem_u32(q_lo(smQ, j % kWgQS)));
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) {
              const uint32_t ah = desc_lo(xb + mt * 2 * kWgPTile), al = desc_lo(xb + 3 * kWgPTile + mt * 2 * kWgPTile);
              const uint32_t d = tmem_base + (uint32_t)acc_col(dyi, mt);
#pragma unroll
              for (int ks = 0; ks < 4; ++ks) {
                mma<0>(d, al + 2 * ks, qh + 2 * ks, fused::kHi128, idesc1, 1u);
                mma<1>(d, ah + 2 * ks, ql + 2 * ks, fused::kHi128, idesc1, 1u);
                mma<2>(d, ah + 2 * ks, qh + 2 * ks, fused::kHi128, idesc1, 1u);
              }
            }
          }
        }
      }
      __syncwarp();
      if (elect_one_sync()) {
        mma_commit(&x_empty[xs]);
        mma_commit(&q_empty[i % kWgQS]);      // Q row j = i (flattened row t - 1): its last use was this step
        if (i == n - 1) mma_commit(&acc_done);
      }
      __syncwarp();
    }
  } else {
    // ---------------- accumulators <- 0 (every MMA accumulates; a tap whose rows never meet in this range stays zero)
    {
      const uint32_t tl = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
      float zero16[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) zero16[j] = 0.0f;
      for (int c0 = 0; c0 < 6 * NB; c0 += 16) fused::tmem_st16(tl + (uint32_t)c0, zero16);
      fused::tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_zero);
    }
    // ---------------- operand preparation in shared memory (conflict-free 16-byte accesses: a thread per 128-byte row)
    {
      const int tt = threadIdx.x - 64;
      if (tt < 64) {   // P rows: transform, split, and the dx = -1 / +1 copies
        const int c = tt, sw = c & 7;
        const bool aff = k.p_ab != nullptr && c < k.Pc;
        const float ta = aff ? __ldg(k.p_ab + c) : 1.0f, tb = aff ? __ldg(k.p_ab + k.Pc + c) : 0.0f;
        for (int i = 0; i < n; ++i) {
          const int slot = i % kWgXS, rs = i % kWgRS;
          mbar_wait(&r_full[rs], (uint32_t)((i / kWgRS) & 1));
          if (tt == 0) CTRACE(2, i);
          const uint8_t* raw = smR + (size_t)rs * kWgPTile + c * 128;
          uint8_t* base = sm + (size_t)slot * kWgXSlot + c * 128;
          float f[34];
          f[0] = 0.0f; f[33] = 0.0f;
#pragma unroll
          for (int qq = 0; qq < 8; ++qq) {
            const float4 r = *reinterpret_cast<const float4*>(raw + ((qq ^ sw) << 4));
            f[1 + 4 * qq] = r.x; f[2 + 4 * qq] = r.y; f[3 + 4 * qq] = r.z; f[4 + 4 * qq] = r.w;
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&r_empty[rs]);        // the raw tile is in registers: the TMA may refill the stage
          wg_transform(f, aff, ta, tb, k.p_act, c, k.Pc);
          float l[34];
#pragma unroll
          for (int x = 0; x < 34; ++x) { const float h = tf32_rna(f[x]); l[x] = tf32_rna(f[x] - h); f[x] = h; }
          // the MMAs that read this P slot two (three) steps ago are done
          if (i >= kWgXS) mbar_wait(&x_empty[slot], (uint32_t)(((i / kWgXS) - 1) & 1));
#pragma unroll
          for (int qq = 0; qq < 8; ++qq) {   // f[1 + x] = P[x]: the dx = -1 tile holds P[x - 1], the dx = +1 tile P[x + 1]
            const int o = (qq ^ sw) << 4;
            *reinterpret_cast<float4*>(base + o) = make_float4(f[4 * qq], f[4 * qq + 1], f[4 * qq + 2], f[4 * qq + 3]);
            *reinterpret_cast<float4*>(base + kWgPTile + o) = make_float4(f[4 * qq + 1], f[4 * qq + 2], f[4 * qq + 3], f[4 * qq + 4]);
            *reinterpret_cast<float4*>(base + 2 * kWgPTile + o) = make_float4(f[4 * qq + 2], f[4 * qq + 3], f[4 * qq + 4], f[4 * qq + 5]);
            *reinterpret_cast<float4*>(base + 3 * kWgPTile + o) = make_float4(l[4 * qq], l[4 * qq + 1], l[4 * qq + 2], l[4 * qq + 3]);
            *reinterpret_cast<float4*>(base + 4 * kWgPTile + o) = make_float4(l[4 * qq + 1], l[4 * qq + 2], l[4 * qq + 3], l[4 * qq + 4]);
            *reinterpret_cast<float4*>(base + 5 * kWgPTile + o) = make_float4(l[4 * qq + 2], l[4 * qq + 3], l[4 * qq + 4], l[4 * qq + 5]);
          }
          fence_proxy_async();
          __syncwarp();
          if (tt == 0) CTRACE(3, i);
          if (lane == 0) mbar_arrive(&s_full[slot]);
        }
      } else {         // Q rows: transform and split
        const int c = tt - 64, sw = c & 7;
        const bool aff = k.q_ab != nullptr && c < k.Qc;
        const float ta = aff ? __ldg(k.q_ab + c) : 1.0f, tb = aff ? __ldg(k.q_ab + k.Qc + c) : 0.0f;
        for (int j = 0; j < n + 2; ++j) {
          const int slot = j % kWgQS;
          mbar_wait(&q_full[slot], (uint32_t)((j / kWgQS) & 1));
          if (c < NB) {
            uint8_t* base = q_hi(smQ, slot) + c * 128;
            uint8_t* base_lo = q_lo(smQ, slot) + c * 128;
            const bool mirror = STACK && slot < 2;
            float f[34];
            f[0] = 0.0f; f[33] = 0.0f;
#pragma unroll
            for (int qq = 0; qq < 8; ++qq) {
              const float4 r = *reinterpret_cast<const float4*>(base + ((qq ^ sw) << 4));
              f[1 + 4 * qq] = r.x; f[2 + 4 * qq] = r.y; f[3 + 4 * qq] = r.z; f[4 + 4 * qq] = r.w;
            }
            wg_transform(f, aff, ta, tb, k.q_act, c, k.Qc);
#pragma unroll
            for (int qq = 0; qq < 8; ++qq) {
              const int o = (qq ^ sw) << 4;
              float4 h, lo4;
              h.x = tf32_rna(f[4 * qq + 1]); h.y = tf32_rna(f[4 * qq + 2]); h.z = tf32_rna(f[4 * qq + 3]); h.w = tf32_rna(f[4 * qq + 4]);
              lo4.x = tf32_rna(f[4 * qq + 1] - h.x); lo4.y = tf32_rna(f[4 * qq + 2] - h.y);
              lo4.z = tf32_rna(f[4 * qq + 3] - h.z); lo4.w = tf32_rna(f[4 * qq + 4] - h.w);
              *reinterpret_cast<float4*>(base + o) = h;
              *reinterpret_cast<float4*>(base_lo + o) = lo4;
              if (mirror) {
                *reinterpret_cast<float4*>(base + 4 * QT + o) = h;
                *reinterpret_cast<float4*>(base_lo + 4 * QT + o) = lo4;
              }
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&qs_full[slot]);
        }
      }
    }
    // ---------------- epilogue: accumulators -> part[split][Lux layout]
    const int q = warp & 3;
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    float* out = k.part + (size_t)blockIdx.x * k.block;
    if (k.CinTot > (k.swapped ? k.Qc : k.Pc))   // time channel: its entries of this split are zero (time_wgrad_kernel fills split 0)
      for (int e = threadIdx.x - 64; e < 9 * k.Cd; e += 128)
        out[(e % 9) + 9 * ((k.CinTot - 1) + (size_t)k.CinTot * (e / 9))] = 0.0f;
    if (threadIdx.x == 64) CTRACE(5, 0);
    mbar_wait(&acc_done, 0u);
    tc_fence_after();
    if (threadIdx.x == 64) CTRACE(5, 1);
    const int m = q * 32 + lane;
    const int pc = m & 63;
#pragma unroll 1
    for (int dyi = 0; dyi < 3; ++dyi) {
#pragma unroll 1
      for (int mt = 0; mt < 2; ++mt) {
        const int dxi = mt * 2 + (m >> 6);
#pragma unroll 1
        for (int cc = 0; cc < NB / 16; ++cc) {
          float v[16];
          tmem_ld16(tlane + (uint32_t)(acc_col(dyi, mt) + cc * 16), v);
          if (dxi < 3 && pc < k.Pc) {
            // normal : P = X (ci = pc), Q = Delta (co = qc): accumulator (dy, dx) is the tap itself -> w[2 - dx, 2 - dy]
            // swapped: P = Delta (co = pc) shifted by s, Q = X (ci = qc): tap d = -s -> w[dx, dy]
            const int kx = k.swapped ? dxi : 2 - dxi, ky = k.swapped ? dyi : 2 - dyi;
#pragma unroll
            for (int jn = 0; jn < 16; ++jn) {
              const int qc = cc * 16 + jn;
              if (qc < k.Qc) {
                const int ci = k.swapped ? qc : pc, co = k.swapped ? pc : qc;
                out[kx + 3 * (ky + 3 * (ci + (size_t)k.CinTot * co))] = v[jn];
              }
            }
          }
        }
      }
    }
  }
  if (threadIdx.x == 64) CTRACE(5, 2);
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TCOLS) : "memory");
  }
}

// weight gradient of the time channel: dW[tap, time, co] = t * sum_{b, pixels p with p + d_tap inside the image} Delta[p, co]
// = t * (T - [dy=-1] Top - [dy=+1] Bottom - [dx=-1] Left - [dx=+1] Right + the excluded corner), from the per-row
// (sum, first, last) triples the pack kernel wrote while it formed Delta.  grid (Cd, nsp): block (c, s) takes the rows
// r = s (mod nsp) and writes into split s of `part` (wgrad_kernel zero-fills the time entries of every split; nsp <= splits).
// Fixed partition, fixed-order tree: deterministic.
__global__ void __launch_bounds__(256) time_wgrad_kernel(const float* __restrict__ rowsum, int Cd, int Ht, int nrows_total,
                                                         const LinComb* tdesc, int Cx, int CinTot, float* part, size_t block,
                                                         const int* done) {
  if (done && *done) return;
  __shared__ double red[9][256];
  const int c = blockIdx.x, tid = threadIdx.x;
  part += (size_t)blockIdx.y * block;
  double a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // T, top, bottom, left, right, tl, tr, bl, br
  for (int r = blockIdx.y + gridDim.y * tid; r < nrows_total; r += 256 * gridDim.y) {
    const float* src = rowsum + ((size_t)r * Cd + c) * 3;
    const float s = src[0], v0 = src[1], v31 = src[2];
    const int y = r % Ht;
    a[0] += s; a[3] += v0; a[4] += v31;
    if (y == 0) { a[1] += s; a[5] += v0; a[6] += v31; }
    if (y == Ht - 1) { a[2] += s; a[7] += v0; a[8] += v31; }
  }
#pragma unroll
  for (int q = 0; q < 9; ++q) red[q][tid] = a[q];
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o)
#pragma unroll
      for (int q = 0; q < 9; ++q) red[q][tid] += red[q][tid + o];
    __syncthreads();
  }
  if (tid < 9) {
    const int dy = tid / 3 - 1, dx = tid % 3 - 1;
    double r = red[0][0];
    if (dy == -1) r -= red[1][0];
    if (dy == 1) r -= red[2][0];
    if (dx == -1) r -= red[3][0];
    if (dx == 1) r -= red[4][0];
    if (dy == -1 && dx == -1) r += red[5][0];
    if (dy == -1 && dx == 1) r += red[6][0];
    if (dy == 1 && dx == -1) r += red[7][0];
    if (dy == 1 && dx == 1) r += red[8][0];
    part[(1 - dx) + 3 * ((1 - dy) + 3 * (Cx + (size_t)CinTot * c))] = tdesc->t * (float)r;   // w[2 - (dx + 1), 2 - (dy + 1), time, co]
  }
}

}  // namespace convtc

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
void convtc_pack(lrnde_ctx* ctx, const ConvTcGeom& g, const ConvTcPackP& p) {
  convtc::PackK k;
  k.p = p; k.Wd = g.Wd; k.Ht = g.Ht; k.PW = g.PW; k.PH = g.PH; k.IMG = g.IMG; k.NPA = g.NPA;
  convtc::pack_kernel<<<g.B * g.Ht, 32 * std::max(1, std::min(8, p.C / 4)), 0, ctx->stream>>>(k);
  LCT_COUNT(ctx);
}

static void convtc_attrs() {   // outside stream capture: the first call is prepare() -> convtc_wpack
  static bool attr_set = false;
  if (attr_set) return;
  const int s64 = convtc::kStages * convtc::stage_bytes(64) + 256, s16 = convtc::kStages * convtc::stage_bytes(16) + 256;
  LR_CUDA(cudaFuncSetAttribute(convtc::conv_kernel<64, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, s64));
  LR_CUDA(cudaFuncSetAttribute(convtc::conv_kernel<64, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, s64));
  LR_CUDA(cudaFuncSetAttribute(convtc::conv_kernel<64, 2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, s64));
  LR_CUDA(cudaFuncSetAttribute(convtc::conv_kernel<16, 0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, s16));
  LR_CUDA(cudaFuncSetAttribute(convtc::conv_kernel<16, 1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, s16));
  LR_CUDA(cudaFuncSetAttribute(convtc::conv_kernel<16, 2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, s16));
  LR_CUDA(cudaFuncSetAttribute(convtc::conv_kernel<64, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, s64));
  LR_CUDA(cudaFuncSetAttribute(convtc::conv_kernel<64, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, s64));
  LR_CUDA(cudaFuncSetAttribute(convtc::conv_kernel<16, 0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, s16));
  LR_CUDA(cudaFuncSetAttribute(convtc::conv_kernel<16, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, s16));
  attr_set = true;
}

size_t convtc_wimg_bytes(int K, int NOUT) { return (size_t)(K / 8) * convtc::wstage_bytes(NOUT); }

void convtc_wpack(lrnde_ctx* ctx, const float* w, int CinTot, int Cout, int transposed, int keep, int td, uint8_t* img,
                  float* tsum) {
  const int K = transposed ? Cout : CinTot - (td ? 1 : 0);
  const int Nreal = transposed ? keep : Cout;
  const int NOUT = convtc_nout(Nreal);
  const int n = (K / 8) * convtc::wstage_bytes(NOUT) / 4;
  convtc_attrs();
  convtc::wpack_kernel<<<std::min(148 * 4, (n + 255) / 256), 256, 0, ctx->stream>>>(w, CinTot, Cout, transposed, K, Nreal, NOUT, td,
                                                                                 reinterpret_cast<float*>(img), tsum);
  LCT_COUNT(ctx);
}

void convtc_conv(lrnde_ctx* ctx, const ConvTcGeom& g, const ConvTcP& p) {
  const int NOUT = convtc_nout(p.Cout);
  const int smem64 = convtc::kStages * convtc::stage_bytes(64) + 256, smem16 = convtc::kStages * convtc::stage_bytes(16) + 256;
  convtc_attrs();
  convtc::ConvK k;
  k.p = p; k.Wd = g.Wd; k.Ht = g.Ht; k.PW = g.PW; k.IMG = g.IMG; k.ngroups = g.ngroups; k.NP = g.NP; k.NPA = g.NPA;
  static const char* dbg_env = getenv("LRNDE_CT_DBG");
  k.dbg = dbg_env ? atoi(dbg_env) : 0;
  const int grid = std::min(g.ngroups, 148);
  const int mode = !p.stat_part ? 0 : (p.bwd_z ? 2 : 1);
  const int T = convtc::kConvThreads, T1 = convtc::kConvThreads + convtc::kConvProd;
  if (p.srcZ) {   // the A operand is formed inside the kernel (forward convolutions on a stored raw conv output)
    if (mode == 2) { lr_set_error("convtc_conv: srcZ with the pullback epilogue is not instantiated"); throw LrError(LRNDE_EINVAL); }
    if (NOUT == 64) {
      if (mode == 0) convtc::conv_kernel<64, 0, 1><<<grid, T1, smem64, ctx->stream>>>(k);
      else convtc::conv_kernel<64, 1, 1><<<grid, T1, smem64, ctx->stream>>>(k);
    } else {
      if (mode == 0) convtc::conv_kernel<16, 0, 1><<<grid, T1, smem16, ctx->stream>>>(k);
      else convtc::conv_kernel<16, 1, 1><<<grid, T1, smem16, ctx->stream>>>(k);
    }
  } else if (NOUT == 64) {
    if (mode == 0) convtc::conv_kernel<64, 0, 0><<<grid, T, smem64, ctx->stream>>>(k);
    else if (mode == 1) convtc::conv_kernel<64, 1, 0><<<grid, T, smem64, ctx->stream>>>(k);
    else convtc::conv_kernel<64, 2, 0><<<grid, T, smem64, ctx->stream>>>(k);
  } else {
    if (mode == 0) convtc::conv_kernel<16, 0, 0><<<grid, T, smem16, ctx->stream>>>(k);
    else if (mode == 1) convtc::conv_kernel<16, 1, 0><<<grid, T, smem16, ctx->stream>>>(k);
    else convtc::conv_kernel<16, 2, 0><<<grid, T, smem16, ctx->stream>>>(k);
  }
  LCT_COUNT(ctx);
}

// ---- weight gradient
bool convtc_wgrad_ok(const ConvTcGeom& g) { return g.Wd == 32; }
static int wg_rows_per_cta(const ConvTcGeom& g) { return std::max(1, (g.B * g.Ht + 147) / 148); }
int convtc_wgrad_splits(const ConvTcGeom& g) {
  const int rows = g.B * g.Ht, rpc = wg_rows_per_cta(g);
  return (rows + rpc - 1) / rpc;
}
static CUtensorMap wg_map(const float* base, const ConvTcGeom& g, int C, int boxC) {
  CUtensorMap m;
  const cuuint64_t dims[4] = {(cuuint64_t)g.Wd, (cuuint64_t)g.Ht, (cuuint64_t)C, (cuuint64_t)g.B};
  const cuuint64_t strides[3] = {(cuuint64_t)g.Wd * 4, (cuuint64_t)g.Wd * g.Ht * 4, (cuuint64_t)g.Wd * g.Ht * C * 4};
  const cuuint32_t box[4] = {32, 1, (cuuint32_t)boxC, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  // the driver entry point is looked up at run time: libLRNDE.so must load (and export its symbols) on a box without libcuda.so.1
  typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeTiled encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    LR_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) {
      lr_set_error("cuTensorMapEncodeTiled is not available from this driver");
      throw LrError(LRNDE_ECUDA);
    }
    encode = reinterpret_cast<EncodeTiled>(fn);
  }
  const CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    lr_set_error("cuTensorMapEncodeTiled failed (%d) for a [%d,%d,%d,%d] tensor", (int)r, g.Wd, g.Ht, C, g.B);
    throw LrError(LRNDE_ECUDA);
  }
  return m;
}
void convtc_wgrad(lrnde_ctx* ctx, const ConvTcGeom& g, const ConvTcWgP& p) {
  static bool attr_set = false;
  // P ring + Q ring of 4 + 2 mirrored slots (hi and lo) + raw P staging ring + alignment
  const int smem64 = convtc::wg_xs(64) * convtc::kWgXSlot + 12 * 64 * 128 + convtc::kWgRS * convtc::kWgPTile + 1024;
  const int smem16 = convtc::wg_xs(16) * convtc::kWgXSlot + 12 * 16 * 128 + convtc::kWgRS * convtc::kWgPTile + 1024;
  if (!attr_set) {
    LR_CUDA(cudaFuncSetAttribute(convtc::wgrad_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem64));
    LR_CUDA(cudaFuncSetAttribute(convtc::wgrad_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem16));
    attr_set = true;
  }
  const bool swapped = p.Cx < p.Cd;                 // the operand with fewer channels is Q
  const int Pc = swapped ? p.Cd : p.Cx, Qc = swapped ? p.Cx : p.Cd;
  const int NB = convtc_nout(Qc);
  const CUtensorMap mP = wg_map(swapped ? p.D : p.X, g, Pc, 64), mQ = wg_map(swapped ? p.X : p.D, g, Qc, NB);
  convtc::WgK k;
  k.Ht = g.Ht; k.nrows_total = g.B * g.Ht; k.rows_per_cta = wg_rows_per_cta(g);
  k.Pc = Pc; k.Qc = Qc; k.swapped = swapped ? 1 : 0; k.CinTot = p.CinTot; k.Cd = p.Cd;
  k.part = p.part; k.block = p.block; k.done = p.done;
  k.p_ab = swapped ? nullptr : p.x_ab; k.p_act = swapped ? ACT_IDENTITY : p.x_act;
  k.q_ab = swapped ? p.x_ab : nullptr; k.q_act = swapped ? p.x_act : ACT_IDENTITY;
  const char* dbg = getenv("LRNDE_WG_DBG");
  k.dbg = dbg ? atoi(dbg) : 0;
  const int grid = convtc_wgrad_splits(g);
  if (k.dbg & 8) return;
  if (NB == 64) convtc::wgrad_kernel<64><<<grid, convtc::kThreads, smem64, ctx->stream>>>(mP, mQ, k);
  else convtc::wgrad_kernel<16><<<grid, convtc::kThreads, smem16, ctx->stream>>>(mP, mQ, k);
  LCT_COUNT(ctx);
  if (p.tdesc && !(k.dbg & 4)) {
    convtc::time_wgrad_kernel<<<dim3(p.Cd, std::min(8, grid)), 256, 0, ctx->stream>>>(p.Drowsum, p.Cd, g.Ht, k.nrows_total, p.tdesc, p.Cx,
                                                                                   p.CinTot, p.part, p.block, p.done);
    LCT_COUNT(ctx);
  }
}

