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
// grid (B * (Ht + 2), 1), block 256 = 64 haloed columns x 4 channel-chunk lanes
// ---------------------------------------------------------------------------------------------------------
struct PackK {
  ConvTcPackP p;
  int Wd, Ht, PW, PH, IMG;
  long NPA;
};

__global__ void __launch_bounds__(256) pack_kernel(PackK k) {
  const ConvTcPackP& p = k.p;
  if (p.done && *p.done) return;
  __shared__ LinComb xd;
  if (p.xdesc) {
    if (threadIdx.x == 0) xd = *p.xdesc;
    __syncthreads();
  }
  const int xh = threadIdx.x & 63, cl = threadIdx.x >> 6;
  if (xh >= k.PW) return;
  const int b = blockIdx.x / k.PH, yh = blockIdx.x % k.PH;
  const bool inside = xh >= 1 && xh <= k.Wd && yh >= 1 && yh <= k.Ht;
  const int x = xh - 1, y = yh - 1;
  const size_t pos = (size_t)kCtGuard + (size_t)b * k.IMG + (size_t)yh * k.PW + xh;
  float* side = p.side_to_desc_dst ? xd.dst : p.side;
  const int C4 = p.C >> 2;
  for (int c4 = cl; c4 < C4; c4 += 4) {
    float hi[4] = {0.0f, 0.0f, 0.0f, 0.0f}, lo[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (inside) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = c4 * 4 + j;
        const size_t idx = x + (size_t)k.Wd * (y + (size_t)k.Ht * (c + (size_t)p.C * b));
        float v;
        if (p.xdesc) {
          v = lr_lincomb_at_tc(xd, idx);
          if (side) side[idx] = v;
        } else v = __ldcg(p.X + idx);
        if (p.bwd_g) {
          const float g = __ldcg(p.bwd_g + idx);
          if (p.bwd_coef) {
            const float a = p.in_ab[c], bb = p.in_ab[p.C + c];
            const float gh = g * lr_dact(p.bwd_act, fmaf(a, v, bb));
            const float xhat = (v - p.bwd_stat[c]) * p.bwd_stat[p.C + c];
            v = a * (gh - p.bwd_coef[c] - xhat * p.bwd_coef[p.C + c]);
          } else v = g * lr_dact(p.bwd_act, v);
        } else {
          if (p.in_ab) v = fmaf(p.in_ab[c], v, p.in_ab[p.C + c]);
          if (p.in_act != ACT_IDENTITY) v = lr_act(p.in_act, v);
        }
        hi[j] = tf32_rna(v);
        lo[j] = tf32_rna(v - hi[j]);
        if (p.Phi) { p.Phi[idx] = hi[j]; p.Plo[idx] = lo[j]; }
      }
    }
    if (p.Fhi) {
      const size_t o = ((size_t)c4 * (size_t)k.NPA + pos) * 4;
      *reinterpret_cast<float4*>(p.Fhi + o) = make_float4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<float4*>(p.Flo + o) = make_float4(lo[0], lo[1], lo[2], lo[3]);
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
    const int n = r % NOUT; r /= NOUT;
    const int k4 = r & 1; r >>= 1;
    const int tap = r % 9; r /= 9;
    const int part = r & 1; r >>= 1;
    const int s = r;
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

template <int NOUT>
__global__ void __launch_bounds__(kThreads, 1) conv_kernel(ConvK k) {
  const ConvTcP& p = k.p;
  if (p.done && *p.done) return;
  constexpr int STB = stage_bytes(NOUT), WSB = wstage_bytes(NOUT);
  constexpr int TCOLS = 2 * kTiles * NOUT;   // 512 (NOUT = 64) or 128 (NOUT = 16)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)smem_raw + 127) & ~(uintptr_t)127);
  __shared__ uint64_t full_bar[kStages], empty_bar[kStages], acc_full[2], tmem_free[2];
  __shared__ uint32_t tmem_slot;

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int nst = p.K >> 3;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1u); mbar_init(&empty_bar[s], 1u); }
    for (int s = 0; s < 2; ++s) { mbar_init(&acc_full[s], 1u); mbar_init(&tmem_free[s], 4u); }
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

  if (warp == 0) {
    // ---------------- bulk copies: per stage of 8 input channels, [hi k4=0 | hi k4=1 | lo k4=0 | lo k4=1] runs + weights
    uint32_t ks = 0;
    for (int g = blockIdx.x; g < k.ngroups; g += gridDim.x) {
      for (int s = 0; s < nst; ++s, ++ks) {
        const uint32_t slot = ks % kStages, ph = (ks / kStages) & 1u;
        mbar_wait(&empty_bar[slot], ph ^ 1u);
        if (elect_one_sync()) {
          uint8_t* dst = sm + (size_t)slot * STB;
          mbar_arrive_expect_tx(&full_bar[slot], (uint32_t)STB);
#pragma unroll
          for (int part = 0; part < 2; ++part)
#pragma unroll
            for (int k4 = 0; k4 < 2; ++k4) {
              const float* src = (part ? p.Flo : p.Fhi) + ((size_t)(2 * s + k4) * (size_t)k.NPA + (size_t)g * kCtGroup) * 4;
              bulk_g2s(dst + (part * 2 + k4) * kRunB, src, (uint32_t)kRunB, &full_bar[slot]);
            }
          bulk_g2s(dst + 4 * kRunB, p.Wimg + (size_t)s * WSB, (uint32_t)WSB, &full_bar[slot]);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ---------------- tcgen05.mma issue: D[128 positions x NOUT] += A(patch shifted by the tap) x B(weights of the tap)
    const uint32_t idesc = make_idesc(128, NOUT);
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
          const uint32_t base = smem_u32(sm + (size_t)slot * STB);
          const uint32_t wb = base + 4 * kRunB;
#pragma unroll 1
          for (int tap = 0; tap < 9; ++tap) {
            const int sh = (tap / 3 - 1) * k.PW + (tap % 3 - 1);
            const uint32_t bh = dlo(wb + tap * (2 * NOUT * 16), NOUT * 16);
            const uint32_t bl = dlo(wb + 9 * 2 * NOUT * 16 + tap * (2 * NOUT * 16), NOUT * 16);
            const uint32_t acc = (s > 0 || tap > 0) ? 1u : 0u;
#pragma unroll
            for (int i = 0; i < kTiles; ++i) {
              const uint32_t aoff = (uint32_t)((kCtGuard + i * 128 + sh) * 16);
              const uint32_t ah = dlo(base + aoff, kRunB), al = dlo(base + 2 * kRunB + aoff, kRunB);
              const uint32_t d = tmem_base + (uint32_t)(db * kTiles * NOUT + i * NOUT);
              mma<0>(d, al, bh, kHiNone, idesc, acc);
              mma<1>(d, ah, bl, kHiNone, idesc, 1u);
              mma<2>(d, ah, bh, kHiNone, idesc, 1u);
            }
          }
          mma_commit(&empty_bar[slot]);
          if (s == nst - 1) mma_commit(&acc_full[db]);
        }
        __syncwarp();
      }
    }
  } else {
    // ---------------- epilogue: lane = position (coalesced along x), registers = 16 output channels at a time
    const int q = warp & 3;
    const uint32_t tlane = tmem_base + ((uint32_t)(q * 32) << 16);
    float* out = p.ydesc ? p.ydesc->dst : p.Y;
    const float tval = p.tdesc ? p.tdesc->t : 0.0f;
    const size_t HW = (size_t)k.Wd * k.Ht;
    int it = 0;
    for (int g = blockIdx.x; g < k.ngroups; g += gridDim.x, ++it) {
      const int db = it & 1;
      size_t obase[kTiles];
      int cls[kTiles];
      bool valid[kTiles];
#pragma unroll
      for (int i = 0; i < kTiles; ++i) {
        const long pp = (long)g * kCtGroup + i * 128 + q * 32 + lane;
        const int b = (int)(pp / k.IMG), r = (int)(pp % k.IMG);
        const int yh = r / k.PW, xh = r % k.PW;
        valid[i] = pp < k.NP && xh >= 1 && xh <= k.Wd && yh >= 1 && yh <= k.Ht;
        const int x = xh - 1, y = yh - 1;
        obase[i] = (size_t)x + (size_t)k.Wd * ((size_t)y + (size_t)k.Ht * ((size_t)p.Cout * b));
        const int ym = (y > 0 ? 1 : 0) | 2 | (y < k.Ht - 1 ? 4 : 0), xm = (x > 0 ? 1 : 0) | 2 | (x < k.Wd - 1 ? 4 : 0);
        cls[i] = (ym * 8 + xm) * NOUT;
      }
      mbar_wait(&acc_full[db], (uint32_t)((it >> 1) & 1));
      tc_fence_after();
#pragma unroll 1
      for (int cc = 0; cc < NOUT / 16; ++cc) {
        float st[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) st[j] = 0.0f;
#pragma unroll
        for (int i = 0; i < kTiles; ++i) {
          float v[16];
          tmem_ld16(tlane + (uint32_t)(db * kTiles * NOUT + i * NOUT + cc * 16), v);
          if (valid[i]) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int co = cc * 16 + j;
              if (co < p.Cout) {
                float r = v[j];
                if (p.tsum) r = fmaf(tval, __ldg(p.tsum + cls[i] + co), r);
                st[j] += r;
                st[16 + j] = fmaf(r, r, st[16 + j]);
                out[obase[i] + (size_t)co * HW] = r * p.out_scale;
              }
            }
          }
        }
        if (p.stat_part) {
          const float tot = warp_transpose_sum(st, lane);
          const int co = cc * 16 + (lane & 15);
          if (co < p.Cout)
            reinterpret_cast<float*>(p.stat_part + ((size_t)g * 4 + q) * p.Cout + co)[lane >> 4] = tot;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_free[db]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)TCOLS) : "memory");
  }
}

}  // namespace convtc

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
void convtc_pack(lrnde_ctx* ctx, const ConvTcGeom& g, const ConvTcPackP& p) {
  convtc::PackK k;
  k.p = p; k.Wd = g.Wd; k.Ht = g.Ht; k.PW = g.PW; k.PH = g.PH; k.IMG = g.IMG; k.NPA = g.NPA;
  convtc::pack_kernel<<<g.B * g.PH, 256, 0, ctx->stream>>>(k);
  LCT_COUNT(ctx);
}

static void convtc_attrs() {   // outside stream capture: the first call is prepare() -> convtc_wpack
  static bool attr_set = false;
  if (attr_set) return;
  LR_CUDA(cudaFuncSetAttribute(convtc::conv_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               convtc::kStages * convtc::stage_bytes(64) + 256));
  LR_CUDA(cudaFuncSetAttribute(convtc::conv_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               convtc::kStages * convtc::stage_bytes(16) + 256));
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
  const int grid = std::min(g.ngroups, 148);
  if (NOUT == 64) convtc::conv_kernel<64><<<grid, convtc::kThreads, smem64, ctx->stream>>>(k);
  else convtc::conv_kernel<16><<<grid, convtc::kThreads, smem16, ctx->stream>>>(k);
  LCT_COUNT(ctx);
}
