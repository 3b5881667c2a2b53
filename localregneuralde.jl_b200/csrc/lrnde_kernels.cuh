// FP32 SIMT kernels of the hot path (precision mode LRNDE_PREC_FP32_SIMT) plus every
// precision-independent kernel: stage linear combinations, Hairer error norm, the on-device
// step-size controller, initial-dt heuristic, regulariser read-out and its reverse pass.
//
// Reference arithmetic restated (file:line in /root/reference):
//   stage inputs  uprev + dt*(a_j1 k1 + ...)            src/perform_step.jl:11-18
//   utilde / residual / RMS                              src/perform_step.jl:21-27,34-38,208-212
//   stiffness estimate                                   src/perform_step.jl:40-47
//   TDChain time row appended before every Dense         src/layers/common.jl:19-33
// and the un-vendored solver loop of SURVEY App. A (lrnde_controller.h).
#pragma once
#include <cuda_runtime.h>

#include "lrnde_device.h"
#include "lrnde_act.cuh"

__device__ __forceinline__ float lr_lincomb_at(const LinComb& d, size_t i) {
  float inner = 0.0f;
  for (int k = 0; k < d.n; ++k) inner = fmaf(d.coef[k], d.src[k][i], inner);
  float b = d.base ? d.base[i] : 0.0f;
  return d.n ? fmaf(d.scale, inner, b) : b;
}

// ------------------------------------------------------------------------------------------
// Dense layer as one GEMM:  Y[M x N] = epi( A[M x Kaug] * Xaug[Kaug x N] )
//   A    : column-major slice of the flat parameter vector (Lux weight [out x in(+1)] followed
//          directly by the bias = one [out x (in + td + 1)] block), or a pre-transposed weight
//   Xaug : rows 0..K-1 from X (sample-contiguous, ld = ldx) or from a LinComb descriptor (the
//          stage combination is formed while the tile is loaded); row K = t (TDChain time row)
//          when td; next row = 1 (bias) when bias
//   epi  : act(acc), optionally acc * act'(dpre) (reverse pass), times out_scale
// ------------------------------------------------------------------------------------------
struct DenseP {
  const float* A; long lda; int M; int K; int td; int bias;
  const float* X; const LinComb* xdesc; long ldx; int N; int in_act;
  float* side; const LinComb* side_desc;  // optional: also store the (pre-in_act) X tile, fp32
  float* Y; const LinComb* ydesc; long y_off; long ldy;
  float* pre; long ldpre;
  int act;
  const float* dpre; long lddpre; int dact;
  float out_scale;
  const LinComb* tdesc;
  const int* done;
};

#define DN_BM 128
#define DN_BN 64
#define DN_BK 16

__global__ void __launch_bounds__(256) dense_nn_kernel(DenseP p) {
  if (p.done && *p.done) return;
  __shared__ __align__(16) float As[2][DN_BK][DN_BM];
  __shared__ __align__(16) float Bs[2][DN_BK][DN_BN + 4];
  __shared__ LinComb sdesc;
  __shared__ float s_t;

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * DN_BM, n0 = blockIdx.y * DN_BN;
  const int Kaug = p.K + p.td + p.bias;

  if (tid == 0) {
    if (p.xdesc) sdesc = *p.xdesc;
    else { sdesc.base = p.X; sdesc.n = 0; sdesc.scale = 0.f; }
    s_t = p.tdesc ? p.tdesc->t : 0.0f;
  }
  __syncthreads();
  const float tval = s_t;
  float* side = p.side_desc ? p.side_desc->dst : p.side;
  const bool a_vec = ((((uintptr_t)p.A) & 15) == 0) && (p.lda % 4 == 0);
  bool x_vec = (p.ldx % 4 == 0) && ((((uintptr_t)sdesc.base) & 15) == 0) && ((((uintptr_t)side) & 15) == 0);
  for (int k = 0; k < sdesc.n; ++k) x_vec = x_vec && ((((uintptr_t)sdesc.src[k]) & 15) == 0);

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  float4 ra[2];
  float rx[4];

  auto load_tiles = [&](int k0) {
    // A tile: 128 x 16
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int idx = tid + i * 256;
      int m4 = idx & 31, k = idx >> 5;
      int gm = m0 + m4 * 4, gk = k0 + k;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gk < Kaug) {
        const float* ap = p.A + (size_t)gk * p.lda + gm;
        if (a_vec && gm + 3 < p.M) {
          v = *reinterpret_cast<const float4*>(ap);
        } else {
          if (gm + 0 < p.M) v.x = ap[0];
          if (gm + 1 < p.M) v.y = ap[1];
          if (gm + 2 < p.M) v.z = ap[2];
          if (gm + 3 < p.M) v.w = ap[3];
        }
      }
      ra[i] = v;
    }
    // X tile: 16 (k) x 64 (n)
    {
      int n = tid >> 2, kq = (tid & 3) * 4;
      int gn = n0 + n, gk = k0 + kq;
      rx[0] = rx[1] = rx[2] = rx[3] = 0.0f;
      if (gn < p.N) {
        size_t off = (size_t)gn * p.ldx + gk;
        if (x_vec && gk + 3 < p.K) {
          float4 inner = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int s = 0; s < sdesc.n; ++s) {
            float4 v = *reinterpret_cast<const float4*>(sdesc.src[s] + off);
            float c = sdesc.coef[s];
            inner.x = fmaf(c, v.x, inner.x); inner.y = fmaf(c, v.y, inner.y);
            inner.z = fmaf(c, v.z, inner.z); inner.w = fmaf(c, v.w, inner.w);
          }
          float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
          if (sdesc.base) b = *reinterpret_cast<const float4*>(sdesc.base + off);
          if (sdesc.n) {
            rx[0] = fmaf(sdesc.scale, inner.x, b.x); rx[1] = fmaf(sdesc.scale, inner.y, b.y);
            rx[2] = fmaf(sdesc.scale, inner.z, b.z); rx[3] = fmaf(sdesc.scale, inner.w, b.w);
          } else { rx[0] = b.x; rx[1] = b.y; rx[2] = b.z; rx[3] = b.w; }
          if (side && blockIdx.x == 0)
            *reinterpret_cast<float4*>(side + off) = make_float4(rx[0], rx[1], rx[2], rx[3]);
          if (p.in_act) {
#pragma unroll
            for (int e = 0; e < 4; ++e) rx[e] = lr_act(p.in_act, rx[e]);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            int k = gk + e;
            float v = 0.0f;
            if (k < p.K) {
              v = lr_lincomb_at(sdesc, off + e);
              if (side && blockIdx.x == 0) side[off + e] = v;
              if (p.in_act) v = lr_act(p.in_act, v);
            } else if (p.td && k == p.K) v = tval;
            else if (p.bias && k == p.K + p.td) v = 1.0f;
            rx[e] = v;
          }
        }
      }
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int idx = tid + i * 256;
      int m4 = idx & 31, k = idx >> 5;
      *reinterpret_cast<float4*>(&As[buf][k][m4 * 4]) = ra[i];
    }
    int n = tid >> 2, kq = (tid & 3) * 4;
#pragma unroll
    for (int e = 0; e < 4; ++e) Bs[buf][kq + e][n] = rx[e];
  };

  const int nk = (Kaug + DN_BK - 1) / DN_BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles((kt + 1) * DN_BK);
#pragma unroll
    for (int kk = 0; kk < DN_BK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][tx * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + tx * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[buf][kk][ty * 4]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) store_tiles(buf ^ 1);
    __syncthreads();
  }

  float* Y = p.ydesc ? (p.ydesc->dst + p.y_off) : p.Y;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int gn = n0 + ty * 4 + j;
    if (gn >= p.N) continue;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int gm = m0 + (i < 4 ? tx * 4 + i : 64 + tx * 4 + (i - 4));
      if (gm >= p.M) continue;
      float v = acc[i][j];
      if (p.pre) p.pre[(size_t)gn * p.ldpre + gm] = v;
      if (p.dact >= 0) v = v * lr_dact(p.dact, p.dpre[(size_t)gn * p.lddpre + gm]);
      else v = lr_act(p.act, v);
      Y[(size_t)gn * p.ldy + gm] = v * p.out_scale;
    }
  }
}

// ------------------------------------------------------------------------------------------
// Weight gradient of one Dense:  dWaug[M x Naug] = sum_b delta[M x b] * Xaug[Naug x b]^T
// (Xaug rows: layer input, then t when td, then 1 when bias => the [weight | bias] block of
// the flat gradient).  Split over the batch; partials are reduced in a fixed order.
// ------------------------------------------------------------------------------------------
struct WgradP {
  const float* Dl; long ldd; int M;
  const float* X; long ldx; int Nin; int td; int bias; int in_act;
  int B; int chunk;
  float* part;
  const LinComb* tdesc;
  const int* done;
};

__global__ void __launch_bounds__(256) wgrad_nt_kernel(WgradP p) {
  if (p.done && *p.done) return;
  __shared__ __align__(16) float As[2][DN_BK][DN_BM];
  __shared__ __align__(16) float Bs[2][DN_BK][DN_BN];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.x * DN_BM, n0 = blockIdx.y * DN_BN;
  const int Naug = p.Nin + p.td + p.bias;
  const int b_begin = blockIdx.z * p.chunk;
  const int b_end = min(p.B, b_begin + p.chunk);
  const float tval = p.tdesc ? p.tdesc->t : 0.0f;
  const bool a_vec = ((((uintptr_t)p.Dl) & 15) == 0) && (p.ldd % 4 == 0);
  const bool x_vec = ((((uintptr_t)p.X) & 15) == 0) && (p.ldx % 4 == 0);

  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  float4 ra[2];
  float4 rb;

  auto load_tiles = [&](int b0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int idx = tid + i * 256;
      int m4 = idx & 31, k = idx >> 5;
      int gm = m0 + m4 * 4, gb = b0 + k;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (gb < b_end) {
        const float* ap = p.Dl + (size_t)gb * p.ldd + gm;
        if (a_vec && gm + 3 < p.M) v = *reinterpret_cast<const float4*>(ap);
        else {
          if (gm + 0 < p.M) v.x = ap[0];
          if (gm + 1 < p.M) v.y = ap[1];
          if (gm + 2 < p.M) v.z = ap[2];
          if (gm + 3 < p.M) v.w = ap[3];
        }
      }
      ra[i] = v;
    }
    {
      int k = tid >> 4, n4 = (tid & 15) * 4;
      int gn = n0 + n4, gb = b0 + k;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
      if (gb < b_end) {
        const float* xp = p.X + (size_t)gb * p.ldx + gn;
        if (x_vec && gn + 3 < p.Nin) {
          float4 q = *reinterpret_cast<const float4*>(xp);
          v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
          if (p.in_act) {
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = lr_act(p.in_act, v[e]);
          }
        } else {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            int n = gn + e;
            if (n < p.Nin) { v[e] = xp[e]; if (p.in_act) v[e] = lr_act(p.in_act, v[e]); }
            else if (p.td && n == p.Nin) v[e] = tval;
            else if (p.bias && n == p.Nin + p.td) v[e] = 1.0f;
          }
        }
      }
      rb = make_float4(v[0], v[1], v[2], v[3]);
    }
  };
  auto store_tiles = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      int idx = tid + i * 256;
      int m4 = idx & 31, k = idx >> 5;
      *reinterpret_cast<float4*>(&As[buf][k][m4 * 4]) = ra[i];
    }
    int k = tid >> 4, n4 = (tid & 15) * 4;
    *reinterpret_cast<float4*>(&Bs[buf][k][n4]) = rb;
  };

  const int nk = (b_end - b_begin + DN_BK - 1) / DN_BK;
  if (nk > 0) {
    load_tiles(b_begin);
    store_tiles(0);
  }
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tiles(b_begin + (kt + 1) * DN_BK);
#pragma unroll
    for (int kk = 0; kk < DN_BK; ++kk) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][tx * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + tx * 4]);
      float4 b = *reinterpret_cast<const float4*>(&Bs[buf][kk][ty * 4]);
      float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) store_tiles(buf ^ 1);
    __syncthreads();
  }
  float* out = p.part + (size_t)blockIdx.z * p.M * Naug;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int gn = n0 + ty * 4 + j;
    if (gn >= Naug) continue;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int gm = m0 + (i < 4 ? tx * 4 + i : 64 + tx * 4 + (i - 4));
      if (gm >= p.M) continue;
      out[(size_t)gn * p.M + gm] = acc[i][j];
    }
  }
}

// dst[i] = beta * dst[i] + scale * sum_z part[z][i]   (z ascending: deterministic)
__global__ void wgrad_reduce_kernel(const float* part, int S, size_t n, float* dst,
                                    const LinComb* dstdesc, size_t dst_off, float scale,
                                    float beta, const int* done) {
  if (done && *done) return;
  float* out = dstdesc ? (dstdesc->dst + dst_off) : dst;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x) {
    float s = 0.0f;
    for (int z = 0; z < S; ++z) s += part[(size_t)z * n + i];
    float r = scale * s;
    if (beta != 0.0f) r = fmaf(beta, out[i], r);
    out[i] = r;
  }
}

// ------------------------------------------------------------------------------------------
// Elementwise
// ------------------------------------------------------------------------------------------
// out[off + i] = lincomb(off + i), i < n   (dst == nullptr: out = descriptor's dst)
__global__ void lincomb_kernel(const LinComb* dp, float* dst, size_t n, const int* done, size_t off = 0) {
  if (done && *done) return;
  __shared__ LinComb d;
  if (threadIdx.x == 0) d = *dp;
  __syncthreads();
  float* out = dst ? dst : d.dst;
  // 16-byte path (every operand aligned, which library buffers are): one float4 per source in flight
  bool vec = ((n & 3) == 0) && ((off & 3) == 0) && ((((uintptr_t)out) & 15) == 0) && ((((uintptr_t)d.base) & 15) == 0);
  for (int k = 0; k < d.n; ++k) vec = vec && ((((uintptr_t)d.src[k]) & 15) == 0);
  if (vec) {
    const size_t n4 = n >> 2, o4 = off >> 2;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
      float4 v[LR_MAXSRC];
#pragma unroll
      for (int k = 0; k < LR_MAXSRC; ++k)
        if (k < d.n) v[k] = __ldcg(reinterpret_cast<const float4*>(d.src[k]) + o4 + i);
      float4 b = d.base ? __ldcg(reinterpret_cast<const float4*>(d.base) + o4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 in = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < LR_MAXSRC; ++k)
        if (k < d.n) {
          in.x = fmaf(d.coef[k], v[k].x, in.x); in.y = fmaf(d.coef[k], v[k].y, in.y);
          in.z = fmaf(d.coef[k], v[k].z, in.z); in.w = fmaf(d.coef[k], v[k].w, in.w);
        }
      if (d.n) { b.x = fmaf(d.scale, in.x, b.x); b.y = fmaf(d.scale, in.y, b.y); b.z = fmaf(d.scale, in.z, b.z); b.w = fmaf(d.scale, in.w, b.w); }
      reinterpret_cast<float4*>(out)[o4 + i] = b;
    }
    return;
  }
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    out[off + i] = lr_lincomb_at(d, off + i);
}

// out[i] = a[i] * act'(pre[i])
__global__ void dact_mul_kernel(const LinComb* adesc, const float* pre, int act, float* out,
                                size_t n, const int* done) {
  if (done && *done) return;
  __shared__ LinComb d;
  if (threadIdx.x == 0) d = *adesc;
  __syncthreads();
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    out[i] = lr_lincomb_at(d, i) * lr_dact(act, pre[i]);
}

// y[i] += alpha * x[i]  (y = current state of the solve: lambda jump of the adjoint)
__global__ void jump_kernel(SolveDev* S, const float* d, size_t n) {
  if (S->failed) return;
  float* z = lr_slot_u(S, S->slot);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    z[i] += d[i];
}

__global__ void axpy_kernel(float* y, const float* x, float alpha, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    y[i] = fmaf(alpha, x[i], y[i]);
}

__device__ __forceinline__ double lr_block_sum(double v) {
  __shared__ double sh[32];
  __syncthreads();
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) sh[w] = v;
  __syncthreads();
  double r = 0.0;
  if (w == 0) {
    int nw = (blockDim.x + 31) >> 5;
    r = (l < nw) ? sh[l] : 0.0;
    for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
  }
  return r;  // valid in thread 0
}

// Hairer error norm partial sums of one step attempt (perform_step.jl:21-27,34-38,210-212):
// utilde = dt * sum btilde_i k_i ; r = utilde / (abstol + max(|uprev|,|u|) * reltol).
// For the adjoint in a data-parallel group the mu block [lam_len, len) holds per-rank partial
// batch sums; it is skipped here and handled by err_norm_mu_kernel over the global values.
__global__ void __launch_bounds__(256) err_norm_kernel(SolveDev* S) {
  if (S->done) return;
  __shared__ LinComb d;
  if (threadIdx.x == 0) d = S->err;
  __syncthreads();
  const float abstol = S->abstol, reltol = S->reltol;
  const size_t n = (S->reduce_mu && S->nranks > 1) ? S->lam_len : S->len;
  const float* uprev = d.base;
  const float* unew = d.dst;
  double acc = 0.0;
  bool vec = ((n & 3) == 0) && ((((uintptr_t)uprev) & 15) == 0) && ((((uintptr_t)unew) & 15) == 0);
#pragma unroll
  for (int k = 0; k < 7; ++k) vec = vec && ((((uintptr_t)d.src[k]) & 15) == 0);
  if (vec) {   // nine float4 loads in flight per thread (the kernel is a pure HBM stream: 9 arrays in)
    const size_t n4 = n >> 2;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
      float4 v[7];
#pragma unroll
      for (int k = 0; k < 7; ++k) v[k] = __ldcg(reinterpret_cast<const float4*>(d.src[k]) + i);
      const float4 up = __ldcg(reinterpret_cast<const float4*>(uprev) + i);
      const float4 un = __ldcg(reinterpret_cast<const float4*>(unew) + i);
      float4 in = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        in.x = fmaf(d.coef[k], v[k].x, in.x); in.y = fmaf(d.coef[k], v[k].y, in.y);
        in.z = fmaf(d.coef[k], v[k].z, in.z); in.w = fmaf(d.coef[k], v[k].w, in.w);
      }
      const float r0 = (d.scale * in.x) / (abstol + fmaxf(fabsf(up.x), fabsf(un.x)) * reltol);
      const float r1 = (d.scale * in.y) / (abstol + fmaxf(fabsf(up.y), fabsf(un.y)) * reltol);
      const float r2 = (d.scale * in.z) / (abstol + fmaxf(fabsf(up.z), fabsf(un.z)) * reltol);
      const float r3 = (d.scale * in.w) / (abstol + fmaxf(fabsf(up.w), fabsf(un.w)) * reltol);
      acc += (double)(r0 * r0) + (double)(r1 * r1) + (double)(r2 * r2) + (double)(r3 * r3);
    }
  } else {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
         i += (size_t)gridDim.x * blockDim.x) {
      float inner = 0.0f;
#pragma unroll
      for (int k = 0; k < 7; ++k) inner = fmaf(d.coef[k], d.src[k][i], inner);
      float ut = d.scale * inner;
      float r = ut / (abstol + fmaxf(fabsf(uprev[i]), fabsf(unew[i])) * reltol);
      acc += (double)(r * r);
    }
  }
  double s = lr_block_sum(acc);
  if (threadIdx.x == 0) S->partials[blockIdx.x] = s;
}

// initdt norms, pass 1: sum (u0/sk)^2, sum (f0/sk)^2, count of non-finite f0.  16-byte loads, two groups in flight
// per thread (the scalar grid-stride loop was latency-bound: 55 us for 51 MB); scalar path for unaligned slots.
__device__ __forceinline__ void lr_initdt1_elem(float u, float f, float abstol, float reltol, double& a0, double& a1,
                                                unsigned int& bad) {
  const float sk = abstol + fabsf(u) * reltol;
  const float q0 = u / sk, q1 = f / sk;
  a0 += (double)(q0 * q0);
  a1 += (double)(q1 * q1);
  if (!isfinite(f)) bad++;
}
__global__ void __launch_bounds__(256) initdt_norm1_kernel(SolveDev* S) {
  if (S->failed) return;
  const float* u0 = lr_slot_u(S, S->slot);
  const float* f0 = lr_slot_k(S, S->slot, 1);
  const float abstol = S->abstol, reltol = S->reltol;
  double a0 = 0.0, a1 = 0.0;
  unsigned int bad = 0;
  const size_t nloc = (S->reduce_mu && S->nranks > 1) ? S->lam_len : S->len;
  const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
  size_t done = 0;
  if (((((uintptr_t)u0) | ((uintptr_t)f0)) & 15) == 0) {
    const size_t n4 = nloc >> 2;
    const float4* u4 = reinterpret_cast<const float4*>(u0);
    const float4* f4 = reinterpret_cast<const float4*>(f0);
    size_t i = tid;
    for (; i + nthr < n4; i += 2 * nthr) {
      const float4 ua = __ldcg(u4 + i), fa = __ldcg(f4 + i), ub = __ldcg(u4 + i + nthr), fb = __ldcg(f4 + i + nthr);
      lr_initdt1_elem(ua.x, fa.x, abstol, reltol, a0, a1, bad); lr_initdt1_elem(ua.y, fa.y, abstol, reltol, a0, a1, bad);
      lr_initdt1_elem(ua.z, fa.z, abstol, reltol, a0, a1, bad); lr_initdt1_elem(ua.w, fa.w, abstol, reltol, a0, a1, bad);
      lr_initdt1_elem(ub.x, fb.x, abstol, reltol, a0, a1, bad); lr_initdt1_elem(ub.y, fb.y, abstol, reltol, a0, a1, bad);
      lr_initdt1_elem(ub.z, fb.z, abstol, reltol, a0, a1, bad); lr_initdt1_elem(ub.w, fb.w, abstol, reltol, a0, a1, bad);
    }
    for (; i < n4; i += nthr) {
      const float4 ua = __ldcg(u4 + i), fa = __ldcg(f4 + i);
      lr_initdt1_elem(ua.x, fa.x, abstol, reltol, a0, a1, bad); lr_initdt1_elem(ua.y, fa.y, abstol, reltol, a0, a1, bad);
      lr_initdt1_elem(ua.z, fa.z, abstol, reltol, a0, a1, bad); lr_initdt1_elem(ua.w, fa.w, abstol, reltol, a0, a1, bad);
    }
    done = n4 << 2;
  }
  for (size_t i = done + tid; i < nloc; i += nthr) lr_initdt1_elem(u0[i], f0[i], abstol, reltol, a0, a1, bad);
  double s0 = lr_block_sum(a0);
  double s1 = lr_block_sum(a1);
  if (threadIdx.x == 0) {
    S->partials[blockIdx.x] = s0;
    S->partials[LR_ERR_BLOCKS + blockIdx.x] = s1;
  }
  if (bad) atomicAdd(&S->counters[0], bad);
}

// pass 2: sum ((f1-f0)/sk)^2 and count of f1 != f0; f1 sits in K(slot, 2)
__device__ __forceinline__ void lr_initdt2_elem(float u, float f0v, float f1v, float abstol, float reltol, double& a2,
                                                unsigned int& neq) {
  const float sk = abstol + fabsf(u) * reltol;
  const float q = (f1v - f0v) / sk;
  a2 += (double)(q * q);
  if (!(f1v == f0v)) neq++;
}
__global__ void __launch_bounds__(256) initdt_norm2_kernel(SolveDev* S) {
  if (S->failed) return;
  const float* u0 = lr_slot_u(S, S->slot);
  const float* f0 = lr_slot_k(S, S->slot, 1);
  const float* f1 = lr_slot_k(S, S->slot, 2);
  const float abstol = S->abstol, reltol = S->reltol;
  double a2 = 0.0;
  unsigned int neq = 0;
  const size_t nloc = (S->reduce_mu && S->nranks > 1) ? S->lam_len : S->len;
  const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
  size_t done = 0;
  if (((((uintptr_t)u0) | ((uintptr_t)f0) | ((uintptr_t)f1)) & 15) == 0) {
    const size_t n4 = nloc >> 2;
    const float4* u4 = reinterpret_cast<const float4*>(u0);
    const float4* p4 = reinterpret_cast<const float4*>(f0);
    const float4* q4 = reinterpret_cast<const float4*>(f1);
    size_t i = tid;
    for (; i + nthr < n4; i += 2 * nthr) {
      const float4 ua = __ldcg(u4 + i), pa = __ldcg(p4 + i), qa = __ldcg(q4 + i);
      const float4 ub = __ldcg(u4 + i + nthr), pb = __ldcg(p4 + i + nthr), qb = __ldcg(q4 + i + nthr);
      lr_initdt2_elem(ua.x, pa.x, qa.x, abstol, reltol, a2, neq); lr_initdt2_elem(ua.y, pa.y, qa.y, abstol, reltol, a2, neq);
      lr_initdt2_elem(ua.z, pa.z, qa.z, abstol, reltol, a2, neq); lr_initdt2_elem(ua.w, pa.w, qa.w, abstol, reltol, a2, neq);
      lr_initdt2_elem(ub.x, pb.x, qb.x, abstol, reltol, a2, neq); lr_initdt2_elem(ub.y, pb.y, qb.y, abstol, reltol, a2, neq);
      lr_initdt2_elem(ub.z, pb.z, qb.z, abstol, reltol, a2, neq); lr_initdt2_elem(ub.w, pb.w, qb.w, abstol, reltol, a2, neq);
    }
    for (; i < n4; i += nthr) {
      const float4 ua = __ldcg(u4 + i), pa = __ldcg(p4 + i), qa = __ldcg(q4 + i);
      lr_initdt2_elem(ua.x, pa.x, qa.x, abstol, reltol, a2, neq); lr_initdt2_elem(ua.y, pa.y, qa.y, abstol, reltol, a2, neq);
      lr_initdt2_elem(ua.z, pa.z, qa.z, abstol, reltol, a2, neq); lr_initdt2_elem(ua.w, pa.w, qa.w, abstol, reltol, a2, neq);
    }
    done = n4 << 2;
  }
  for (size_t i = done + tid; i < nloc; i += nthr) lr_initdt2_elem(u0[i], f0[i], f1[i], abstol, reltol, a2, neq);
  double s2 = lr_block_sum(a2);
  if (threadIdx.x == 0) S->partials[2 * LR_ERR_BLOCKS + blockIdx.x] = s2;
  if (neq) atomicAdd(&S->counters[1], neq);
}

// ------------------------------------------------------------------------------------------
// Data-parallel group: sum up to 4 doubles (+ 2 counters folded in as doubles) over ranks
// through peer-mapped mailboxes; every rank adds in rank order => bit-identical results, so
// the replicated controllers take identical decisions.  Single calling thread.
// ------------------------------------------------------------------------------------------
__device__ inline void lr_group_sum(SolveDev* S, double* v) {
  if (S->nranks <= 1) return;
  const unsigned long long seq = S->seq;
  S->seq = seq + 1;
  const int par = (int)(seq & 1ull);
  const int me = S->rank;
  for (int r = 0; r < S->nranks; ++r) {
    volatile double* dst = S->mbox[r]->val[par][me];
    for (int k = 0; k < 4; ++k) dst[k] = v[k];
  }
  __threadfence_system();
  for (int r = 0; r < S->nranks; ++r) {
    volatile unsigned long long* f = &S->mbox[r]->flag[par][me];
    *f = seq + 1;
  }
  LrMailbox* mine = S->mbox[me];
  for (int r = 0; r < S->nranks; ++r) {
    volatile unsigned long long* f = &mine->flag[par][r];
    long long spins = 0;
    while (*f != seq + 1) {
      __nanosleep(20);
      if (++spins > (1ll << 27)) { lr_peer_timeout(S); break; }   // a peer never arrived (~10 s): fail the solve instead of hanging
    }
    if (S->failed) break;
  }
  __threadfence_system();
  for (int k = 0; k < 4; ++k) {
    double s = 0.0;
    for (int r = 0; r < S->nranks; ++r) s += ((volatile double*)mine->val[par][r])[k];
    v[k] = s;
  }
}

// ------------------------------------------------------------------------------------------
// Adjoint in a data-parallel group: z = [lambda_local ; mu_partial].  mu is a batch sum, and the
// step controller's norms are non-linear in it, so the vectors they read are exchanged:
// every rank publishes its partial vectors in its peer-mapped staging area, then every rank adds
// all ranks' vectors in rank order (identical bits everywhere).  mode 0: f0_mu ; mode 1:
// (f1 - f0)_mu ; mode 2: utilde_mu, mu_prev, mu_new of the attempt.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mu_publish_kernel(SolveDev* S, int mode, const int* done) {
  if (done && *done) return;
  __shared__ LinComb e;
  if (threadIdx.x == 0) e = S->err;
  __syncthreads();
  const size_t P = S->mu_len, off = S->lam_len;
  const int par = (int)(S->mseq & 1ull);
  LrMailbox* mine = S->mbox[S->rank];
  float* v0 = lr_mbox_stage(mine, par, 0);
  float* v1 = lr_mbox_stage(mine, par, 1);
  float* v2 = lr_mbox_stage(mine, par, 2);
  const float* f0 = lr_slot_k(S, S->slot, 1) + off;
  const float* f1 = lr_slot_k(S, S->slot, 2) + off;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < P; i += (size_t)gridDim.x * blockDim.x) {
    if (mode == 0) v0[i] = f0[i];
    else if (mode == 1) v0[i] = f1[i] - f0[i];
    else {
      float inner = 0.0f;
#pragma unroll
      for (int k = 0; k < 7; ++k) inner = fmaf(e.coef[k], e.src[k][off + i], inner);
      v0[i] = e.scale * inner;
      v1[i] = e.base[off + i];
      v2[i] = e.dst[off + i];
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int prev = atomicAdd(S->mucounter, 1u);
    if (prev == gridDim.x - 1) {  // last block: everything of this rank is visible, raise the flags
      *S->mucounter = 0;
      __threadfence_system();
      for (int r = 0; r < S->nranks; ++r) {
        volatile unsigned long long* f = &S->mbox[r]->mflag[par][S->rank];
        *f = S->mseq + 1;
      }
    }
  }
}

__global__ void __launch_bounds__(256) mu_gather_kernel(SolveDev* S, int nvec, const int* done) {
  if (done && *done) return;
  const int par = (int)(S->mseq & 1ull);
  LrMailbox* mine = S->mbox[S->rank];
  if (threadIdx.x == 0) {
    for (int r = 0; r < S->nranks; ++r) {
      volatile unsigned long long* f = &mine->mflag[par][r];
      long long spins = 0;
      while (*f != S->mseq + 1) {
        __nanosleep(40);
        if (++spins > (1ll << 26)) { lr_peer_timeout(S); break; }
      }
      if (S->failed) break;
    }
    __threadfence_system();
  }
  __syncthreads();
  const size_t P = S->mu_len;
  for (int v = 0; v < nvec; ++v) {
    float* out = S->muglob + (size_t)v * P;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < P; i += (size_t)gridDim.x * blockDim.x) {
      float s = 0.0f;
      for (int r = 0; r < S->nranks; ++r) {
        const volatile float* src = lr_mbox_stage(S->mbox[r], par, v);
        s += src[i];
      }
      out[i] = s;
    }
  }
}

// mu part of the norms from the exchanged global vectors -> partials[3 * LR_ERR_BLOCKS + b]
__global__ void __launch_bounds__(256) mu_norm_kernel(SolveDev* S, int mode, const int* done) {
  if (done && *done) return;
  const size_t P = S->mu_len;
  const float abstol = S->abstol, reltol = S->reltol;
  const float* g0 = S->muglob;
  const float* g1 = S->muglob + P;
  const float* g2 = S->muglob + 2 * P;
  double acc = 0.0;
  unsigned int cnt = 0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < P; i += (size_t)gridDim.x * blockDim.x) {
    float r;
    if (mode == 2) r = g0[i] / (abstol + fmaxf(fabsf(g1[i]), fabsf(g2[i])) * reltol);
    else r = g0[i] / abstol;  // mu(t2) = 0  =>  sk = abstol
    acc += (double)(r * r);
    if (mode == 0 && !isfinite(g0[i])) cnt++;
    if (mode == 1 && !(g0[i] == 0.0f)) cnt++;
  }
  double s = lr_block_sum(acc);
  if (threadIdx.x == 0) S->partials[3 * LR_ERR_BLOCKS + blockIdx.x] = s;
  if (cnt) atomicAdd(&S->counters[3], cnt);
  if (blockIdx.x == 0 && threadIdx.x == 0) S->mseq += 1;  // next exchange uses the other parity
}

// fixed-order sum of LR_ERR_BLOCKS partials by one warp
__device__ __forceinline__ double lr_sum_partials(const double* p) {
  // all loads in flight before the (order-preserving) adds: the controller is a single warp on the critical path
  constexpr int NV = (LR_ERR_BLOCKS + 31) / 32;
  double v[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int i = (int)threadIdx.x + 32 * k;
    v[k] = (i < LR_ERR_BLOCKS) ? __ldcg(p + i) : 0.0;
  }
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < NV; ++k)
    if ((int)threadIdx.x + 32 * k < LR_ERR_BLOCKS) s += v[k];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
  return __shfl_sync(0xffffffffu, s, 0);
}

__device__ __forceinline__ void lr_desc_clear(LinComb& d) {
  d.base = nullptr; d.dst = nullptr; d.n = 0; d.scale = 0.0f; d.t = 0.0f; d.pad_ = 0;
  for (int k = 0; k < LR_MAXSRC; ++k) { d.src[k] = nullptr; d.coef[k] = 0.0f; }
}

// interpolant of the dense forward solution at tval (oracle ODESolution.__call__)
__device__ inline void lr_write_interp(const SolveDev* S, LinComb& d, float tval) {
  lr_desc_clear(d);
  int n = lr_locate(S->fts, S->fnsteps, S->ftdir, tval);
  float dt = lr_sub(S->fts[n + 1], S->fts[n]);
  float th = lr_div(lr_sub(tval, S->fts[n]), dt);
  float b[7];
  lr_tsit5_interp(th, b);
  const float* slot = S->ftape + (size_t)n * 7 * S->flen;
  d.base = slot;
  for (int k = 0; k < 6; ++k) d.src[k] = slot + (size_t)(k + 1) * S->flen;
  d.src[6] = slot + (size_t)7 * S->flen + S->flen;  // k7 == k1 of the next slot
  for (int k = 0; k < 7; ++k) d.coef[k] = b[k];
  d.scale = dt;
  d.n = 7;
  d.t = tval;
  d.pad_ = __float_as_int(S->fts[n]);   // start time of the forward interval (latent-space adjoint: stage times)
}

// Stage descriptors of the attempt about to run from (slot, c.t, c.dt).
__device__ inline void lr_write_stage_descs(SolveDev* S) {
  const int s = S->slot, s1 = lr_next_slot(S, s);
  const float t = S->c.t, dt = S->c.dt;
  const float* uprev = lr_slot_u(S, s);
  for (int j = 0; j < 6; ++j) {  // row j: stage j+2
    LinComb& d = S->st[j];
    lr_desc_clear(d);
    d.base = uprev;
    d.n = j + 1;
    for (int i = 0; i <= j; ++i) {
      d.src[i] = lr_slot_k(S, s, i + 1);
      d.coef[i] = lr_tsit5_a(j, i);
    }
    d.scale = dt;
    d.t = lr_add(t, lr_mul(lr_tsit5_c(j), dt));
    d.dst = (j < 5) ? lr_slot_k(S, s, j + 2) : lr_slot_u(S, s1);
  }
  LinComb& k7 = S->st[6];
  lr_desc_clear(k7);
  k7.base = lr_slot_u(S, s1);
  k7.t = lr_add(t, dt);
  k7.dst = lr_slot_k(S, s1, 1);
  LinComb& e = S->err;
  lr_desc_clear(e);
  e.base = uprev;
  e.dst = lr_slot_u(S, s1);
  e.n = 7;
  for (int i = 0; i < 6; ++i) e.src[i] = lr_slot_k(S, s, i + 1);
  e.src[6] = lr_slot_k(S, s1, 1);
  for (int i = 0; i < 7; ++i) e.coef[i] = lr_tsit5_btilde(i);
  e.scale = dt;
  if (S->is_adjoint) {
    for (int j = 0; j < 5; ++j) lr_write_interp(S, S->yint[j + 1], S->st[j].t);
    lr_write_interp(S, S->yint[6], k7.t);
  }
}

// header of the next attempt (or end of segment / failure / tape full)
__device__ inline void lr_begin_attempt(SolveDev* S) {
  const float td = (float)S->c.tdir;
  const bool more = lr_mul(td, S->c.t) < lr_mul(td, S->c.tstop);
  if (more && !S->ring && S->slot + 1 >= S->cap) {
    S->tape_full = 1;
    S->done = 1;
    return;
  }
  if (lr_ctrl_header(S->c, S->u_nan)) {
    lr_write_stage_descs(S);
    S->done = 0;
  } else {
    S->done = 1;
    if (S->c.retcode != LR_RET_SUCCESS) S->failed = 1;
  }
}

__device__ __forceinline__ void lr_set_cond(SolveDev* S) {
  if (S->use_cond) cudaGraphSetConditional((cudaGraphConditionalHandle)S->cond_handle, S->done ? 0u : 1u);
}

// Descriptor for evaluating fsalfirst = f(u, t) at the current state.
__global__ void k1_desc_kernel(SolveDev* S) {
  if (threadIdx.x != 0) return;
  LinComb& d = S->st[6];
  lr_desc_clear(d);
  d.base = lr_slot_u(S, S->slot);
  d.t = S->c.t;
  d.dst = lr_slot_k(S, S->slot, 1);
  LinComb& c = S->cur;
  lr_desc_clear(c);
  c.base = lr_slot_u(S, S->slot);
  c.t = S->c.t;
  if (S->is_adjoint) {
    lr_write_interp(S, S->yint[6], S->c.t);
    lr_write_interp(S, S->yint[0], S->c.t);   // kept until the next call (latent-space adjoint: stage 1 of the segment)
  }
}

// initdt part A (after initdt_norm1): dt0 and the Euler probe descriptor (st[0])
__global__ void initdt_a_kernel(SolveDev* S) {
  if (S->failed) return;
  double v[4];
  v[0] = lr_sum_partials(S->partials);
  v[1] = lr_sum_partials(S->partials + LR_ERR_BLOCKS);
  if (threadIdx.x != 0) return;
  v[2] = (double)S->counters[0];
  v[3] = 0.0;
  lr_group_sum(S, v);
  if (S->reduce_mu && S->nranks > 1) {  // mu block from the exchanged global vector
    double m = 0.0;
    for (int i = 0; i < LR_ERR_BLOCKS; ++i) m += S->partials[3 * LR_ERR_BLOCKS + i];
    v[1] += m;
    v[2] += (double)S->counters[3];
    S->counters[3] = 0;
  }
  S->nf += 2;
  const double n = (double)S->total_len;
  float d0 = sqrtf((float)v[0] / (float)n);
  float d1 = sqrtf((float)v[1] / (float)n);
  S->d0 = d0;
  S->d1 = d1;
  S->u_nan = isnan(d0) ? 1 : 0;
  S->counters[2] = (v[2] > 0.0) ? 1u : 0u;  // non-finite f0: initdt returns tdir*dtmin
  float dt0 = lr_initdt_a(d0, d1, S->c.dtmax, S->c.tdir);
  S->dt0 = dt0;
  LinComb& d = S->st[0];
  lr_desc_clear(d);
  d.base = lr_slot_u(S, S->slot);
  d.src[0] = lr_slot_k(S, S->slot, 1);
  d.coef[0] = 1.0f;
  d.n = 1;
  d.scale = dt0;
  d.t = lr_add(S->c.t, dt0);
  d.dst = lr_slot_k(S, S->slot, 2);
  if (S->is_adjoint) lr_write_interp(S, S->yint[1], d.t);
}

// initdt part B (after initdt_norm2) + header of the first attempt
__global__ void initdt_b_kernel(SolveDev* S, int begin) {
  if (S->failed) return;
  double v[4];
  v[0] = lr_sum_partials(S->partials + 2 * LR_ERR_BLOCKS);
  if (threadIdx.x != 0) return;
  v[1] = (double)S->counters[1];
  v[2] = v[3] = 0.0;
  lr_group_sum(S, v);
  if (S->reduce_mu && S->nranks > 1) {
    double m = 0.0;
    for (int i = 0; i < LR_ERR_BLOCKS; ++i) m += S->partials[3 * LR_ERR_BLOCKS + i];
    v[0] += m;
    v[1] += (double)S->counters[3];
    S->counters[3] = 0;
  }
  float d2rms = sqrtf((float)v[0] / (float)(double)S->total_len);
  float dt;
  if (S->counters[2]) dt = lr_mul((float)S->c.tdir, S->c.dtmin);
  else dt = lr_initdt_b(S->dt0, S->d1, d2rms, v[1] == 0.0, S->c.dtmax, S->c.dtmin, S->c.tdir);
  S->c.dt = dt;
  S->c.dtpropose = dt;
  if (begin) {
    lr_begin_attempt(S);
  } else {
    // regulariser integrator: the reference reads integrator.dt straight after init
    // (src/perform_step.jl:4) -- no loopheader
    lr_write_stage_descs(S);
  }
}

// footer of the attempt that just ran + header of the next one
__global__ void controller_kernel(SolveDev* S) {
  if (S->done) { if (threadIdx.x == 0) lr_set_cond(S); return; }
  double v[4];
  v[0] = lr_sum_partials(S->partials);
  // mu block of the adjoint state: row 1 = latent-space engine (adj_mu_kernel / adj_mu_finish_kernel), row 3 = the
  // exchanged global vectors of the per-layer engine; in a group both are already global and identical on every rank
  v[1] = S->lat_mu_row ? lr_sum_partials(S->partials + LR_ERR_BLOCKS)
                       : ((S->reduce_mu && S->nranks > 1) ? lr_sum_partials(S->partials + 3 * LR_ERR_BLOCKS) : 0.0);
  if (threadIdx.x != 0) return;
  v[2] = v[3] = 0.0;
  if (S->lat_mu_row && S->reduce_mu && S->nranks > 1) S->mseq += 1;   // next vector exchange uses the other parity
  { double w[4] = {v[0], 0, 0, 0}; lr_group_sum(S, w); v[0] = w[0] + v[1]; }
  if (S->failed) { S->done = 1; lr_set_cond(S); return; }   // peer timeout (or a failure flagged by a step kernel)
  float EEst = sqrtf((float)v[0] / (float)(double)S->total_len);
  const float t_before = S->c.t, dt_before = S->c.dt;
  float dt_taken = 0.0f;
  int accept = lr_ctrl_footer(S->c, EEst, &dt_taken);
  S->nf += 6;
  if (S->nlog < S->logcap) {
    S->log_t[S->nlog] = t_before;
    S->log_dt[S->nlog] = dt_before;
    S->log_eest[S->nlog] = EEst;
    S->log_acc[S->nlog] = (unsigned char)accept;
  }
  S->nlog += 1;
  if (accept) {
    S->slot = lr_next_slot(S, S->slot);
    if (!S->ring && S->ts) S->ts[S->slot] = S->c.t;
  }
  lr_begin_attempt(S);
  lr_set_cond(S);
}

// (re)start a segment towards `tstop` (adjoint: after a lambda jump; forward: after growing
// the tape when tstop_valid == 0)
__global__ void segment_begin_kernel(SolveDev* S, float tstop, int set_tstop) {
  if (threadIdx.x != 0) return;
  if (S->failed) { S->done = 1; return; }
  if (set_tstop) S->c.tstop = tstop;
  S->tape_full = 0;
  lr_begin_attempt(S);
}

// ------------------------------------------------------------------------------------------
// Regulariser read-out (perform_step.jl:34-47) on the regulariser integrator R (a 2-slot ring:
// uprev = U(0), k1..k6 = K(0,.), u = U(1), k7 = K(1,1)); and the seeds of its reverse pass.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) reg_stiff_sums_kernel(SolveDev* R) {
  if (R->failed) return;
  __shared__ LinComb g6d;
  if (threadIdx.x == 0) g6d = R->st[4];
  __syncthreads();
  const float* u = R->err.dst;
  const float* k6 = R->err.src[5];
  const float* k7 = R->err.src[6];
  double a = 0.0, b = 0.0;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < R->len;
       i += (size_t)gridDim.x * blockDim.x) {
    float d = u[i] - lr_lincomb_at(g6d, i);
    float kk = k7[i] - k6[i];
    a += (double)(d * d);
    b += (double)(kk * kk);
  }
  double sa = lr_block_sum(a);
  double sb = lr_block_sum(b);
  if (threadIdx.x == 0) {
    R->partials[LR_ERR_BLOCKS + blockIdx.x] = sa;
    R->partials[2 * LR_ERR_BLOCKS + blockIdx.x] = sb;
  }
}

__global__ void reg_value_kernel(SolveDev* R, int reg_type) {
  if (R->failed) { if (threadIdx.x == 0) R->reg_val = NAN; return; }
  double v[4];
  v[0] = lr_sum_partials(R->partials);
  v[1] = reg_type ? lr_sum_partials(R->partials + LR_ERR_BLOCKS) : 0.0;
  v[2] = reg_type ? lr_sum_partials(R->partials + 2 * LR_ERR_BLOCKS) : 0.0;
  if (threadIdx.x != 0) return;
  v[3] = 0.0;
  lr_group_sum(R, v);
  const float n = (float)(double)R->total_len;
  R->nf += 6;
  if (reg_type == 0) {
    float root = sqrtf((float)v[0] / n);
    R->reg_ss_root = root;
    R->reg_val = root * R->c.dt;
  } else {
    float den = sqrtf((float)v[1] / n);
    float num = sqrtf((float)v[2] / n);
    R->reg_aux[0] = den;
    R->reg_aux[1] = num;
    R->reg_val = (den == 0.0f) ? 0.0f : fabsf(num / (den + 1.1920929e-07f)) / 3.5068f;
  }
}

// Seeds of the reverse pass of _perform_step w.r.t. the parameters (uprev, k1, dt constant:
// neural_ode.jl:40, utils.jl:60).  dk[0..5] <- cotangents of k2..k7, du <- of u, dg6 <- of g6.
struct RegSeedP {
  SolveDev* R; int reg_type; float d_reg;
  float* dk[6]; float* du; float* dg6;
};
__global__ void __launch_bounds__(256) reg_seed_kernel(RegSeedP p) {
  SolveDev* R = p.R;
  __shared__ LinComb e, g6d;
  if (threadIdx.x == 0) { e = R->err; g6d = R->st[4]; }
  __syncthreads();
  const float n = (float)(double)R->total_len;
  const float dt = R->c.dt, abstol = R->abstol, reltol = R->reltol;
  const float* uprev = e.base;
  const float* u = e.dst;
  float c = 0.0f, dnum_s = 0.0f, dden_s = 0.0f;
  if (p.reg_type == 0) {
    float root = R->reg_ss_root;
    c = (root == 0.0f || R->failed) ? 0.0f : p.d_reg * dt / (n * root);
  } else {
    float den = R->reg_aux[0], num = R->reg_aux[1];
    const float eps = 1.1920929e-07f;
    if (den != 0.0f && !R->failed) {
      float cc = p.d_reg / 3.5068f;
      float dnum = cc / (den + eps);
      float dden = -cc * num / ((den + eps) * (den + eps));
      dnum_s = (num != 0.0f) ? dnum / (n * num) : 0.0f;
      dden_s = dden / (n * den);
    }
  }
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < R->len;
       i += (size_t)gridDim.x * blockDim.x) {
    if (p.reg_type == 0) {
      float inner = 0.0f;
#pragma unroll
      for (int k = 0; k < 7; ++k) inner = fmaf(e.coef[k], e.src[k][i], inner);
      float ut = dt * inner;
      float au = fabsf(u[i]), ap = fabsf(uprev[i]);
      float denom = abstol + fmaxf(ap, au) * reltol;
      float r = ut / denom;
      float dr = c * r;
      float dut = dr / denom;
      float sg = (u[i] > 0.0f) ? 1.0f : ((u[i] < 0.0f) ? -1.0f : 0.0f);
      p.du[i] = (-dr * ut / (denom * denom)) * reltol * sg * ((au > ap) ? 1.0f : 0.0f);
#pragma unroll
      for (int k = 0; k < 6; ++k) p.dk[k][i] = dt * e.coef[k + 1] * dut;
      p.dg6[i] = 0.0f;
    } else {
      float d = u[i] - lr_lincomb_at(g6d, i);
      float kk = e.src[6][i] - e.src[5][i];
      float dkk = dnum_s * kk;
      float dd = dden_s * d;
#pragma unroll
      for (int k = 0; k < 4; ++k) p.dk[k][i] = 0.0f;
      p.dk[4][i] = -dkk;
      p.dk[5][i] = dkk;
      p.du[i] = dd;
      p.dg6[i] = -dd;
    }
  }
}

// after the VJP of stage `row+2` (row = 5: k7 = f(u); row 0..4: k_{row+2} = f(g)) produced `a`:
//   row == 5:  du += a;  dk_i += dt*a7i*du  (i = 2..6)
//   row <  5:  (row == 4: a += dg6);  dk_i += dt*a_{row+2,i}*a  (i = 2..row+1)
struct RegBwdP {
  SolveDev* R; int row;
  const float* a; float* du; const float* dg6; float* dk[6];
  float coef[6];  // tableau row a_{row+2, .} (host-filled: no local-memory table in the loop)
};
__global__ void __launch_bounds__(256) reg_bwd_stage_kernel(RegBwdP p) {
  const float dt = p.R->c.dt;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < p.R->len;
       i += (size_t)gridDim.x * blockDim.x) {
    float a = p.a[i];
    int nk;
    if (p.row == 5) { a += p.du[i]; p.du[i] = a; nk = 5; }
    else { if (p.row == 4) a += p.dg6[i]; nk = p.row; }
    // k index i+1 (i>=1) receives dt * a[row][i] * a ; dk[] starts at k2
#pragma unroll
    for (int k = 1; k <= 5; ++k)
      if (k <= nk) p.dk[k - 1][i] = fmaf(dt * p.coef[k], a, p.dk[k - 1][i]);
  }
}

// transpose of the weight block without its time column / bias:  WT[in x out] <- W[out x in]
__global__ void transpose_kernel(const float* W, int out, int in, float* WT) {
  __shared__ float tile[32][33];
  int o0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int i = i0 + r, o = o0 + threadIdx.x;
    tile[r][threadIdx.x] = (i < in && o < out) ? W[(size_t)i * out + o] : 0.0f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int o = o0 + r, i = i0 + threadIdx.x;
    if (i < in && o < out) WT[(size_t)o * in + i] = tile[threadIdx.x][r];
  }
}

__global__ void fill_kernel(float* p, float v, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    p[i] = v;
}
