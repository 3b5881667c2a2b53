// Conv-dynamics engine of the neural-ODE path (BASELINE configs[3] "cifar10"):
//
//   TDChain(Chain(Chain(Conv((3,3), 9 => 64; pad=1, use_bias=false), BatchNorm(64, gelu)),
//                 Chain(Conv((3,3), 65 => 64; pad=1, use_bias=false), BatchNorm(64, gelu)),
//                 Conv((3,3), 65 => 8; pad=1, use_bias=false)))
//                                                   experiments/src/construct.jl:212-218
//
// evaluated behind the same LinComb-descriptor contract as MlpEval::forward / vjp, so that the
// on-device controller, the CUDA-graph WHILE loop, the tape, the regulariser and the adjoint are
// shared with the Dense path.  FP32 SIMT direct convolutions (first correct path of this row;
// the implicit-GEMM tcgen05 version is the next step):
//
//   * state / activations in the reference's WHCN column-major layout [Wd, Ht, C, B]
//   * the stage linear combination is formed while the first convolution loads its input patch
//     (perform_step.jl:11-18), the TDChain time channel (common.jl:19-33: t * ones concatenated on
//     the channel dimension before every layer) is a virtual channel: t inside the image, 0 in the
//     padding -- it is NOT a bias because of the zero padding
//   * NNlib `conv` is a true convolution (flipped kernel): y[o] = sum_k w[k] x[o + 1 - k]
//   * BatchNorm in training mode (batch mean / biased variance over W,H,B, eps 1e-5): the conv
//     epilogue emits per-block (sum, sum of squares), bn_finalize_kernel turns them into a per-channel
//     scale/shift that the NEXT convolution applies (with the activation) while loading its patch,
//     so the normalised activation is never written
//   * reverse pass (ZygoteVJP of the closure): recompute, then per layer the weight gradient
//     (split over the batch, fixed-order reduce), the transposed convolution and the BatchNorm
//     pullback (two per-channel sums, then an in-place element-wise pass)
#pragma once
#include <cuda_runtime.h>

#include "lrnde_kernels.cuh"

struct ConvP {
  const float* X; const LinComb* xdesc;  // input [Wd,Ht,Cin,B]: plain buffer or stage combination
  const LinComb* side_desc; float* side; // also store the combined input (u_{n+1} / y(t))
  const float* in_ab; int in_act;        // per-channel a[Cin], b[Cin] and activation applied on load
  int td; const LinComb* tdesc;          // virtual time channel (index Cin) = tdesc->t
  const float* Wp;                       // packed weights [9][CinTot][Cout]
  int Wd, Ht, Cin, Cout, B;
  float* Y; const LinComb* ydesc; float out_scale;
  const float* bias; int out_act; float* pre;   // standalone Conv layer: y = act(conv + bias), pre-activation kept
  float2* stat_part;                     // [gridDim.x][Cout] (sum, sum of squares) of the raw output
  const int* done;
};

// Lux weight w[kx,ky,ci,co] (kx fastest) -> Wp[tap][ci][co], tap = (dy+1)*3 + (dx+1) the spatial offset
// of the input element: kx = 1 - dx, ky = 1 - dy.  transposed: the data-gradient convolution
// g_x[i,ci] = sum_d sum_co Wp[-d][ci][co] delta[i + d, co] over the first `keep` input channels.
__global__ void conv_pack_kernel(const float* w, int CinTot, int Cout, int transposed, int keep, float* out) {
  const int n = transposed ? 9 * Cout * keep : 9 * CinTot * Cout;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    int tap, ci, co, kx, ky;
    if (!transposed) {
      co = e % Cout; ci = (e / Cout) % CinTot; tap = e / (Cout * CinTot);
      kx = 2 - tap % 3; ky = 2 - tap / 3;
    } else {
      ci = e % keep; co = (e / keep) % Cout; tap = e / (keep * Cout);
      kx = tap % 3; ky = tap / 3;
    }
    out[e] = w[kx + 3 * (ky + 3 * (ci + (size_t)CinTot * co))];
  }
}

// One CTA: TR rows x Wd columns of one image x CB output channels; a thread owns 4 consecutive pixels of a
// row x TC output channels.  Input channels stream through shared memory CK at a time.
template <int CB, int TC, int CK>
__global__ void __launch_bounds__(256) conv3x3_kernel(ConvP p) {
  if (p.done && *p.done) return;
  constexpr int NCG = CB / TC, PT = 256 / NCG;
  constexpr int PATCH = (PT + 2) * 7;   // (TR + 2) * (Wd + 3) is largest at Wd = 4
  __shared__ float xs[CK][PATCH];
  __shared__ __align__(16) float ws[9][CK][CB];
  __shared__ LinComb xd;
  __shared__ float tval;
  __shared__ float red[2][8][TC];
  const int tid = threadIdx.x;
  const int Wd = p.Wd, Ht = p.Ht, CGR = Wd >> 2, RS = Wd + 3;
  const int TR = min(PT / CGR, Ht);
  const int tilesY = (Ht + TR - 1) / TR;
  const int b = blockIdx.x / tilesY, row0 = (blockIdx.x % tilesY) * TR;
  const int co0 = blockIdx.y * CB;
  const int cg = tid / PT, pt = tid % PT;
  const int prow = pt / CGR, pcol = (pt % CGR) * 4;
  const bool active = prow < TR && (row0 + prow) < Ht;
  const int CinTot = p.Cin + (p.td ? 1 : 0);
  if (tid == 0) {
    if (p.xdesc) xd = *p.xdesc;
    tval = p.tdesc ? p.tdesc->t : 0.0f;
  }
  float acc[4][TC];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TC; ++j) acc[i][j] = 0.0f;
  const int prow_n = TR + 2, pw = Wd + 2;
  float* side_out = p.side_desc ? nullptr : p.side;

  for (int c0 = 0; c0 < CinTot; c0 += CK) {
    __syncthreads();
    if (p.side_desc && !side_out) side_out = xd.dst;   // xd is visible after the first barrier
    {   // patch: one (or half a) warp per input channel, lanes over the (row, x) pairs; the row comes from a
        // multiply-shift (exact for e < 2048 and every pw = Wd + 2 <= 34; checked exhaustively) instead of integer divisions
      constexpr int WPC = 8 / CK;
      const int c = (tid >> 5) / WPC;
      const int gc = c0 + c;
      const unsigned inv_pw = (65536u + pw - 1) / pw;
#pragma unroll 4
      for (int e = (tid & 31) + 32 * ((tid >> 5) % WPC); e < prow_n * pw; e += 32 * WPC) {
        const int r = (int)(((unsigned)e * inv_pw) >> 16), x = e - r * pw;
        const int gy = row0 + r - 1, gx = x - 1;
        float v = 0.0f;
        if (gc < CinTot && gy >= 0 && gy < Ht && gx >= 0 && gx < Wd) {
          if (gc == p.Cin) v = tval;
          else {
            const size_t idx = gx + (size_t)Wd * (gy + (size_t)Ht * (gc + (size_t)p.Cin * b));
            if (p.xdesc) {
              v = lr_lincomb_at(xd, idx);
              if (side_out && blockIdx.y == 0 && r >= 1 && r <= TR) side_out[idx] = v;
            } else v = p.X[idx];
            if (p.in_ab) v = fmaf(p.in_ab[gc], v, p.in_ab[p.Cin + gc]);
            if (p.in_act != ACT_IDENTITY) v = lr_act(p.in_act, v);
          }
        }
        xs[c][r * RS + x] = v;
      }
    }
    for (int e = tid; e < 9 * CK * CB; e += 256) {
      const int j = e % CB, c = (e / CB) % CK, tap = e / (CB * CK);
      float v = 0.0f;
      if (c0 + c < CinTot && co0 + j < p.Cout) v = p.Wp[((size_t)tap * CinTot + c0 + c) * p.Cout + co0 + j];
      ws[tap][c][j] = v;
    }
    __syncthreads();
    if (active) {
#pragma unroll
      for (int c = 0; c < CK; ++c) {
        if (c0 + c >= CinTot) break;
        float xin[3][6];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
          for (int j = 0; j < 6; ++j) xin[dy][j] = xs[c][(prow + dy) * RS + pcol + j];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
          for (int dx = 0; dx < 3; ++dx) {
            const float4* w4 = reinterpret_cast<const float4*>(&ws[dy * 3 + dx][c][cg * TC]);
#pragma unroll
            for (int q = 0; q < TC / 4; ++q) {
              const float4 w = w4[q];
#pragma unroll
              for (int pp = 0; pp < 4; ++pp) {
                const float xv = xin[dy][pp + dx];
                acc[pp][4 * q + 0] = fmaf(xv, w.x, acc[pp][4 * q + 0]);
                acc[pp][4 * q + 1] = fmaf(xv, w.y, acc[pp][4 * q + 1]);
                acc[pp][4 * q + 2] = fmaf(xv, w.z, acc[pp][4 * q + 2]);
                acc[pp][4 * q + 3] = fmaf(xv, w.w, acc[pp][4 * q + 3]);
              }
            }
          }
      }
    }
  }

  float* out = p.ydesc ? p.ydesc->dst : p.Y;
  const bool out_vec = (((uintptr_t)out) & 15) == 0;
  if (active) {
#pragma unroll
    for (int j = 0; j < TC; ++j) {
      const int co = co0 + cg * TC + j;
      if (co < p.Cout) {
        const size_t idx = pcol + (size_t)Wd * ((row0 + prow) + (size_t)Ht * (co + (size_t)p.Cout * b));
        float4 v = make_float4(acc[0][j] * p.out_scale, acc[1][j] * p.out_scale, acc[2][j] * p.out_scale,
                               acc[3][j] * p.out_scale);
        if (p.bias) { const float bv = p.bias[co]; v.x += bv; v.y += bv; v.z += bv; v.w += bv; }
        if (p.pre) *reinterpret_cast<float4*>(p.pre + idx) = v;
        if (p.out_act != ACT_IDENTITY) {
          v.x = lr_act(p.out_act, v.x); v.y = lr_act(p.out_act, v.y); v.z = lr_act(p.out_act, v.z); v.w = lr_act(p.out_act, v.w);
        }
        // the lambda block of the adjoint state sits at a multiple of (D*B + nparams) floats: not always 16-byte aligned
        if (out_vec) *reinterpret_cast<float4*>(out + idx) = v;
        else { out[idx] = v.x; out[idx + 1] = v.y; out[idx + 2] = v.z; out[idx + 3] = v.w; }
      }
    }
  }
  if (p.stat_part) {   // per-channel (sum, sum of squares) of this tile: warp shuffle, then the warps of a group
    const int warp = tid >> 5, lane = tid & 31;
#pragma unroll
    for (int j = 0; j < TC; ++j) {
      float s1 = (acc[0][j] + acc[1][j]) + (acc[2][j] + acc[3][j]);
      float s2 = fmaf(acc[0][j], acc[0][j], acc[1][j] * acc[1][j]) + fmaf(acc[2][j], acc[2][j], acc[3][j] * acc[3][j]);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
      }
      if (lane == 0) { red[0][warp][j] = s1; red[1][warp][j] = s2; }
    }
    __syncthreads();
    if (tid < CB && co0 + tid < p.Cout) {
      constexpr int WPG = PT / 32;   // warps per output-channel group
      const int g = tid / TC, j = tid % TC;
      float s1 = 0.0f, s2 = 0.0f;
      for (int w = 0; w < WPG; ++w) { s1 += red[0][g * WPG + w][j]; s2 += red[1][g * WPG + w][j]; }
      p.stat_part[(size_t)blockIdx.x * p.Cout + co0 + tid] = make_float2(s1, s2);
    }
  }
}

// ------------------------------------------------------------------------------------------
// weight gradient: dW[kx,ky,ci,co] = sum_{b,o} x[o + d, ci, b] * delta[o, co, b]
//
// Two operands: P ("patch", with a one-pixel halo, <= 16 channels per CTA) and Q ("tile", no halo, 64 channels
// per CTA, stored pixel-major so that a thread reads its 4 channels with one 16-byte load).  A thread owns one P
// channel x 4 Q channels x the 9 taps and walks the pixels of a row strip with a sliding window over P:
//   acc[tap][j] = sum_px P[px + d_tap][pc] * Q[px][qc_j]
// normal : P = x (input, transformed on load), Q = delta           -> dW[2-dx, 2-dy][ci = pc][co = qc]
// swapped: P = delta, Q = x  (few output channels: 65 => 8)        -> dW[dx, dy][ci = qc][co = pc]
// grid (P-channel chunks, batch splits, 64-wide Q blocks); fixed-order reduce over the splits afterwards
// ------------------------------------------------------------------------------------------
struct ConvWgOp {
  const float* ptr; const float* ab; int act; int C; int td;   // [Wd,Ht,C,B]; per-channel a,b + activation on load; virtual time channel
};
struct ConvWgP {
  ConvWgOp P, Q;
  int swapped;
  const LinComb* tdesc;
  int Wd, Ht, B;
  int pcc, img_per_split;      // P channels per CTA (<= 16)
  int CinTot, Cout;            // layout of the result
  float* part; size_t block;   // part[split][block], Lux weight layout
  const int* done;
};

__device__ __forceinline__ float conv_wg_load(const ConvWgOp& o, int ch, size_t pix, int b, size_t HW, float tval) {
  if (ch == o.C) return tval;
  float v = o.ptr[pix + HW * (ch + (size_t)o.C * b)];
  if (o.ab) v = fmaf(o.ab[ch], v, o.ab[o.C + ch]);
  if (o.act != ACT_IDENTITY) v = lr_act(o.act, v);
  return v;
}

#define CWG_NPX 96
#define CWG_QS 68     // Q tile row stride (floats): 16-byte aligned rows, 4-way conflict on the (rare) stores only
__global__ void __launch_bounds__(256) conv3x3_wgrad_kernel(ConvWgP p) {
  if (p.done && *p.done) return;
  __shared__ __align__(16) float qs[CWG_NPX * CWG_QS];
  __shared__ float xs[16][184];
  __shared__ float tval;
  const int tid = threadIdx.x;
  const int Wd = p.Wd, Ht = p.Ht, RS = Wd + 3, pw = Wd + 2;
  const int SR = max(1, min(Ht, CWG_NPX / Wd)), NPX = SR * Wd;
  const int PT = p.P.C + p.P.td, QT = p.Q.C + p.Q.td;
  const int c0 = blockIdx.x * p.pcc;
  const int ncl = min(p.pcc, PT - c0);
  const int q0 = blockIdx.z * 64;
  const int pc_l = tid >> 4, cog = tid & 15;
  const int b0 = blockIdx.y * p.img_per_split, b1 = min(p.B, b0 + p.img_per_split);
  if (tid == 0) tval = p.tdesc ? p.tdesc->t : 0.0f;
  float acc[9][4];
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[k][j] = 0.0f;
  const size_t HW = (size_t)Wd * Ht;
  const bool qactive = q0 + cog * 4 < QT;       // warps whose threads are all idle skip the pixel loop
  for (int b = b0; b < b1; ++b) {
    for (int row0 = 0; row0 < Ht; row0 += SR) {
      __syncthreads();
      for (int c = tid >> 5; c < ncl; c += 8) {        // patch operand: a warp per channel, lanes along x
        for (int r = 0; r < SR + 2; ++r) {
          const int gy = row0 + r - 1;
          const bool row_ok = gy >= 0 && gy < Ht;
          for (int x = tid & 31; x < pw; x += 32) {
            const int gx = x - 1;
            float v = 0.0f;
            if (row_ok && gx >= 0 && gx < Wd) v = conv_wg_load(p.P, c0 + c, gx + (size_t)Wd * gy, b, HW, tval);
            xs[c][r * RS + x] = v;
          }
        }
      }
      for (int c = tid >> 5; c < 64; c += 8) {         // tile operand: lanes along the pixels of the strip
        const bool ch_ok = q0 + c < QT;
        for (int px = tid & 31; px < NPX; px += 32) {
          const size_t pix = (size_t)row0 * Wd + px;
          float v = 0.0f;
          if (ch_ok && pix < HW) v = conv_wg_load(p.Q, q0 + c, pix, b, HW, tval);
          qs[px * CWG_QS + c] = v;
        }
      }
      __syncthreads();
      if (pc_l < ncl && qactive) {
        for (int r = 0; r < SR; ++r) {
          float x0[3], x1[3], x2[3];
#pragma unroll
          for (int dy = 0; dy < 3; ++dy) { x0[dy] = xs[pc_l][(r + dy) * RS]; x1[dy] = xs[pc_l][(r + dy) * RS + 1]; }
#pragma unroll 4
          for (int x = 0; x < Wd; ++x) {
            const float4 d4 = *reinterpret_cast<const float4*>(&qs[(r * Wd + x) * CWG_QS + cog * 4]);
            const float d[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
            for (int dy = 0; dy < 3; ++dy) {
              x2[dy] = xs[pc_l][(r + dy) * RS + x + 2];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                acc[dy * 3 + 0][j] = fmaf(x0[dy], d[j], acc[dy * 3 + 0][j]);
                acc[dy * 3 + 1][j] = fmaf(x1[dy], d[j], acc[dy * 3 + 1][j]);
                acc[dy * 3 + 2][j] = fmaf(x2[dy], d[j], acc[dy * 3 + 2][j]);
              }
              x0[dy] = x1[dy]; x1[dy] = x2[dy];
            }
          }
        }
      }
    }
  }
  if (pc_l < ncl) {
    float* out = p.part + (size_t)blockIdx.y * p.block;
    const int pc = c0 + pc_l;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int qc = q0 + cog * 4 + j;
      if (qc >= QT) continue;
      const int ci = p.swapped ? qc : pc, co = p.swapped ? pc : qc;
#pragma unroll
      for (int k = 0; k < 9; ++k) {
        const int kx = p.swapped ? (k % 3) : (2 - k % 3), ky = p.swapped ? (k / 3) : (2 - k / 3);
        out[kx + 3 * (ky + 3 * (ci + (size_t)p.CinTot * co))] = acc[k][j];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// BatchNorm (training mode: batch statistics over W, H, B)
// ------------------------------------------------------------------------------------------
// per-channel scale / shift from the conv-epilogue partials: ab = [gamma * invstd ; beta - mean * gamma * invstd],
// stat = [mean ; invstd].  grid = C blocks of 128 threads; fixed partition, fixed-order tree => deterministic.
__global__ void __launch_bounds__(128) bn_finalize_kernel(const float2* part, int nblk, int C, double count,
                                                          const float* gamma_beta, float eps, float* ab,
                                                          float* stat, float* running, int testmode, int update,
                                                          const int* done, BnDist dist) {
  if (done && *done) return;
  __shared__ double s1[128], s2[128];
  const int c = blockIdx.x, tid = threadIdx.x;
  if (testmode) {   // Lux.testmode: normalise with the running statistics (running = [mean[C] ; var[C]])
    if (tid == 0) {
      const float mean = running[c];
      const float invstd = 1.0f / sqrtf(running[C + c] + eps);
      const float a = gamma_beta[c] * invstd;
      ab[c] = a; ab[C + c] = fmaf(-mean, a, gamma_beta[C + c]);
      stat[c] = mean; stat[C + c] = invstd;
    }
    return;
  }
  double a1 = 0.0, a2 = 0.0;
  for (int i = tid; i < nblk; i += 128) {
    const float2 v = part[(size_t)i * C + c];
    a1 += (double)v.x; a2 += (double)v.y;
  }
  s1[tid] = a1; s2[tid] = a2;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (tid < o) { s1[tid] += s1[tid + o]; s2[tid] += s2[tid + o]; }
    __syncthreads();
  }
  if (tid == 0) {
    // data-parallel group: the two sums and the element count are exchanged, so every rank
    // normalises with (and tracks) the statistics of the whole batch
    lr_bn_group_sum(dist, c, s1[0], s2[0], count);
    const double mean = s1[0] / count;
    double var = s2[0] / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float a = gamma_beta[c] * invstd;
    ab[c] = a;
    ab[C + c] = fmaf(-(float)mean, a, gamma_beta[C + c]);
    stat[c] = (float)mean;
    stat[C + c] = invstd;
    if (running) {   // Lux BatchNorm: momentum 0.1, running_var tracks the UNBIASED batch variance; `update` = how
                     // many closure calls this evaluation stands for (f(u0, t0) is called twice at the start of a solve)
      const float mom = 0.1f;
      for (int i = 0; i < update; ++i) {
        running[c] = (1.0f - mom) * running[c] + mom * (float)mean;
        running[C + c] = (1.0f - mom) * running[C + c] + mom * (float)(var * (count / (count - 1.0)));
      }
    }
  }
}

// pullback of h = act(a z + b), z -> (z - mean) invstd gamma + beta with batch statistics:
//   ghat = g * act'(a z + b), xhat = (z - mean) invstd
//   d_gamma = sum ghat xhat, d_beta = sum ghat, g_z = gamma invstd (ghat - mean(ghat) - xhat mean(ghat xhat))
// grid (C, S): partial sums of (ghat, ghat xhat) per channel over a slice of the batch
__global__ void __launch_bounds__(256) bn_bwd_stats_kernel(const float* g, const float* z, const float* ab,
                                                           const float* stat, int act, int C, size_t HW, int B,
                                                           int img_per_split, double2* part, const int* done) {
  if (done && *done) return;
  __shared__ double r1[8], r2[8];
  const int c = blockIdx.x, tid = threadIdx.x;
  const float a = ab[c], bb = ab[C + c], mean = stat[c], invstd = stat[C + c];
  const int b0 = blockIdx.y * img_per_split, b1 = min(B, b0 + img_per_split);
  double a1 = 0.0, a2 = 0.0;
  for (int b = b0; b < b1; ++b) {
    const size_t base = HW * (c + (size_t)C * b);
    float f1 = 0.0f, f2 = 0.0f;
    for (size_t i = tid; i < HW; i += 256) {
      const float zz = z[base + i];
      const float gh = g[base + i] * lr_dact(act, fmaf(a, zz, bb));
      f1 += gh;
      f2 = fmaf(gh, (zz - mean) * invstd, f2);
    }
    a1 += (double)f1; a2 += (double)f2;
  }
  for (int o = 16; o > 0; o >>= 1) {
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    a2 += __shfl_xor_sync(0xffffffffu, a2, o);
  }
  if ((tid & 31) == 0) { r1[tid >> 5] = a1; r2[tid >> 5] = a2; }
  __syncthreads();
  if (tid == 0) {
    double t1 = 0.0, t2 = 0.0;
    for (int w = 0; w < 8; ++w) { t1 += r1[w]; t2 += r2[w]; }
    part[(size_t)c * gridDim.y + blockIdx.y] = make_double2(t1, t2);
  }
}

// coef = [mean(ghat) ; mean(ghat xhat)], dgb = [d_gamma ; d_beta]
__global__ void bn_bwd_finalize_kernel(const double2* part, int S, int C, double count, float* coef, float* dgb,
                                       int testmode, const int* done, BnDist dist) {
  if (done && *done) return;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double t1 = 0.0, t2 = 0.0;
  for (int s = 0; s < S; ++s) { t1 += part[(size_t)c * S + s].x; t2 += part[(size_t)c * S + s].y; }
  dgb[c] = (float)t2;            // parameter gradients stay per-rank partial sums (all-reduced once per iteration)
  dgb[C + c] = (float)t1;
  if (!testmode) lr_bn_group_sum(dist, c, t1, t2, count);   // the means are over the GLOBAL batch
  coef[c] = testmode ? 0.0f : (float)(t1 / count);          // testmode: the statistics are constants
  coef[C + c] = testmode ? 0.0f : (float)(t2 / count);
}

// the same from the (sum ghat, sum ghat xhat) partials of the tensor-core data-gradient convolution's epilogue
// (part[block][C] float2, like bn_finalize_kernel); grid = C blocks of 128 threads
__global__ void __launch_bounds__(128) bn_bwd_finalize_tc_kernel(const float2* part, int nblk, int C, double count, float* coef,
                                                                 float* dgb, int testmode, const int* done, BnDist dist) {
  if (done && *done) return;
  __shared__ double s1[128], s2[128];
  const int c = blockIdx.x, tid = threadIdx.x;
  double a1 = 0.0, a2 = 0.0;
  for (int i = tid; i < nblk; i += 128) {
    const float2 v = part[(size_t)i * C + c];
    a1 += (double)v.x; a2 += (double)v.y;
  }
  s1[tid] = a1; s2[tid] = a2;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (tid < o) { s1[tid] += s1[tid + o]; s2[tid] += s2[tid + o]; }
    __syncthreads();
  }
  if (tid == 0) {
    double t1 = s1[0], t2 = s2[0];
    dgb[c] = (float)t2;
    dgb[C + c] = (float)t1;
    if (!testmode) lr_bn_group_sum(dist, c, t1, t2, count);
    coef[c] = testmode ? 0.0f : (float)(t1 / count);
    coef[C + c] = testmode ? 0.0f : (float)(t2 / count);
  }
}

// g <- gamma invstd (ghat - m1 - xhat m2) in place; coef == nullptr: plain activation pullback g <- g act'(z)
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(float* g, const float* z, const float* ab,
                                                           const float* stat, const float* coef, int act, int C,
                                                           size_t HW, size_t n, const int* done) {
  if (done && *done) return;
  // HW is a multiple of 4 (width % 4 == 0) and the buffers are 16-byte aligned: one float4 never straddles a channel
  const bool vec = ((HW & 3) == 0) && ((((uintptr_t)g) & 15) == 0) && ((((uintptr_t)z) & 15) == 0);
  if (vec) {
    const size_t n4 = n >> 2, hw4 = HW >> 2;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
      const int c = (int)((i / hw4) % C);
      const float4 zz = __ldcg(reinterpret_cast<const float4*>(z) + i);
      float4 gg = reinterpret_cast<float4*>(g)[i];
      if (coef) {
        const float a = ab[c], b = ab[C + c], mean = stat[c], invstd = stat[C + c], m1 = coef[c], m2 = coef[C + c];
        gg.x = a * (gg.x * lr_dact(act, fmaf(a, zz.x, b)) - m1 - (zz.x - mean) * invstd * m2);
        gg.y = a * (gg.y * lr_dact(act, fmaf(a, zz.y, b)) - m1 - (zz.y - mean) * invstd * m2);
        gg.z = a * (gg.z * lr_dact(act, fmaf(a, zz.z, b)) - m1 - (zz.z - mean) * invstd * m2);
        gg.w = a * (gg.w * lr_dact(act, fmaf(a, zz.w, b)) - m1 - (zz.w - mean) * invstd * m2);
      } else {
        gg.x *= lr_dact(act, zz.x); gg.y *= lr_dact(act, zz.y); gg.z *= lr_dact(act, zz.z); gg.w *= lr_dact(act, zz.w);
      }
      reinterpret_cast<float4*>(g)[i] = gg;
    }
    return;
  }
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)((i / HW) % C);
    const float zz = z[i];
    if (coef) {
      const float a = ab[c];
      const float gh = g[i] * lr_dact(act, fmaf(a, zz, ab[C + c]));
      const float xh = (zz - stat[c]) * stat[C + c];
      g[i] = a * (gh - coef[c] - xh * coef[C + c]);
    } else {
      g[i] = g[i] * lr_dact(act, zz);
    }
  }
}

// ------------------------------------------------------------------------------------------
// standalone BatchNorm layer (the BatchNorm(8) in front of the cifar10 NeuralODE, construct.jl:222)
// ------------------------------------------------------------------------------------------
// per-channel (sum, sum of squares) of x [HW, C, B] over a slice of the batch, in the layout bn_finalize_kernel
// reads: part[split][C].  grid (C, S)
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* x, int C, size_t HW, int B, int img_per_split,
                                                       float2* part) {
  __shared__ double r1[8], r2[8];
  const int c = blockIdx.x, tid = threadIdx.x;
  const int b0 = blockIdx.y * img_per_split, b1 = min(B, b0 + img_per_split);
  double a1 = 0.0, a2 = 0.0;
  for (int b = b0; b < b1; ++b) {
    const size_t base = HW * (c + (size_t)C * b);
    float f1 = 0.0f, f2 = 0.0f;
    for (size_t i = tid; i < HW; i += 256) { const float v = x[base + i]; f1 += v; f2 = fmaf(v, v, f2); }
    a1 += (double)f1; a2 += (double)f2;
  }
  for (int o = 16; o > 0; o >>= 1) {
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    a2 += __shfl_xor_sync(0xffffffffu, a2, o);
  }
  if ((tid & 31) == 0) { r1[tid >> 5] = a1; r2[tid >> 5] = a2; }
  __syncthreads();
  if (tid == 0) {
    double t1 = 0.0, t2 = 0.0;
    for (int w = 0; w < 8; ++w) { t1 += r1[w]; t2 += r2[w]; }
    part[(size_t)blockIdx.y * C + c] = make_float2((float)t1, (float)t2);
  }
}

// y = act(a[c] x + b[c])
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* x, const float* ab, int act, int C, size_t HW,
                                                       size_t n, float* y) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)((i / HW) % C);
    y[i] = lr_act(act, fmaf(ab[c], x[i], ab[C + c]));
  }
}

// dst[c] = sum over (HW, B) of g[., c, .]   (bias gradient of a Conv layer); grid C
__global__ void __launch_bounds__(256) channel_sum_kernel(const float* g, int C, size_t HW, int B, float* dst) {
  __shared__ double r1[8];
  const int c = blockIdx.x, tid = threadIdx.x;
  double a1 = 0.0;
  for (int b = 0; b < B; ++b) {
    const size_t base = HW * (c + (size_t)C * b);
    float f1 = 0.0f;
    for (size_t i = tid; i < HW; i += 256) f1 += g[base + i];
    a1 += (double)f1;
  }
  for (int o = 16; o > 0; o >>= 1) a1 += __shfl_xor_sync(0xffffffffu, a1, o);
  if ((tid & 31) == 0) r1[tid >> 5] = a1;
  __syncthreads();
  if (tid == 0) {
    double t1 = 0.0;
    for (int w = 0; w < 8; ++w) t1 += r1[w];
    dst[c] = (float)t1;
  }
}
