// Reverse pass of the regulariser's Tsit5 step w.r.t. the parameters in HIDDEN space, for the
// TDChain(Dense(D(+1) => H, tanh), Dense(H(+1) => D)) family and regularize_type = :error_estimate.
//
// Reference: Zygote's pullback of `_perform_step(integrator, cache, ps, Val(:error_estimate))`
// (src/perform_step.jl:3-38) with uprev, k1 and dt constants (neural_ode.jl:40, utils.jl:60), i.e. the discrete
// adjoint of the six stages k_j = f(g_j, ps, t_j), g_j = uprev + dt sum_i a_ji k_i, j = 2..7 (g_7 = u_{n+1}).
//
// With G = d reg / d utilde, Q = d reg / d u_{n+1} (through the residual's denominator), hh_j = [h_j ; t_j ; 1] the
// hidden images the forward chain kernel kept, s_j = 1 - h_j^2, Mh = W1[:, :D] W2 and
//     A = W2^T G,  Bq = W2^T Q                                                 (one D -> H GEMM over 2B columns)
// the cotangents of the stages are a recurrence in H dimensions (stage 7 first):
//     alpha_7 = dt bt_7 A,   alpha_j = dt (bt_j A + a_7j Bq + Mh^T sum_{m>j} a_mj delta_m),   delta_j = s_j .* alpha_j
// and the parameter gradient needs D-dimensional data only in two batch contractions:
//     dW2a        = [G Q] [HB HA]^T + W1^T (SS - dt UV hh_1^T)       HB = dt sum_j bt_j hh_j, HA = dt sum_{j<=6} a_7j hh_j
//     dW1[:, :D]  = (sum_j delta_j) uprev^T + SS W2a^T                  SS = sum_j delta_j c_j^T, c_j = dt sum_{i<j} a_ji hh_i
//     dW1[:, t/1] = sum_j delta_j [t_j ; 1]^T                           UV = sum_m a_m1 delta_m
// (the per-layer path does the same arithmetic as six layer-by-layer VJPs: 45 launches, 1.4 ms at B = 8192).
#pragma once
#include "lrnde_kernels.cuh"

// G and Q ([D, 2B]: G in the first B columns) -- reg_seed_kernel's arithmetic for regularize_type 0
struct RegSeedGQP {
  SolveDev* R; float d_reg; float* gq; size_t DB;
};
__global__ void __launch_bounds__(256) reg_seed_gq_kernel(RegSeedGQP p) {
  SolveDev* R = p.R;
  __shared__ LinComb e;
  if (threadIdx.x == 0) e = R->err;
  __syncthreads();
  const float n = (float)(double)R->total_len;
  const float dt = R->c.dt, abstol = R->abstol, reltol = R->reltol;
  const float* uprev = e.base;
  const float* u = e.dst;
  const float root = R->reg_ss_root;
  const float c = (root == 0.0f || R->failed) ? 0.0f : p.d_reg * dt / (n * root);
  auto elem = [&](float inner, float uu, float up, float& g, float& q) {
    const float ut = dt * inner;
    const float au = fabsf(uu), ap = fabsf(up);
    const float denom = abstol + fmaxf(ap, au) * reltol;
    const float r = ut / denom;
    const float dr = c * r;
    const float sg = (uu > 0.0f) ? 1.0f : ((uu < 0.0f) ? -1.0f : 0.0f);
    g = dr / denom;
    q = (-dr * ut / (denom * denom)) * reltol * sg * ((au > ap) ? 1.0f : 0.0f);
  };
  const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
  // 16-byte groups (every array of the regulariser's ring starts at a multiple of D * B floats, D % 4 == 0)
  uintptr_t al = ((uintptr_t)uprev) | ((uintptr_t)u) | ((uintptr_t)p.gq) | (uintptr_t)(p.DB & 3);
#pragma unroll
  for (int k = 0; k < 7; ++k) al |= (uintptr_t)e.src[k];
  size_t done = 0;
  if ((al & 15) == 0) {
    const size_t n4 = p.DB >> 2;
    for (size_t i = tid; i < n4; i += nthr) {
      float4 in = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(e.src[k]) + i);
        in.x = fmaf(e.coef[k], v.x, in.x); in.y = fmaf(e.coef[k], v.y, in.y);
        in.z = fmaf(e.coef[k], v.z, in.z); in.w = fmaf(e.coef[k], v.w, in.w);
      }
      const float4 uu = __ldcg(reinterpret_cast<const float4*>(u) + i), up = __ldcg(reinterpret_cast<const float4*>(uprev) + i);
      float4 g, q;
      elem(in.x, uu.x, up.x, g.x, q.x); elem(in.y, uu.y, up.y, g.y, q.y);
      elem(in.z, uu.z, up.z, g.z, q.z); elem(in.w, uu.w, up.w, g.w, q.w);
      reinterpret_cast<float4*>(p.gq)[i] = g;
      reinterpret_cast<float4*>(p.gq + p.DB)[i] = q;
    }
    done = n4 << 2;
  }
  for (size_t i = done + tid; i < p.DB; i += nthr) {
    float inner = 0.0f;
#pragma unroll
    for (int k = 0; k < 7; ++k) inner = fmaf(e.coef[k], __ldcg(e.src[k] + i), inner);
    elem(inner, u[i], uprev[i], p.gq[i], p.gq[p.DB + i]);
  }
}

struct RegRevP {
  const SolveDev* R;
  const float* Mz;       // plain [128][128]: Mz[r * 128 + k], r = hidden row, k = column of [W2 | w2t | b2]
  const float* ab;       // [2B][LR_ZROW]: A = W2^T G in the first B rows, Bq = W2^T Q after them
  const float* hh[7];    // [B][LR_ZROW]: [h_j ; (t_j) ; 1], j = 1..7
  float* del;            // [6B][LR_ZROW]: delta_j, j = 2..7 (block j - 2)
  float* cc;             // [6B][LR_ZROW]: [c_j ; (t_j) ; 1]
  float* dsum;           // [B][LR_ZROW]:  sum_j delta_j
  float* uv;             // [B][LR_ZROW]:  sum_m a_m1 delta_m
  float* hba;            // [2B][LR_ZROW]: HB, then HA
  int B, H, Kaug, td;
  float a[6][6];         // a[j - 2][i - 1] = a_ji
  float bt[7];
};
constexpr int kRegRevS = 4;   // samples per group of 128 threads
// block = 256 threads = 2 groups; thread k of a group owns hidden row k of 4 samples
__global__ void __launch_bounds__(256) reg_rev_chain_kernel(RegRevP p) {
  extern __shared__ __align__(16) float rr_smem[];
  float* Ms = rr_smem;                       // [H][128]
  float4* Es = reinterpret_cast<float4*>(rr_smem + p.H * 128);   // [2][128]
  const int tid = threadIdx.x, k = tid & 127, g = tid >> 7;
  for (int i = tid; i < p.H * 128; i += 256) Ms[i] = p.Mz[i];
  const float dt = p.R->c.dt;
  const int H = p.H, Kaug = p.Kaug, ext = p.td + 1;
  const int ngroups = (p.B + kRegRevS - 1) / kRegRevS;
  const int npairs = (ngroups + 1) / 2;
  __syncthreads();
  for (int pr = blockIdx.x; pr < npairs; pr += gridDim.x) {
    const int s0 = (2 * pr + g) * kRegRevS;
    float A[kRegRevS], Bq[kRegRevS], h[7][kRegRevS], dl[6][kRegRevS];
    bool valid[kRegRevS];
#pragma unroll
    for (int s = 0; s < kRegRevS; ++s) {
      const int b = s0 + s;
      valid[s] = b < p.B;
      const bool hrow = valid[s] && k < H;
      A[s] = hrow ? p.ab[(size_t)b * LR_ZROW + k] : 0.0f;
      Bq[s] = hrow ? p.ab[(size_t)(p.B + b) * LR_ZROW + k] : 0.0f;
#pragma unroll
      for (int j = 0; j < 7; ++j) h[j][s] = (valid[s] && k < Kaug) ? __ldcg(p.hh[j] + (size_t)b * LR_ZROW + k) : 0.0f;
    }
    // combinations of the hidden images: c_j, HB, HA (rows < Kaug), then the [t_j ; 1] rows of the c arrays
#pragma unroll
    for (int j = 2; j <= 7; ++j) {
#pragma unroll
      for (int s = 0; s < kRegRevS; ++s) {
        if (!valid[s]) continue;
        float v = 0.0f;
        if (k < Kaug) {
#pragma unroll
          for (int i = 1; i < j; ++i) v = fmaf(p.a[j - 2][i - 1], h[i - 1][s], v);
          v *= dt;
        } else if (k < Kaug + ext) {
          v = __ldcg(p.hh[j - 1] + (size_t)(s0 + s) * LR_ZROW + (k - ext));   // rows H .. Kaug-1 of hh_j = (t_j,) 1
        }
        p.cc[((size_t)(j - 2) * p.B + s0 + s) * LR_ZROW + k] = v;
      }
    }
#pragma unroll
    for (int s = 0; s < kRegRevS; ++s) {
      if (!valid[s]) continue;
      float hb = 0.0f, ha = 0.0f;
#pragma unroll
      for (int j = 2; j <= 7; ++j) hb = fmaf(p.bt[j - 1], h[j - 1][s], hb);
#pragma unroll
      for (int j = 2; j <= 6; ++j) ha = fmaf(p.a[5][j - 1], h[j - 1][s], ha);
      p.hba[(size_t)(s0 + s) * LR_ZROW + k] = (k < Kaug) ? dt * hb : 0.0f;
      p.hba[(size_t)(p.B + s0 + s) * LR_ZROW + k] = (k < Kaug) ? dt * ha : 0.0f;
    }
    // the stage cotangents, stage 7 first
#pragma unroll
    for (int s = 0; s < kRegRevS; ++s) dl[5][s] = (k < H) ? (1.0f - h[6][s] * h[6][s]) * (dt * p.bt[6] * A[s]) : 0.0f;
#pragma unroll
    for (int m = 6; m >= 2; --m) {
      float4 E;
      float* Ev = reinterpret_cast<float*>(&E);
#pragma unroll
      for (int s = 0; s < kRegRevS; ++s) {
        float v = 0.0f;
#pragma unroll
        for (int mm = m + 1; mm <= 7; ++mm) v = fmaf(p.a[mm - 2][m - 1], dl[mm - 2][s], v);
        Ev[s] = v;
      }
      Es[g * 128 + k] = E;
      __syncthreads();
      float acc[kRegRevS] = {0.0f, 0.0f, 0.0f, 0.0f};
      if (k < H) {
#pragma unroll 4
        for (int r = 0; r < H; ++r) {
          const float w = Ms[r * 128 + k];
          const float4 e = Es[g * 128 + r];
          acc[0] = fmaf(w, e.x, acc[0]); acc[1] = fmaf(w, e.y, acc[1]);
          acc[2] = fmaf(w, e.z, acc[2]); acc[3] = fmaf(w, e.w, acc[3]);
        }
      }
      __syncthreads();
#pragma unroll
      for (int s = 0; s < kRegRevS; ++s) {
        const float alpha = dt * (fmaf(p.bt[m - 1], A[s], fmaf(p.a[5][m - 1], Bq[s], acc[s])));
        dl[m - 2][s] = (k < H) ? (1.0f - h[m - 1][s] * h[m - 1][s]) * alpha : 0.0f;
      }
    }
#pragma unroll
    for (int s = 0; s < kRegRevS; ++s) {
      if (!valid[s]) continue;
      float ds = 0.0f, uvv = 0.0f;
#pragma unroll
      for (int j = 2; j <= 7; ++j) {
        p.del[((size_t)(j - 2) * p.B + s0 + s) * LR_ZROW + k] = dl[j - 2][s];
        ds += dl[j - 2][s];
        uvv = fmaf(p.a[j - 2][0], dl[j - 2][s], uvv);
      }
      p.dsum[(size_t)(s0 + s) * LR_ZROW + k] = ds;
      p.uv[(size_t)(s0 + s) * LR_ZROW + k] = uvv;
    }
  }
}

// d_ps += the four contractions (fixed-order sums of their batch splits) and the two small products.
//   t1 [S1][D x KP]  = [G Q] [HB HA]^T        (column-major, ld D)
//   t2 [S2][H x D]   = dsum uprev^T           (ld H)
//   ss [S3][H x KP]  = sum_j delta_j [c_j ; (t_j) ; 1]^T   (ld H; columns Kaug .. Kaug+td: the time / bias columns of dW1)
//   uh [S4][H x KP]  = UV hh_1^T              (ld H)
struct RegAsmP {
  const SolveDev* R;
  const float* ps; float* dps;
  const float *t1, *t2, *ss, *uh;
  float* sm;             // [2][H * Kaug]: S1 = SS - dt UH, then SS (row-major h * Kaug + k)
  int S1, S2, S3, S4;
  int D, H, Kaug, td, KP;
  long w1_off, w2_off;   // layer blocks [H x (D + td + 1)], [D x Kaug] in the flat parameter vector
};
__global__ void __launch_bounds__(256) reg_ss_kernel(RegAsmP p) {
  const int H = p.H, Kaug = p.Kaug;
  const float dt = p.R->c.dt;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= H * Kaug) return;
  const int h = e / Kaug, k = e % Kaug;
  float ssv = 0.0f, uhv = 0.0f;
  for (int z = 0; z < p.S3; ++z) ssv += p.ss[(size_t)z * H * p.KP + h + (size_t)k * H];
  for (int z = 0; z < p.S4; ++z) uhv += p.uh[(size_t)z * H * p.KP + h + (size_t)k * H];
  p.sm[e] = fmaf(-dt, uhv, ssv);
  p.sm[H * Kaug + e] = ssv;
}
__global__ void __launch_bounds__(256) reg_assemble_kernel(RegAsmP p) {
  const int H = p.H, D = p.D, Kaug = p.Kaug;
  const float* S1m = p.sm;
  const float* SSm = p.sm + H * Kaug;
  const float* W1 = p.ps + p.w1_off;   // [H x D] column-major (then the time column and the bias)
  const float* W2a = p.ps + p.w2_off;  // [D x Kaug]
  const size_t n2 = (size_t)D * Kaug, n1 = (size_t)H * D, nx = (size_t)H * (p.td + 1);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n2 + n1 + nx; i += (size_t)gridDim.x * blockDim.x) {
    if (i < n2) {            // dW2a[d, k]
      const int d = (int)(i % D), k = (int)(i / D);
      float v = 0.0f;
      for (int z = 0; z < p.S1; ++z) v += p.t1[(size_t)z * D * p.KP + i];
      float w = 0.0f;
      for (int h = 0; h < H; ++h) w = fmaf(W1[(size_t)d * H + h], S1m[h * Kaug + k], w);
      p.dps[p.w2_off + i] += v + w;
    } else if (i < n2 + n1) {   // dW1[h, d]
      const size_t j = i - n2;
      const int h = (int)(j % H), d = (int)(j / H);
      float v = 0.0f;
      for (int z = 0; z < p.S2; ++z) v += p.t2[(size_t)z * n1 + j];
      float w = 0.0f;
      for (int k = 0; k < Kaug; ++k) w = fmaf(SSm[h * Kaug + k], W2a[(size_t)k * D + d], w);
      p.dps[p.w1_off + j] += v + w;
    } else {                 // time column (TDChain) and bias of layer 1
      const size_t j = i - n2 - n1;
      const int h = (int)(j % H), col = (int)(j / H);
      float v = 0.0f;
      for (int z = 0; z < p.S3; ++z) v += p.ss[(size_t)z * H * p.KP + h + (size_t)(Kaug + col) * H];
      p.dps[p.w1_off + n1 + j] += v;
    }
  }
}
