// Small multilayer perceptrons evaluated out of shared memory: the building block of every path whose
// networks fit one SM (neural SDE, latent-ODE encoder, small-state neural ODEs such as the physionet
// decoder).  A CTA keeps the whole network in shared memory ([in + td + 1][out_padded] blocks: rows = inputs,
// time row, bias row; float4 along the outputs) and the activations of a tile of S samples as [dim][S + 1]
// (conflict-free for both the forward and the transposed access of the reverse pass).
#pragma once
#include <cuda_runtime.h>

#include "lrnde_kernels.cuh"

#define SDE_MAXL 10
#define SDE_STK 64
#define SDE_THREADS 256

struct SdeNet {
  int nl, td, D;
  int in[SDE_MAXL], out[SDE_MAXL], outp[SDE_MAXL], act[SDE_MAXL];
  int w_off[SDE_MAXL];    // float offset of the layer's [(in + td + 1) x outp] block in shared memory
  int hid_off[SDE_MAXL];  // row offset of the layer's pre / post activation buffers
  long long ps_w[SDE_MAXL], ps_b[SDE_MAXL];
  int wfloats;            // shared-memory floats of all blocks
  int hid_rows;           // total rows of the activation buffers
  int maxdim;
  int nparams;
};

struct SdeConsts {
  float beta1, beta2, gamma, qmin, qmax, qoldinit, delta, discard, order;
  int pow_mode;
};

struct SdeTile { int S, SP, D, nvalid, nthr, tid; };

// ---------------------------------------------------------------- networks in shared memory
__device__ void sde_load_weights(const SdeNet& n, const float* __restrict__ ps, float* W, int tid, int nthr) {
  for (int l = 0; l < n.nl; ++l) {
    const int rows = n.in[l] + n.td + 1, outp = n.outp[l], out = n.out[l];
    for (int e = tid; e < rows * outp; e += nthr) {
      const int i = e / outp, r = e % outp;
      float v = 0.0f;
      if (r < out) v = (i < rows - 1) ? ps[n.ps_w[l] + (long long)i * out + r] : ps[n.ps_b[l] + r];
      W[n.w_off[l] + e] = v;
    }
  }
}

// y = net(x, t) for the S samples of the tile; x, y: [D][SP]; pre/post: [hid_rows][SP]
__device__ void sde_mlp_fwd(const SdeNet& n, const float* W, const float* x, float t, float* y, float* pre,
                            float* post, const SdeTile& T) {
  const float* cur = x;
  for (int l = 0; l < n.nl; ++l) {
    const int in = n.in[l], out = n.out[l], outp = n.outp[l], act = n.act[l];
    const float* Wl = W + n.w_off[l];
    float* pl = pre + n.hid_off[l] * T.SP;
    float* dst = (l == n.nl - 1) ? y : post + n.hid_off[l] * T.SP;
    for (int item = T.tid; item < (outp >> 2) * T.S; item += T.nthr) {
      const int s = item % T.S, r0 = (item / T.S) << 2;
      float4 acc = *reinterpret_cast<const float4*>(Wl + (in + n.td) * outp + r0);
      if (n.td) {
        const float4 w = *reinterpret_cast<const float4*>(Wl + in * outp + r0);
        acc.x = fmaf(w.x, t, acc.x); acc.y = fmaf(w.y, t, acc.y); acc.z = fmaf(w.z, t, acc.z); acc.w = fmaf(w.w, t, acc.w);
      }
#pragma unroll 8
      for (int i = 0; i < in; ++i) {
        const float xv = cur[i * T.SP + s];
        const float4 w = *reinterpret_cast<const float4*>(Wl + i * outp + r0);
        acc.x = fmaf(w.x, xv, acc.x); acc.y = fmaf(w.y, xv, acc.y); acc.z = fmaf(w.z, xv, acc.z); acc.w = fmaf(w.w, xv, acc.w);
      }
      const float a[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (r0 + j < out) {
          pl[(r0 + j) * T.SP + s] = a[j];
          dst[(r0 + j) * T.SP + s] = lr_act(act, a[j]);
        }
    }
    __syncthreads();
    cur = dst;
  }
}

// (J_x^T cot, J_p^T cot) of net at (x, t): xbar += J_x^T cot (when xbar != nullptr), G += J_p^T cot.
// cot: [D][SP] (read only); dA/dB: scratch [maxdim][SP].  Samples >= nvalid contribute nothing.
__device__ void sde_mlp_vjp(const SdeNet& n, const float* W, float* G, const float* x, float t, const float* cot,
                            float* xbar, float* y_scratch, float* pre, float* post, float* dA, float* dB,
                            const SdeTile& T) {
  sde_mlp_fwd(n, W, x, t, y_scratch, pre, post, T);
  const int L = n.nl - 1;
  {
    const float* pl = pre + n.hid_off[L] * T.SP;
    for (int idx = T.tid; idx < n.out[L] * T.S; idx += T.nthr) {
      const int r = idx / T.S, s = idx % T.S, a = r * T.SP + s;
      dA[a] = (s < T.nvalid) ? cot[a] * lr_dact(n.act[L], pl[a]) : 0.0f;
    }
  }
  __syncthreads();
  float* dcur = dA;
  float* dnext = dB;
  for (int l = L; l >= 0; --l) {
    const int in = n.in[l], out = n.out[l], outp = n.outp[l];
    const float* Wl = W + n.w_off[l];
    float* Gl = G + n.w_off[l];
    const float* xin = (l == 0) ? x : post + n.hid_off[l - 1] * T.SP;
    const int rows = in + n.td + 1;
    for (int item = T.tid; item < rows * outp; item += T.nthr) {
      const int i = item / outp, r = item % outp;
      if (r >= out) continue;
      float acc = 0.0f;
      if (i < in) {
#pragma unroll 8
        for (int s = 0; s < T.S; ++s) acc = fmaf(xin[i * T.SP + s], dcur[r * T.SP + s], acc);
      } else {
        for (int s = 0; s < T.S; ++s) acc += dcur[r * T.SP + s];
        if (n.td && i == in) acc *= t;
      }
      Gl[item] += acc;
    }
    if (l > 0 || xbar) {
      const float* pprev = (l > 0) ? pre + n.hid_off[l - 1] * T.SP : nullptr;
      const int actprev = (l > 0) ? n.act[l - 1] : 0;
      const int nb = (in + 3) >> 2;
      for (int item = T.tid; item < nb * T.S; item += T.nthr) {
        const int s = item % T.S, i0 = (item / T.S) << 2;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
        for (int r = 0; r < out; ++r) {
          const float d = dcur[r * T.SP + s];
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (i0 + j < in) acc[j] = fmaf(Wl[(i0 + j) * outp + r], d, acc[j]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (i0 + j < in) {
            const int a = (i0 + j) * T.SP + s;
            if (l > 0) dnext[a] = acc[j] * lr_dact(actprev, pprev[a]);
            else xbar[a] += acc[j];
          }
      }
    }
    __syncthreads();
    float* tmp = dcur; dcur = dnext; dnext = tmp;
  }
}


// device limits, queried once per device (cudaGetDeviceProperties is a slow, lock-taking call)
static void sde_device_limits(int device, int* n_sm, size_t* smem_optin) {
  static int cached_dev = -1, c_sm = 0, c_smem = 0;
  if (cached_dev != device) {
    LR_CUDA(cudaDeviceGetAttribute(&c_sm, cudaDevAttrMultiProcessorCount, device));
    LR_CUDA(cudaDeviceGetAttribute(&c_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    cached_dev = device;
  }
  *n_sm = c_sm;
  *smem_optin = (size_t)c_smem;
}


// ==========================================================================================
// Small-state engine of the neural-ODE path (precision LRNDE_PREC_SMEM): when the dynamics network fits
// shared memory, one stage evaluation f(lincomb(in), ps, t) is ONE launch (instead of one tcgen05 GEMM per
// layer) and one adjoint right-hand side (J_u^T lam, J_p^T lam) is one launch plus a fixed-order
// reduction (instead of ~4 launches per layer).  Same LinComb descriptor contract as MlpEval::forward / vjp.
// ==========================================================================================
struct SmallP {
  SdeNet n;
  int in_act;
  const float* ps;
  const LinComb* in;      // stage combination (forward) / y(t) interpolant (vjp)
  const LinComb* out;     // forward: output descriptor (nullptr: in->dst)
  int side_to_in_dst;     // forward: also store the combined input to in->dst
  const LinComb* lam;     // vjp: cotangent combination
  float* out_a; const LinComb* out_desc; float a_scale;   // vjp: a_scale * J_u^T lam
  float* gpart;           // vjp: [grid][wfloats] partial parameter gradients
  const int* done;
  int B, D, S, SP, tiles_per_cta;
};

static size_t small_smem_bytes(const SdeNet& n, int D, int SP, bool bwd) {
  size_t f = (size_t)n.wfloats * (bwd ? 2 : 1);
  int rows = 2 * D + 2 * n.hid_rows + (bwd ? 4 * D + 2 * n.maxdim : D);
  return (f + (size_t)rows * SP) * sizeof(float) + 16;
}

template <bool BWD>
__global__ void __launch_bounds__(SDE_THREADS, 1) small_kernel(SmallP p) {
  if (p.done && *p.done) return;
  extern __shared__ __align__(16) float sde_sm[];
  __shared__ LinComb d, d2;
  const int tid = threadIdx.x, nthr = blockDim.x, D = p.D, SP = p.SP;
  if (tid == 0) {
    d = *p.in;
    if (BWD) d2 = *p.lam;
    else if (p.out) d2 = *p.out;
  }
  float* q = sde_sm;
  float* W = q; q += p.n.wfloats;
  float* G = nullptr;
  if (BWD) { G = q; q += p.n.wfloats; }
  auto take = [&](int rows) { float* r = q; q += rows * SP; return r; };
  float* xraw = take(D); float* xin = take(D);
  float* pre = take(p.n.hid_rows); float* post = take(p.n.hid_rows);
  float *yout = nullptr, *cot = nullptr, *xbar = nullptr, *dA = nullptr, *dB = nullptr;
  if (BWD) { cot = take(D); xbar = take(D); yout = take(D); dA = take(p.n.maxdim); dB = take(p.n.maxdim); (void)take(D); }
  else yout = take(D);
  sde_load_weights(p.n, p.ps, W, tid, nthr);
  if (BWD) for (int e = tid; e < p.n.wfloats; e += nthr) G[e] = 0.0f;
  __syncthreads();
  const float t = d.t;
  float* outp = BWD ? (p.out_a ? p.out_a : p.out_desc->dst) : (p.out ? d2.dst : d.dst);
  float* side = (!BWD && p.side_to_in_dst) ? d.dst : nullptr;
  SdeTile T{p.S, SP, D, 0, nthr, tid};
  for (int ti = 0; ti < p.tiles_per_cta; ++ti) {
    const int b0 = (blockIdx.x * p.tiles_per_cta + ti) * p.S;
    if (b0 >= p.B) break;
    T.nvalid = min(p.S, p.B - b0);
    for (int idx = tid; idx < D * T.S; idx += nthr) {
      const int s = idx / D, r = idx % D, a = r * SP + s;
      float x = 0.0f, c = 0.0f;
      if (s < T.nvalid) {
        const size_t g = (size_t)(b0 + s) * D + r;
        x = lr_lincomb_at(d, g);
        if (side) side[g] = x;
        if (BWD) c = lr_lincomb_at(d2, g);
      }
      xraw[a] = x;
      xin[a] = p.in_act ? lr_act(p.in_act, x) : x;
      if (BWD) { cot[a] = c; xbar[a] = 0.0f; }
    }
    __syncthreads();
    if (!BWD) {
      sde_mlp_fwd(p.n, W, xin, t, yout, pre, post, T);
      for (int idx = tid; idx < D * T.S; idx += nthr) {
        const int s = idx / D, r = idx % D;
        if (s < T.nvalid) outp[(size_t)(b0 + s) * D + r] = yout[r * SP + s];
      }
    } else {
      sde_mlp_vjp(p.n, W, G, xin, t, cot, xbar, yout, pre, post, dA, dB, T);
      for (int idx = tid; idx < D * T.S; idx += nthr) {
        const int s = idx / D, r = idx % D, a = r * SP + s;
        if (s < T.nvalid) {
          float v = xbar[a];
          if (p.in_act) v *= lr_dact(p.in_act, xraw[a]);
          outp[(size_t)(b0 + s) * D + r] = p.a_scale * v;
        }
      }
    }
    __syncthreads();
  }
  if (BWD) {
    float* mine = p.gpart + (size_t)blockIdx.x * p.n.wfloats;
    for (int e = tid; e < p.n.wfloats; e += nthr) mine[e] = G[e];
  }
}

// dst = beta * dst + scale * sum_cta gpart, un-padded into the flat parameter layout.  One thread per padded
// element, the per-CTA partials added in CTA order (deterministic) with the loads issued 8 at a time.
__global__ void small_dps_reduce_kernel(SdeNet n, const float* gpart, int nblk, float* dst, const LinComb* dstdesc,
                                        size_t dst_off, float scale, float beta, const int* done) {
  if (done && *done) return;
  float* out = dstdesc ? (dstdesc->dst + dst_off) : dst;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n.wfloats; e += gridDim.x * blockDim.x) {
    int l = 0;
    while (l + 1 < n.nl && e >= n.w_off[l + 1]) ++l;
    const int loc = e - n.w_off[l];
    const int rows = n.in[l] + n.td + 1, outp = n.outp[l], o = n.out[l];
    const int i = loc / outp, r = loc % outp;
    if (r >= o) continue;
    float acc = 0.0f;
#pragma unroll 8
    for (int bb = 0; bb < nblk; ++bb) acc += gpart[(size_t)bb * n.wfloats + e];
    const long long j = (i < rows - 1) ? n.ps_w[l] + (long long)i * o + r : n.ps_b[l] + r;
    float v = scale * acc;
    if (beta != 0.0f) v = fmaf(beta, out[j], v);
    out[j] = v;
  }
}
