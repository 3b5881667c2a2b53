// Latent-space engine of the neural-ODE path for TDChain(Dense(D(+1) => H, act), Dense(H(+1) => D)) dynamics
// (the mnist_ode family, experiments/src/construct.jl:184-199).
//
// The first layer is linear in its input, so the pre-activation of every Tsit5 stage
//     W1 [g_j ; t_j ; 1],   g_j = uprev + dt * sum_i a_ji k_i           (src/perform_step.jl:11-19)
// is  Z(uprev) + dt * sum_i a_ji Z(k_i) + w1t t_j + b1  with  Z(x) = W1[:, :D] x,  and
//     Z(k_j) = Z(W2a [h_j ; t_j ; 1]) = (W1[:, :D] W2a) [h_j ; t_j ; 1] =: Mz [h_j ; t_j ; 1]   (an H x (H+2) matrix).
// A step attempt is therefore
//   (1) chain_kernel : the six stages as a recurrence in the H-dimensional latent space (tcgen05 Mz GEMMs,
//       Z's in TMEM), producing the hidden activations h_2..h_7 as ready-made tf32 hi/lo operand images;
//   (2) kgemm_kernel : k_2..k_7 = W2a [h_j ; t_j ; 1] for all six stages in one tcgen05 GEMM whose epilogue
//       writes the k's straight to the dense tape, forms u_{n+1} = uprev + dt * sum a_7i k_i and the Hairer
//       residual partial sums (perform_step.jl:18-27,34-38) -- 2 arrays read, 7 written: the 9 * 4 * D * B
//       byte minimum of SURVEY 8(d);
//   (3) controller_kernel (unchanged).
// Every k_j is still materialised, so the tape, the interpolant, the saves and the adjoint are unchanged.
#pragma once
#include <stdlib.h>

#include "lrnde_host.h"

struct FusedShape {
  int D, H, td, act;
  int Kaug;    // H + td + 1: hidden units, time row, bias row
  int KS;      // K-steps of 8 (tf32) covering Kaug
  int nfull;   // full 32-float K-chunks (SWIZZLE_128B tiles)
  int ntail;   // remaining K-steps (SWIZZLE_32B tiles of 8 floats)
  int MT;      // output features per kgemm tile (multiple of 16, <= 128)
  int n_mt;    // number of feature tiles
};

// shape of an eligible model (false: the model is not a 2-layer TD-MLP the latent-space engines handle)
bool lrf_shape(const lrnde_model* m, FusedShape* out);
// lambda GEMM of the latent-space adjoint (lrnde_adjoint.h): kgemm_kernel<2, true> over units of 16 samples
void lrf_launch_kgemm_adj(lrnde_ctx* ctx, SolveDev* S, const FusedShape& sh, const float* W1, const float* hbuf,
                          size_t unit_bytes, int64_t B, int passes, int nunits, int nclusters);

// Mz = W1[:, :D] W2a as a plain zero-padded [128][128] matrix (device), accumulated in double
void lrf_mz_plain(lrnde_ctx* ctx, const lrnde_model* m, const float* ps, float* out);

struct FusedEngine {
  lrnde_ctx* ctx;
  const lrnde_model* m;
  const float* ps;
  int64_t B;
  int passes;
  FusedShape sh;
  int ntiles = 0;        // 64-sample tiles of the chain kernel
  int nunits = 0;        // 16-sample units of the kgemm kernel
  int nbuf = 2;          // chain kernel: operand tile buffers
  int cluster = 1;       // kgemm: CTAs per cluster (= n_mt when the multicast path is usable)
  int nclusters = 1;     // kgemm: persistent clusters
  int ring = 8;          // kgemm: operand pieces in flight
  float* Mimg = nullptr;   // chain A operand: [hi | lo] images of Mz, 128 rows
  float* Mplain = nullptr; // Mz as a plain zero-padded [128][128] matrix
  float* hbuf = nullptr;   // operand images written by the chain kernel: per 16-sample unit, 96 rows = 6 stages x 16
  size_t imgM = 0, unit_bytes = 0, unit_bytes2 = 0;
  int lean = 0;                     // lean tape: attempts store u_{n+1} only (the hidden tape carries everything else)
  // lean attempts run the two-column GEMM (sum a_7i k_i, sum btilde_i k_i) when both operand buffers exist
  bool combo() const { return lean && nbuf == 2 && !getenv("LRNDE_NO_COMBO"); }
  const float* add_base = nullptr;  // dense_output(): base array of the single-mode kgemm

  static bool eligible(const lrnde_model* m);
  FusedEngine(lrnde_ctx* c, const lrnde_model* mm, const float* p, int64_t b, int npasses);
  ~FusedEngine();
  size_t zlen() const { return (size_t)LR_ZROW * (size_t)B; }
  // Mz = W1 W2a (double accumulation) and its operand image; mz_plain_out (optional): where the plain [128][128] copy goes
  void prepare(float* mz_plain_out = nullptr);
  // one Tsit5 attempt from the descriptors of S (st[0..6], err): launches (1) and (2)
  void step(SolveDev* S, int write_z);
  void step_chain(SolveDev* S, int write_z);   // the two launches of step(), separately (profiling)
  void step_kgemm(SolveDev* S);
  // dst <- f(lincomb(in)) for a descriptor whose arrays live on S's tape; (out ? out : in)->dst receives k
  void eval(SolveDev* S, const LinComb* in, const LinComb* out, const int* done, int write_z);
  // dense output from the hidden tape: out_desc->dst <- xbase + W2a (harr[0] + scale * sum_i coef[i] * harr[i + 1])
  // (harr: C_n, H(k_1..7) of the interval, [B][LR_ZROW] each; out_desc: device descriptor, only dst is used)
  void dense_output(SolveDev* S, const float* const* harr, const float* coef, float scale, const float* xbase,
                    const LinComb* out_desc);
  // the same from a device-resident interpolant descriptor yd of the adjoint solve A (A->fhtape ...): out_desc->dst (=
  // ybuf) <- y(yd->t), and *ydesc_out <- the materialised state as a descriptor {base = ybuf, t = yd->t}
  void dense_output_dev(SolveDev* A, const LinComb* yd, const float* xbase, float* ybuf, LinComb* ydesc_out,
                        const LinComb* out_desc);
};
