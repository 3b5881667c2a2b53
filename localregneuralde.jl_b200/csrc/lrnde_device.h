// Device-resident descriptors of the solve.  Everything a kernel needs that depends on the
// step size (stage coefficients dt*a_ij, stage times, which tape slot is current) lives in
// device memory and is rewritten by the on-device controller, so a captured CUDA graph of one
// step attempt can be replayed (or looped by a WHILE node) with no host round trip.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "lrnde_controller.h"

#define LR_MAXSRC 7
#define LR_ERR_BLOCKS 296  // 2 x 148 SMs: partial sums of every norm, summed in fixed order
#define LR_MAX_RANKS 8

// x[i] = base[i] + scale * sum_{k<n} coef[k] * src[k][i]; the f-evaluation that consumes x
// runs at time `t` and writes its result to `dst`.
struct LinComb {
  const float* base;
  const float* src[LR_MAXSRC];
  float* dst;
  float coef[LR_MAXSRC];
  float scale;
  float t;
  int n;
  int pad_;
};

// Mailbox of the data-parallel group (one per rank, peer-mapped on every other rank).
// slot[seq & 1][r][0..3]: the four partial sums rank r published for sequence number seq,
// flag[seq & 1][r]: seq + 1 once they are visible.
struct LrMailbox {
  double val[2][LR_MAX_RANKS][4];
  unsigned long long flag[2][LR_MAX_RANKS];
  unsigned long long mflag[2][LR_MAX_RANKS];  // vector exchange (adjoint mu block)
  unsigned long long pad_[16];
};
// The exported mailbox allocation = LrMailbox followed by the vector staging area
// stage[parity][3][LR_MU_MAX] floats: per-rank partial batch sums of the mu block that the
// adjoint's error norm needs as GLOBAL values.
#define LR_MU_MAX (1 << 20)
// ... followed by the BatchNorm area: per-channel pairs of batch sums (sum x, sum x^2 of the forward; sum ghat,
// sum ghat xhat of the pullback) that conv dynamics with BatchNorm need as GLOBAL values (SURVEY 8e (4),
// experiments/src/construct.jl:213-216).  val[seq & 1][r][c], flag[seq & 1][r][c] = seq + 1 once visible; every channel
// counts its own exchanges (lrnde_ctx::bn_seq), so the C blocks of a finalize kernel exchange independently.
#define LR_BN_MAXC 256
struct LrBnBox {
  double val[2][LR_MAX_RANKS][LR_BN_MAXC][4];   // (sum 1, sum 2, element count of the rank, unused)
  unsigned long long flag[2][LR_MAX_RANKS][LR_BN_MAXC];
};
#define LR_MU_STAGE_BYTES (sizeof(float) * 2 * 3 * (size_t)LR_MU_MAX)
#define LR_MAILBOX_BYTES (sizeof(LrMailbox) + LR_MU_STAGE_BYTES + sizeof(LrBnBox))
__host__ __device__ inline float* lr_mbox_stage(LrMailbox* m, int parity, int vec) {
  return reinterpret_cast<float*>(m + 1) + ((size_t)parity * 3 + vec) * LR_MU_MAX;
}
__host__ __device__ inline LrBnBox* lr_mbox_bn(LrMailbox* m) {
  return reinterpret_cast<LrBnBox*>(reinterpret_cast<char*>(m + 1) + LR_MU_STAGE_BYTES);
}
// the data-parallel group as a kernel argument of the BatchNorm finalize kernels (nranks <= 1: no exchange)
struct BnDist {
  LrMailbox* mbox[LR_MAX_RANKS];
  int rank, nranks;
  unsigned long long* seq;   // [LR_BN_MAXC] exchanges done per channel (device memory of this rank)
  int* timed_out;            // set when a peer never arrived (~10 s): the host call fails with LRNDE_ESTATE
};
#ifdef __CUDACC__
// (a, b, count) <- sums over ranks in rank order (identical bits on every rank).  One calling thread per channel.
__device__ inline void lr_bn_group_sum(const BnDist& d, int c, double& a, double& b, double& count) {
  if (d.nranks <= 1) return;
  const unsigned long long seq = d.seq[c];
  d.seq[c] = seq + 1;
  const int par = (int)(seq & 1ull);
  for (int r = 0; r < d.nranks; ++r) {
    volatile double* dst = lr_mbox_bn(d.mbox[r])->val[par][d.rank][c];
    dst[0] = a; dst[1] = b; dst[2] = count;
  }
  __threadfence_system();
  for (int r = 0; r < d.nranks; ++r) {
    volatile unsigned long long* f = &lr_mbox_bn(d.mbox[r])->flag[par][d.rank][c];
    *f = seq + 1;
  }
  LrBnBox* mine = lr_mbox_bn(d.mbox[d.rank]);
  for (int r = 0; r < d.nranks; ++r) {
    volatile unsigned long long* f = &mine->flag[par][r][c];
    long long spins = 0;
    while (*f != seq + 1) {
      __nanosleep(40);
      if (++spins > (1ll << 26)) { if (d.timed_out) *d.timed_out = 1; break; }
    }
  }
  __threadfence_system();
  double sa = 0.0, sb = 0.0, sc = 0.0;
  for (int r = 0; r < d.nranks; ++r) {
    const volatile double* src = mine->val[par][r][c];
    sa += src[0]; sb += src[1]; sc += src[2];
  }
  a = sa; b = sb; count = sc;
}
#endif

struct SolveDev {
  LrCtrl c;
  int done;       // segment finished, solve failed, or tape full: step kernels become no-ops
  int failed;     // retcode != SUCCESS
  int tape_full;  // forward only: host must grow the tape and resume
  int u_nan;
  int slot;  // current slot: U(slot)=uprev, K(slot,1)=fsalfirst
  int cap;   // number of slots
  int ring;  // 1: slot index wraps (adjoint / regulariser integrator), 0: dense tape
  int nlog, logcap;
  int nf;  // f evaluations so far
  int reduce_mu;  // adjoint, multi-rank: the mu block of z is a partial batch sum
  float abstol, reltol;
  unsigned long long cond_handle;
  int use_cond;
  int is_adjoint;
  size_t len;      // elements per array of this solve (D*B, or D*B + P for the adjoint)
  size_t lam_len;  // D*B
  float* tape;     // slot s at tape + s * 7 * len : [u, k1, ..., k6]; k7 == k1 of slot s+1
  float* ts;       // accepted times, ts[0] = t0 (dense tape only)
  float* log_t;
  float* log_dt;
  float* log_eest;
  unsigned char* log_acc;
  double* partials;  // [4][LR_ERR_BLOCKS]
  unsigned int* counters;  // [4] order-independent integer flags (non-finite count, f0!=f1 count)
  LinComb st[7];   // [0..4] inputs of stages 2..6 (dst = K(slot, j)); [5] u_new (dst U(slot+1));
                   // [6] input of stage 7 (base = U(slot+1), dst = K(slot+1, 1))
  LinComb err;     // src = k1..k7, coef = btilde, scale = dt, base = uprev, dst = u_new
  LinComb yint[7]; // adjoint: forward-solution interpolants for the RHS evaluations;
                   // [0] at the current time (fsalfirst refresh), [j] for st[j-1] ... see api
  // dense forward solution the adjoint interpolates
  const float* fts;
  const float* ftape;
  size_t flen;
  int fnsteps;
  int ftdir;
  // initdt scratch
  float d0, d1, dt0;
  float reg_ss_root;  // regulariser integrator: sqrt(ss/n) kept for the reverse pass
  float reg_val;
  float reg_aux[3];
  // data-parallel group
  int rank, nranks;
  unsigned long long seq;  // collective sequence number
  unsigned long long total_len;  // elements of the GLOBAL array (norm denominators)
  LrMailbox* mbox[LR_MAX_RANKS];
  unsigned long long mseq;       // vector-exchange sequence number
  float* muglob;                 // [3][mu_len]: globally summed mu vectors of the current exchange
  unsigned int* mucounter;       // blocks-finished counter of mu_publish_kernel
  size_t mu_len;                 // P
  // latent-space engine (lrnde_fused.cu): Z(x) = W1[:, :D] x for every array x of the tape, same slot layout
  // ([B][LR_ZROW] floats per array)
  float* ztape;
  size_t zlen;
  float* htape;   // forward: hidden activations h of every k (H(k) = act(W1a [g; t; 1])); adjoint: delta_1 of every k
  // adjoint: the latent / hidden tapes of the forward solve it interpolates (same slot layout as ftape)
  const float* fztape;
  const float* fhtape;
  size_t fzlen;
  LinComb cur;      // the current state as a descriptor (base = U(slot), t = c.t): written by k1_desc_kernel
  int lat_mu_row;   // latent-space adjoint: the mu block's residual partial sums sit in partials[LR_ERR_BLOCKS ...]
};
// a rank of the group never published its value: the solve ends with retcode PeerTimeout on this rank (the
// controller sees `failed`, closes the WHILE loop, and the host reports the retcode) instead of spinning forever
#ifdef __CUDACC__
__device__ inline void lr_peer_timeout(SolveDev* S) {
  S->failed = 1;
  S->done = 1;
  if (S->c.retcode == LR_RET_SUCCESS) S->c.retcode = LR_RET_PEERTIMEOUT;
}
#endif

#define LR_ZROW 128
// the latent image of a tape array (p must point at the start of an array of this solve's tape)
__host__ __device__ inline float* lr_zof(const SolveDev* S, const float* p) {
  return S->ztape + ((size_t)(p - S->tape) / S->len) * S->zlen;
}
__host__ __device__ inline float* lr_hof(const SolveDev* S, const float* p) {
  return S->htape + ((size_t)(p - S->tape) / S->len) * S->zlen;
}

__host__ __device__ inline float* lr_slot_u(const SolveDev* S, int s) {
  return S->tape + (size_t)s * 7 * S->len;
}
__host__ __device__ inline float* lr_slot_k(const SolveDev* S, int s, int j) {  // j = 1..6
  return S->tape + ((size_t)s * 7 + j) * S->len;
}
__host__ __device__ inline int lr_next_slot(const SolveDev* S, int s) {
  return S->ring ? ((s + 1) % S->cap) : (s + 1);
}
