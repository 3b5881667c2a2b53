/*
 * libLRNDE.so -- C ABI of the B200-native LocalRegNeuralDE hot path.
 *
 * This is the boundary a Julia maintainer binds with `ccall` from the unchanged
 * layer constructors.  Plain C: pointers, sizes, PODs.  No
 * torch / CUDA types in any signature (a CUDA stream is passed as `void*`).
 *
 * Citations are into the reference tree (avik-pal/LocalRegNeuralDE.jl):
 *   lrnde_ode_forward   replaces the body of (n::NeuralODE)(x, ps, st)
 *                       src/layers/neural_ode.jl:62-100, i.e. solve (:42-54),
 *                       _get_ode_integrator (:33-38) and _perform_step
 *                       (src/perform_step.jl:3-47).
 *   lrnde_ode_backward  replaces the Zygote pullback through that functor:
 *                       SciMLSensitivity InterpolatingAdjoint(ZygoteVJP) selected at
 *                       neural_ode.jl:11 / experiments/src/construct.jl:197, plus the
 *                       reverse pass of _perform_step w.r.t. ps only (neural_ode.jl:40,
 *                       src/utils.jl:60; pinned by test/runtests.jl:127-131).
 *   lrnde_model_create  describes the dynamics net the reference passes as `model`:
 *                       Lux Chain of Dense layers, optionally a TDChain
 *                       (src/layers/common.jl:2-45) -- time row appended before every
 *                       layer -- as built in experiments/src/construct.jl:180-189,235-243.
 *   lrnde_sosri_step    replaces _perform_step(::FourStageSRIConstantCache)
 *                       src/perform_step.jl:49-106 for diagonal noise.
 *
 * Layout everywhere is the reference's: column-major [features, batch] Float32, `ps`
 * in ComponentArray order (layer_1.weight[out x in(+1)], layer_1.bias[out], ...).
 *
 * Return value: 0 on success, negative LRNDE_E* otherwise; lrnde_last_error() gives a
 * thread-local message.  Solver outcomes (MaxIters, DtLessThanMin, Unstable) are NOT
 * errors: like the reference they are reported in stats.retcode and the call returns 0.
 *
 * Threading: one ctx <-> one device + one stream; a ctx is not thread-safe, different
 * ctxs are independent.  The library never calls back into the host language.
 */
#ifndef LRNDE_H
#define LRNDE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LRNDE_OK 0
#define LRNDE_EINVAL (-1)   /* bad argument; mirrors ArgumentError of src/utils.jl:53-58 */
#define LRNDE_ECUDA (-2)    /* CUDA runtime failure */
#define LRNDE_ENOMEM (-3)   /* dense tape does not fit the configured budget */
#define LRNDE_ESTATE (-4)   /* tape/ctx used out of order */

/* activations (Lux/NNlib names) */
enum { LRNDE_ACT_IDENTITY = 0, LRNDE_ACT_TANH = 1, LRNDE_ACT_GELU = 2, LRNDE_ACT_SIGMOID = 3,
       LRNDE_ACT_RELU = 4, LRNDE_ACT_NONE = -1 };
/* regularize (src/layers/neural_ode.jl:14-18) */
enum { LRNDE_REG_NONE = 0, LRNDE_REG_UNBIASED = 1, LRNDE_REG_BIASED = 2 };
/* regularize_type (src/perform_step.jl:34,40) */
enum { LRNDE_REGTYPE_ERROR = 0, LRNDE_REGTYPE_STIFFNESS = 1 };
/* arithmetic of the dense contractions */
enum { LRNDE_PREC_AUTO = 0, LRNDE_PREC_FP32_SIMT = 1, LRNDE_PREC_TF32X3 = 2, LRNDE_PREC_TF32 = 3,
       LRNDE_PREC_SMEM = 4 /* FP32, whole network out of shared memory: one launch per evaluation (small nets) */ };
/* solver retcodes (OrdinaryDiffEq ReturnCode subset the loop can produce) */
enum { LRNDE_RET_SUCCESS = 0, LRNDE_RET_MAXITERS = 1, LRNDE_RET_DTMIN = 2, LRNDE_RET_UNSTABLE = 3,
       LRNDE_RET_TAPEFULL = 4,
       LRNDE_RET_PEERTIMEOUT = 5 /* data-parallel group: a rank did not reach a norm / mu exchange within ~10 s */ };
/* controller power function */
enum { LRNDE_POW_FASTPOW2023 = 0, LRNDE_POW_EXACT = 1 };

typedef struct lrnde_ctx lrnde_ctx;
typedef struct lrnde_model lrnde_model;
typedef struct lrnde_tape lrnde_tape;

/* One Lux Dense(in => out, act).  in_dims EXCLUDES the TDChain time row. */
typedef struct {
  int32_t in_dims;
  int32_t out_dims;
  int32_t act; /* LRNDE_ACT_* */
} lrnde_layer_desc;

/* One layer of the conv dynamics (experiments/src/construct.jl:212-218):
 * Conv((3,3), in_ch(+1) => out_ch; pad=(1,1), use_bias=false), optionally wrapped as
 * Chain(Conv, BatchNorm(out_ch, act)).  in_ch EXCLUDES the TDChain time channel.  Parameters in
 * ComponentArray order: weight[3,3,in_ch(+1),out_ch], then (batchnorm) scale[out_ch], bias[out_ch]. */
typedef struct {
  int32_t in_ch;
  int32_t out_ch;
  int32_t batchnorm; /* 0 / 1; training-mode batch statistics over (W,H,B), eps = 1e-5 */
  int32_t act;       /* LRNDE_ACT_*: BatchNorm's activation (or the conv's own when batchnorm == 0) */
} lrnde_conv_layer_desc;

/* kwargs of NeuralODE(...) / solve(...) that reach the hot path. */
typedef struct {
  float t0, t2;          /* tspan (neural_ode.jl:12) */
  float abstol, reltol;  /* splatted kwargs (neural_ode.jl:51) */
  int32_t maxiters;      /* neural_ode.jl:13 */
  int32_t reg_mode;      /* LRNDE_REG_*; eval mode => pass LRNDE_REG_NONE (neural_ode.jl:66) */
  int32_t reg_type;      /* LRNDE_REGTYPE_* */
  float t1;              /* unbiased: host-sampled t1 (neural_ode.jl:71) */
  float u01;             /* biased: host-sampled uniform [0,1) picking a step time (:92) */
  const float* saveat;   /* HOST pointer, nsave sorted times in [t0,t2]; NULL => see nsave */
  int32_t nsave;         /* >0: save at saveat[]; 0: save only t2; -1: every accepted step */
  int32_t save_start;    /* with nsave==-1: include t0 */
  int32_t precision;     /* LRNDE_PREC_* */
  int32_t pow_mode;      /* LRNDE_POW_* */
  int32_t host_buffers;  /* 1: ps/x/u_save/d_* are HOST pointers (Julia Array); staged inside */
  int32_t keep_tape;     /* 1: return a tape for lrnde_ode_backward (training) */
  int32_t loop_mode;     /* 0: CUDA-graph WHILE node; 1: host-chunked replay (diagnostic) */
  int32_t last_only;     /* 1: return only sol.u[end] (the diffeqsol_to_array layer that follows the functor in
                            every classifier of experiments/src/construct.jl, src/utils.jl:37, fused): u_save and
                            d_u_save then hold ONE [D,B] block; t1 / the regulariser are unaffected */
  float* model_state;    /* conv dynamics with BatchNorm: st.model of the layer (neural_ode.jl:44-47, :83), per
                            BatchNorm layer running_mean[C] then running_var[C]; HOST or device pointer as
                            host_buffers says; NULL = batch statistics, nothing tracked.  Training mode: every f
                            evaluation of the main solve updates it through the closure exactly as Lux does
                            (momentum 0.1, unbiased variance) and the buffer holds st'.model on return */
  int32_t model_testmode; /* 1: Lux.testmode(st.model): BatchNorm normalises with model_state, no update */
  int32_t no_dx;          /* 1: lrnde_ode_backward does not return dL/dx (d_x may be NULL): nothing upstream of the
                             layer has parameters, as in every classifier of experiments/src/construct.jl */
  int32_t reserved[2];
} lrnde_opts;

typedef struct {
  int32_t nfe;        /* st'.nfe: sol.destats.nf (+ 6 + 3 when regularised), neural_ode.jl:79 */
  int32_t naccept, nreject;
  int32_t retcode;    /* LRNDE_RET_* of the main solve */
  float reg_val;      /* st'.reg_val */
  float t1_used;      /* the sampled time actually used (biased: a step time) */
  float dt_reg;       /* auto-selected initial dt of the regulariser integrator */
  int32_t nsave_out;  /* number of [D,B] blocks written to u_save */
  int32_t nf_bwd, naccept_bwd, nreject_bwd, retcode_bwd; /* filled by lrnde_ode_backward */
  int32_t gpu_launches; /* kernels this call launched (graph nodes x iterations included) */
  int32_t reserved[7];
} lrnde_stats;

const char* lrnde_last_error(void);
int lrnde_version(void);

/* stream: a cudaStream_t the caller's buffers are ordered on, or NULL: the library then creates a BLOCKING stream,
 * i.e. one that synchronises implicitly with the legacy default stream (stream 0): work the caller queued on stream 0
 * before a call is finished before the library reads its buffers, and later stream-0 work waits for the call. */
int lrnde_ctx_create(lrnde_ctx** out, int device, void* stream);
int lrnde_ctx_destroy(lrnde_ctx* ctx);
/* Blocks until everything this ctx enqueued has finished. */
int lrnde_ctx_sync(lrnde_ctx* ctx);
/* Upper bound (bytes) for the dense-tape arena; default 0 = 60% of free HBM at first use. */
int lrnde_ctx_set_tape_budget(lrnde_ctx* ctx, uint64_t bytes);
/* Data-parallel group: this ctx is `rank` of `nranks`; the per-step error norm (and every
 * other norm the step-size logic takes) is summed over ranks through peer-mapped mailboxes.
 * mailboxes[r] is the DEVICE-visible address of rank r's mailbox (lrnde_ctx_mailbox on r,
 * exchanged by the host through CUDA IPC).  total_batch = sum of all ranks' B. */
int lrnde_ctx_mailbox(lrnde_ctx* ctx, void** dev_ptr, uint64_t* bytes);
int lrnde_ctx_set_dist(lrnde_ctx* ctx, int rank, int nranks, void* const* mailboxes,
                       int64_t total_batch);

int lrnde_model_create(lrnde_ctx* ctx, const lrnde_layer_desc* layers, int nlayers,
                       int time_dependent, int input_act, lrnde_model** out);
/* Conv dynamics on a WHCN state [width, height, channels, B] (state dims D = width*height*channels; the
 * layer functor, solve, regulariser and adjoint entry points below take the model unchanged).  Replaces
 * TDChain(Chain(Chain(Conv, BatchNorm), ..., Conv)) applied through the `dudt` closure (neural_ode.jl:44-48,
 * src/layers/common.jl:10-45).  width % 4 == 0, 4 <= width <= 32; the last layer is a plain Conv.  BatchNorm
 * couples the samples of a batch, so a model with batchnorm layers runs on a single-GPU ctx only. */
int lrnde_conv_model_create(lrnde_ctx* ctx, const lrnde_conv_layer_desc* layers, int nlayers, int width,
                            int height, int time_dependent, lrnde_model** out);
int lrnde_model_destroy(lrnde_model* m);
int64_t lrnde_model_nparams(const lrnde_model* m);
int64_t lrnde_model_state_dims(const lrnde_model* m);

/* One f(u, ps, t) evaluation (the `dudt` closure, neural_ode.jl:45-48); parity hook. */
int lrnde_dynamics_eval(lrnde_ctx* ctx, const lrnde_model* m, const lrnde_opts* o,
                        const float* ps, const float* u, float t, int64_t B, float* du);

/* (J_u^T lam, J_p^T lam) of one f evaluation: what Zygote.pullback((u,p) -> dudt(u,p,t)) returns inside
 * SciMLSensitivity's ZygoteVJP (neural_ode.jl:11); parity hook.  a [D,B], dps [nparams]. */
int lrnde_dynamics_vjp(lrnde_ctx* ctx, const lrnde_model* m, const lrnde_opts* o, const float* ps,
                       const float* u, float t, const float* lam, int64_t B, float* a, float* dps);

/* Forward functor.  u_save receives stats->nsave_out blocks of [D,B]; capacity must be
 * nsave (or 1 when nsave==0; when nsave==-1 pass u_save_cap blocks, the LAST accepted
 * states are written oldest-first and nsave_out reports how many exist).  In unbiased mode
 * with nsave==0 the reference returns [u(t1), u(t2)] (neural_ode.jl:108): two blocks.
 * save_times (HOST, nullable) receives the corresponding times. */
int lrnde_ode_forward(lrnde_ctx* ctx, const lrnde_model* m, const lrnde_opts* o,
                      const float* ps, const float* x, int64_t B, float* u_save,
                      int64_t u_save_cap, float* save_times, lrnde_stats* stats,
                      lrnde_tape** tape);

/* Pullback.  d_u_save: cotangent blocks matching what forward wrote (NULL = all zero),
 * d_reg: cotangent of st'.reg_val.  Outputs d_ps [P] and d_x [D,B] (d reg/d x == 0,
 * test/runtests.jl:129).  Consumes nothing: call lrnde_tape_free when done. */
int lrnde_ode_backward(lrnde_ctx* ctx, const lrnde_model* m, lrnde_tape* tape,
                       const float* d_u_save, float d_reg, float* d_ps, float* d_x,
                       lrnde_stats* stats);
int lrnde_tape_free(lrnde_tape* tape);

/* Parity/diagnostics: every ATTEMPTED step of the last solve recorded on `tape`
 * (which = 0 forward, 1 adjoint).  HOST arrays of capacity cap; *n = number of steps. */
int lrnde_step_log(const lrnde_tape* tape, int which, float* t, float* dt, float* eest,
                   uint8_t* accepted, int cap, int* n);

/* _perform_step(::FourStageSRIConstantCache) for diagonal noise (perform_step.jl:49-106):
 * one SOSRI step from (uprev, t, dt) with injected Wiener increments dW, dZ [D,B].
 * drift / diffusion are two models over the same state; ps_drift / ps_diffusion their
 * flat parameters.  Writes u [D,B] and *reg_val = EEst*dt.  delta = integrator.opts.delta. */
int lrnde_sosri_step(lrnde_ctx* ctx, const lrnde_model* drift, const lrnde_model* diffusion,
                     const lrnde_opts* o, const float* ps_drift, const float* ps_diffusion,
                     const float* uprev, const float* dW, const float* dZ, float t, float dt,
                     float delta, int64_t B, float* u, float* reg_val);

/* ---------------------------------------------------------------------------------------
 * Neural SDE (diagonal noise): replaces (n::NeuralDSDE)(x, ps, st), src/layers/neural_sde.jl:82-123
 *   lrnde_sde_forward   solve(SDEProblem(dudt, g, x, tspan, ps), SOSRI(); saveat, maxiters, abstol,
 *                       reltol) (neural_sde.jl:44-72) + _get_dsde_integrator (:35-40) +
 *                       _perform_step(::FourStageSRIConstantCache) (src/perform_step.jl:49-106)
 *   lrnde_sde_backward  the TrackerAdjoint pullback selected at neural_sde.jl:12
 * drift / diffusion: two lrnde_model over the same state (ps.drift / ps.diffusion, :57,:63).
 * The reference's Wiener process is unseeded (neural_sde.jl:68-69); here the increments come
 * from Philox4x32-10 keyed by `seed` (counter = element, draw, stream), reproducible on CPU.
 * --------------------------------------------------------------------------------------- */
typedef struct lrnde_sde_tape lrnde_sde_tape;

typedef struct {
  float t0, t2;          /* tspan (neural_sde.jl:13) */
  float abstol, reltol;  /* splatted kwargs (neural_sde.jl:69) */
  int32_t maxiters;      /* neural_sde.jl:14 */
  int32_t reg_mode;      /* LRNDE_REG_*; eval mode => LRNDE_REG_NONE (neural_sde.jl:86,108) */
  float t1;              /* unbiased: host-sampled t1 (:89); must be one of saveat[] (:90-91) */
  float u01;             /* biased: uniform [0,1) picking sol.t[1:end-1] (:112) */
  const float* saveat;   /* HOST pointer, nsave sorted times in [t0,t2] */
  int32_t nsave;         /* >0: save at saveat[] (linear interpolation inside a step); 0: only t2;
                            -1: every accepted step (fetch with lrnde_sde_states) */
  int32_t save_start;    /* with nsave==-1: include t0 */
  uint64_t seed;         /* Wiener process seed */
  int32_t pow_mode;      /* LRNDE_POW_* */
  int32_t host_buffers;  /* 1: ps / x / u_save / d_* are HOST pointers */
  int32_t keep_tape;     /* 1: return a tape for lrnde_sde_backward */
  float qmax, gamma, qmin, delta; /* controller constants; 0 => StochasticDiffEq defaults for
                                     SOSRI (9/8, 9/10, 1/5, 1/6) */
  int32_t reserved[8];
} lrnde_sde_opts;

typedef struct {
  int32_t nfe_drift, nfe_diffusion; /* st'.nfe_drift / nfe_diffusion (neural_sde.jl:44-65) */
  int32_t naccept, nreject;
  int32_t retcode;      /* LRNDE_RET_* */
  float reg_val;        /* st'.reg_val = EEst*dt (perform_step.jl:105) */
  float t1_used, dt_reg;
  int32_t nsave_out;    /* blocks written to u_save (nsave==-1: states available) */
  int32_t ndraws;       /* Philox draws consumed (next free draw index) */
  int32_t gpu_launches;
  int32_t reserved[8];
} lrnde_sde_stats;

/* x: [D,B]; u_save: [D,B,nsave] (ignored for nsave==-1). */
int lrnde_sde_forward(lrnde_ctx* ctx, const lrnde_model* drift, const lrnde_model* diffusion,
                      const lrnde_sde_opts* o, const float* ps_drift, const float* ps_diffusion,
                      const float* x, int64_t B, float* u_save, lrnde_sde_stats* stats,
                      lrnde_sde_tape** tape);
/* d_u_save: one [D,B] cotangent block per returned state (NULL = zeros); d_reg: cotangent of
 * reg_val.  Writes d_ps_drift, d_ps_diffusion (flat parameter layouts) and d_x (may be NULL). */
int lrnde_sde_backward(lrnde_ctx* ctx, lrnde_sde_tape* tape, const float* d_u_save, float d_reg,
                       float* d_ps_drift, float* d_ps_diffusion, float* d_x);
int lrnde_sde_tape_free(lrnde_sde_tape* tape);
/* accepted states first .. first+count-1 of the last solve (state 0 = x) and their times (HOST). */
int lrnde_sde_states(lrnde_ctx* ctx, const lrnde_sde_tape* tape, int32_t first, int32_t count,
                     int32_t host_buffers, float* u_out, float* t_out);
/* every ATTEMPTED step (t, dt, EEst, accepted); HOST arrays of capacity cap. */
int lrnde_sde_step_log(const lrnde_sde_tape* tape, float* t, float* dt, float* eest,
                       uint8_t* accepted, int32_t cap, int32_t* n);

/* The other in-tree SDE local-reg steps for diagonal noise with injected increments dW [D,B]:
 * _perform_step(::RKMilCommuteConstantCache) src/perform_step.jl:108-170 (quirk kept: the
 * error computed at :162-163 is overwritten at :165, EEst = RMS relative change) and
 * _perform_step(::LambaEulerHeunConstantCache) :172-206.  Writes u [D,B], *reg_val = EEst*dt. */
enum { LRNDE_SDESTEP_RKMIL = 0, LRNDE_SDESTEP_LAMBA_EULER_HEUN = 1 };
int lrnde_sde_aux_step(lrnde_ctx* ctx, const lrnde_model* drift, const lrnde_model* diffusion,
                       int32_t kind, const float* ps_drift, const float* ps_diffusion,
                       const float* uprev, const float* dW, float t, float dt, float abstol,
                       float reltol, float delta, int64_t B, int32_t host_buffers, float* u,
                       float* reg_val);

/* Latent-ODE encoder: Recurrence(LatentGRUCell(in, h, latent)) of the physionet model,
 * src/layers/latent_ode.jl:1-48 (both quirks kept: new_y_mean uses new_state_std, :37; the mask sums the
 * mask rows and the dt row, :40) applied along the time dimension as at experiments/src/construct.jl:231.
 * F = feature rows of x (2*in + 1); ps = [update_gate; reset_gate; new_state], each Chain(Dense(2L+F => H,
 * tanh), Dense(H => L | L | 2L, sigmoid | sigmoid | tanh)) in ComponentArray order.  x: [F,T,B] column-major;
 * y: [2L,B] = vcat(y_mean, y_std) after the last step.  The pullback is w.r.t. the parameters (x is data). */
typedef struct lrnde_gru_tape lrnde_gru_tape;
int64_t lrnde_gru_nparams(int32_t F, int32_t H, int32_t L);
int lrnde_gru_forward(lrnde_ctx* ctx, int32_t F, int32_t H, int32_t L, const float* ps, const float* x,
                      int32_t T, int64_t B, int32_t host_buffers, int32_t keep_tape, float* y,
                      lrnde_gru_tape** tape);
int lrnde_gru_backward(lrnde_ctx* ctx, lrnde_gru_tape* tape, const float* d_y, float* d_ps);
int lrnde_gru_tape_free(lrnde_gru_tape* tape);

/* The small layers either side of the latent ODE (experiments/src/construct.jl:229-247).
 * lrnde_mlp_*: a Chain of Dense layers (rec_to_gen; gen_to_data applied to the N = T*B saved columns),
 * x [in,N] -> y [out,N]; the backward recomputes the forward from x and returns d_x (may be NULL), d_ps. */
int lrnde_mlp_forward(lrnde_ctx* ctx, const lrnde_layer_desc* layers, int32_t nlayers, const float* ps,
                      const float* x, int64_t N, int32_t host_buffers, float* y);
int lrnde_mlp_backward(lrnde_ctx* ctx, const lrnde_layer_desc* layers, int32_t nlayers, const float* ps,
                       const float* x, const float* d_y, int64_t N, int32_t host_buffers, float* d_x,
                       float* d_ps);
/* ReparameterizeLayer (src/layers/common.jl:48-77): x [2L,B] = vcat(mu, logsigma2) -> y = mu + exp(logsigma2/2)*eps,
 * eps from Philox(seed) (the reference uses randn_like(rng, ...)); training == 0: y = mu.  Pass d_x (and any of
 * d_y, d_mu, d_logsigma2, the latter two being the KL cotangents) to get the pullback in the same call. */
int lrnde_reparameterize(lrnde_ctx* ctx, const float* x, int32_t L, int64_t B, uint64_t seed, int32_t training,
                         int32_t host_buffers, float* y, const float* d_y, const float* d_mu,
                         const float* d_ls, float* d_x);
/* Latent-ODE loss -mean(log_likelihood - w_kl * kl_div) (experiments/src/construct.jl:36-70 with
 * log_likelihood_loss / kl_divergence of experiments/src/utils.jl:94-101; w_reg * reg_val is added by the
 * caller).  pred/data/mask [F,T,B], mu/logsigma2 [L,B]; out3 (HOST) = {loss, -mean(ll), mean(kl)}. */
int lrnde_latent_loss(lrnde_ctx* ctx, const float* pred, const float* data, const float* mask, const float* mu,
                      const float* logsigma2, int32_t F, int32_t T, int32_t L, int64_t B, float w_kl,
                      int32_t host_buffers, float* out3, float* d_pred, float* d_mu, float* d_logsigma2);

/* Layers either side of the cifar10 NeuralODE (experiments/src/construct.jl:220-227), WHCN arrays
 * [width, height, C, B]; backward calls recompute what they need from x (no tape).
 * lrnde_conv2d_*: Lux Conv((3,3), in_ch => out_ch, act; pad=(1,1)) -- the AugmenterLayer's Conv 3 => 5
 * (src/layers/common.jl:80-92; the concatenation with x is the caller's) and the classifier's Conv 8 => 1, gelu.
 * ps = weight[3,3,in_ch,out_ch] then (use_bias) bias[out_ch].  d_x may be NULL. */
int lrnde_conv2d_forward(lrnde_ctx* ctx, int32_t in_ch, int32_t out_ch, int32_t use_bias, int32_t act, int32_t width,
                         int32_t height, const float* ps, const float* x, int64_t B, int32_t host_buffers, float* y);
int lrnde_conv2d_backward(lrnde_ctx* ctx, int32_t in_ch, int32_t out_ch, int32_t use_bias, int32_t act,
                          int32_t width, int32_t height, const float* ps, const float* x, const float* d_y, int64_t B,
                          int32_t host_buffers, float* d_x, float* d_ps);
/* lrnde_batchnorm_*: Lux BatchNorm(C, act) on [HW, C, B]: ps = scale[C] then bias[C]; state (nullable unless
 * testmode) = running_mean[C] then running_var[C], updated in place in training mode (momentum 0.1, unbiased
 * variance), used for normalisation when testmode. */
int lrnde_batchnorm_forward(lrnde_ctx* ctx, int32_t C, int64_t HW, int32_t act, const float* ps, const float* x,
                            int64_t B, float* state, int32_t testmode, int32_t host_buffers, float* y);
int lrnde_batchnorm_backward(lrnde_ctx* ctx, int32_t C, int64_t HW, int32_t act, const float* ps, const float* x,
                             const float* d_y, int64_t B, const float* state, int32_t testmode, int32_t host_buffers,
                             float* d_x, float* d_ps);

/* Next row of the path: classifier Dense(D => C) + logitcrossentropy,
 * experiments/src/construct.jl:199 and experiments/src/utils.jl:88, with its pullback.
 * Wc: flat [C x D] weight then [C] bias; u: [D,B]; labels: class index per sample.
 * Writes *loss (HOST), d_u [D,B] and d_Wc [C*D + C] (either may be NULL). */
int lrnde_head_ce(lrnde_ctx* ctx, const float* Wc, const float* u, const int32_t* labels,
                  int64_t B, int32_t D, int32_t C, int32_t host_buffers, float* loss, float* d_u,
                  float* d_Wc);

/* Adam update on device buffers (Optimisers.jl Adam as built in
 * experiments/src/construct.jl:104-152); step >= 1 is the 1-based iteration count. */
int lrnde_adam_step(lrnde_ctx* ctx, float* p, const float* g, float* m, float* v, int64_t n,
                    float lr, float beta1, float beta2, float eps, int32_t step);

/* The other Optimisers.jl rules experiments/src/construct.jl:104-125 can build: Descent / Momentum(lr, a) /
 * Nesterov(lr, a) (cfg.optimizer == "sgd"), Adam(lr, (a, b), eps) (also "adamw": AdamW(lr) is Adam chained with
 * WeightDecay(0)) and AdaMax(lr, (a, b), eps) (the physionet config, experiments/physionet/physionet.yml:27), each
 * optionally chained with WeightDecay(weight_decay) (construct.jl:124-126: the decay term weight_decay * p is added
 * to the rule's update).  s1 / s2: state buffers of n floats (momentum or first moment; second moment or
 * infinity norm), zero-initialised by the caller, NULL where the rule has none.  DEVICE pointers. */
enum { LRNDE_OPT_DESCENT = 0, LRNDE_OPT_MOMENTUM = 1, LRNDE_OPT_NESTEROV = 2, LRNDE_OPT_ADAM = 3, LRNDE_OPT_ADAMAX = 4 };
int lrnde_opt_step(lrnde_ctx* ctx, int32_t kind, float* p, const float* g, float* s1, float* s2, int64_t n,
                   float lr, float a, float b, float eps, float weight_decay, int32_t step);

/* Sum of the DEVICE vector v[n] over the data-parallel group of lrnde_ctx_set_dist, in place, identical bits on
 * every rank (peer loads over NVLink through the mailboxes; fixed rank order): the once-per-iteration all-reduce
 * of the parameter gradients (Zygote's gradient of the global-mean loss in the reference's
 * FluxMPI-less single-process runs).  Every rank must call it in the same order.  No-op on a single-rank ctx. */
int lrnde_allreduce_sum(lrnde_ctx* ctx, float* v, int64_t n);

/* The saved states of a kept solution, after the call: lrnde_ode_forward may run with u_save = NULL (every-step
 * saves of the :biased path, neural_ode.jl:86-100: the count is only known afterwards, stats->nsave_out) and the
 * caller then sizes u_save.  Pointers as in the forward call (HOST when it had host_buffers). */
int lrnde_ode_saved_states(lrnde_ctx* ctx, const lrnde_model* m, lrnde_tape* tape, float* u_save,
                           int64_t u_save_cap, float* save_times);

/* One gradient evaluation of the MNIST classifier Chain(FlattenLayer, NeuralODE, diffeqsol_to_array,
 * Dense(D => C)) (experiments/src/construct.jl:180-200) under logitcrossentropy + w_reg * reg_val
 * (construct.jl:19-31): what Zygote.pullback computes inside run_training_step (experiments/src/utils.jl:106-115).
 * u(t2) and its cotangent stay on the device; with o->host_buffers only x [D,B], labels (0-based), ps and Wc
 * ([C x D] weight then [C] bias) go in and loss_ce (HOST scalar, the cross-entropy term), d_ps, d_Wc come out.
 * grad_scale multiplies the cross-entropy cotangent (1 / nranks for a global-mean loss).  stats: forward fields
 * of lrnde_ode_forward plus the *_bwd fields. */
int lrnde_classifier_grad(lrnde_ctx* ctx, const lrnde_model* m, const lrnde_opts* o, const float* ps,
                          const float* Wc, const float* x, const int32_t* labels, int64_t B, int32_t C,
                          float w_reg, float grad_scale, float* loss_ce, float* d_ps, float* d_Wc,
                          lrnde_stats* stats);

/* Input pipeline of lrnde_classifier_grad with host buffers: starts the host -> device copy of a batch (x [D,B] and
 * its 0-based labels, PINNED host memory for the copy to overlap) on the library's copy stream and returns at once
 * (the copy is queued behind the next lrnde_classifier_grad's own small parameter copies, or at the next prefetch).
 * A later lrnde_classifier_grad called with the same host pointers and B uses the staged device copy instead of
 * copying inside the call.  Two batches can be in flight (the one being consumed and the next one); the host
 * buffers must stay unchanged until the consuming call returns.  The reference's counterpart is the DataLoader
 * batch moved with `gpu` before run_training_step (experiments/src/utils.jl:106-115). */
int lrnde_prefetch_inputs(lrnde_ctx* ctx, const float* x, const int32_t* labels, int64_t B, int32_t D);

/* Microseconds per launch of one attempt of the latent-space adjoint {chain, lambda GEMM, pairacc, reduce + mu,
 * whole attempt}, measured with CUDA events by the last lrnde_ode_backward that ran with the environment variable
 * LRNDE_PROFILE_ADJ=<iterations> set (roofline probe of bench.py). */
int lrnde_profile_adjoint_last(float* us5);

/* Roofline probe of the latent-space forward engine: `iters` repetitions of one Tsit5 attempt (same descriptors),
 * CUDA events on the ctx stream; us[0] = chain kernel, us[1] = kgemm kernel, us[2] = the attempt (both). */
int lrnde_profile_step(lrnde_ctx* ctx, const lrnde_model* m, const lrnde_opts* o, const float* ps, const float* x,
                       int64_t B, int32_t iters, float* us);

/* Roofline probe: times `iters` back-to-back f(u, ps, t) evaluations (DEVICE pointers) with
 * CUDA events on the ctx stream; *ms_per_eval is the mean, *launches_per_eval the kernels one
 * evaluation launches. */
int lrnde_profile_feval(lrnde_ctx* ctx, const lrnde_model* m, const lrnde_opts* o,
                        const float* ps, const float* u, int64_t B, int32_t iters, float* du,
                        float* ms_per_eval, int32_t* launches_per_eval);

/* CUDA IPC plumbing for lrnde_ctx_set_dist when every rank is its own process: export a
 * device allocation of this ctx as a 64-byte handle / map a peer's handle. */
int lrnde_ipc_export(lrnde_ctx* ctx, void* dev_ptr, void* handle64_out);
int lrnde_ipc_open(lrnde_ctx* ctx, const void* handle64, void** dev_ptr_out);

#ifdef __cplusplus
}
#endif
#endif /* LRNDE_H */
