// Scalar step-size logic of the adaptive explicit solve, shared by the device controller
// kernel and a host build (liblrnde_hostcheck.so) that the CPU tests pin bit-for-bit against
// oracle/lrnde_oracle.py.
//
// What it restates (un-vendored OrdinaryDiffEq / DiffEqBase; the
// reference reaches it through `solve(prob, Tsit5(); ...)` at src/layers/neural_ode.jl:51 and
// `init(...)` at src/utils.jl:51):
//   lr_fastpow            DiffEqBase.fastpow (2023 Float32 approximation)
//   lr_ctrl_header        loopheader! + check_error!
//   lr_ctrl_footer        loopfooter! with the PI controller
//   lr_initdt_a/_b        ode_determine_initdt (out-of-place)
//   lr_tsit5_*            Tsit5 tableau + free 4th-order interpolant weights
//
// Every float operation is a single correctly-rounded IEEE op (no FMA contraction) so host
// and device agree exactly.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define LR_HD __host__ __device__ __forceinline__
#else
#define LR_HD inline
#endif

#if defined(__CUDA_ARCH__)
LR_HD float lr_mul(float a, float b) { return __fmul_rn(a, b); }
LR_HD float lr_add(float a, float b) { return __fadd_rn(a, b); }
LR_HD float lr_sub(float a, float b) { return __fsub_rn(a, b); }
LR_HD float lr_div(float a, float b) { return __fdiv_rn(a, b); }
LR_HD uint32_t lr_f2u(float x) { return __float_as_uint(x); }
LR_HD float lr_u2f(uint32_t b) { return __uint_as_float(b); }
#else
// volatile keeps the host compiler from contracting a*b+c into an fma
LR_HD float lr_mul(float a, float b) { volatile float r = a * b; return r; }
LR_HD float lr_add(float a, float b) { volatile float r = a + b; return r; }
LR_HD float lr_sub(float a, float b) { volatile float r = a - b; return r; }
LR_HD float lr_div(float a, float b) { volatile float r = a / b; return r; }
LR_HD uint32_t lr_f2u(float x) { uint32_t b; memcpy(&b, &x, 4); return b; }
LR_HD float lr_u2f(uint32_t b) { float x; memcpy(&x, &b, 4); return x; }
#endif

// Python's min/max argument semantics (the oracle is the definition): min(a,b) = b if b<a else a
LR_HD float lr_pymin(float a, float b) { return (b < a) ? b : a; }
LR_HD float lr_pymax(float a, float b) { return (b > a) ? b : a; }

// spacing(x) of a non-negative Float32 (eps(x) in Julia, np.spacing)
LR_HD float lr_spacing(float x) {
  x = fabsf(x);
  if (!(x < INFINITY)) return NAN;
  return lr_sub(lr_u2f(lr_f2u(x) + 1u), x);
}

// ---------------------------------------------------------------- DiffEqBase.fastpow (2023)
LR_HD float lr_fastlog2(float x) {
  uint32_t bits = lr_f2u(x);
  int e = (int)((bits & 0x7F800000u) >> 23);
  float s, fe;
  if (bits & 0x00400000u) {
    s = lr_u2f((bits & 0x007FFFFFu) | 0x3F000000u);
    fe = lr_sub((float)e, 126.0f);
  } else {
    s = lr_u2f((bits & 0x007FFFFFu) | 0x3F800000u);
    fe = lr_sub((float)e, 127.0f);
  }
  s = lr_sub(s, 1.0f);
  float num = lr_mul(s, lr_add(lr_mul(0.338953f, s), 2.198599f));
  float den = lr_add(s, 1.523692f);
  return lr_add(fe, lr_div(num, den));
}

LR_HD float lr_fastpow2(float x) {
  float offset = (x < 0.0f) ? 1.0f : 0.0f;
  float clipp = (x < -126.0f) ? -126.0f : x;
  float w = (float)((int)clipp);  // trunc toward zero
  float z = lr_add(lr_sub(clipp, w), offset);
  float inner = lr_add(lr_add(clipp, 121.2740575f), lr_div(27.7280233f, lr_sub(4.84252568f, z)));
  inner = lr_sub(inner, lr_mul(1.49012907f, z));
  float v = lr_mul(8388608.0f, inner);
  return lr_u2f((uint32_t)(long long)v);
}

// pow_mode: 0 = fastpow_2023, 1 = exact (libm pow in double, rounded to Float32)
LR_HD float lr_fastpow(float x, float y, int pow_mode) {
  if (pow_mode == 1) return (float)pow((double)x, (double)y);
  if (x == 0.0f) return 0.0f;
  if (isinf(x) && isinf(y)) return INFINITY;
  return lr_fastpow2(lr_mul(y, lr_fastlog2(x)));
}

// ---------------------------------------------------------------- Tsit5 tableau (App. A.1)
#define LR_TSIT5_C1 0.161f
#define LR_TSIT5_C2 0.327f
#define LR_TSIT5_C3 0.9f
#define LR_TSIT5_C4 0.9800255409045097f

// a[j][i]: stage j+2 (j = 0..5), i.e. rows a2*, a3*, ..., a7*
LR_HD float lr_tsit5_a(int row, int i) {
  const float a[6][6] = {
      {0.161f, 0, 0, 0, 0, 0},
      {-0.008480655492356989f, 0.335480655492357f, 0, 0, 0, 0},
      {2.8971530571054935f, -6.359448489975075f, 4.3622954328695815f, 0, 0, 0},
      {5.325864828439257f, -11.748883564062828f, 7.4955393428898365f, -0.09249506636175525f, 0, 0},
      {5.86145544294642f, -12.92096931784711f, 8.159367898576159f, -0.071584973281401f,
       -0.028269050394068383f, 0},
      {0.09646076681806523f, 0.01f, 0.4798896504144996f, 1.379008574103742f, -3.290069515436081f,
       2.324710524099774f}};
  return a[row][i];
}
LR_HD float lr_tsit5_c(int row) {  // c of stage row+2
  const float c[6] = {LR_TSIT5_C1, LR_TSIT5_C2, LR_TSIT5_C3, LR_TSIT5_C4, 1.0f, 1.0f};
  return c[row];
}
LR_HD float lr_tsit5_btilde(int i) {
  const float b[7] = {-0.00178001105222577714f, -0.0008164344596567469f, 0.007880878010261995f,
                      -0.1447110071732629f,     0.5823571654525552f,     -0.45808210592918697f,
                      0.015151515151515152f};
  return b[i];
}
// b_i(theta), i = 0..6 (oracle Tableau.interp_weights, Horner form, single-rounded ops)
LR_HD void lr_tsit5_interp(float th, float* b) {
  const float r[7][4] = {{1.0f, -2.763706197274826f, 2.9132554618219126f, -1.0530884977290216f},
                         {0, 0.13169999999999998f, -0.2234f, 0.1017f},
                         {0, 3.9302962368947516f, -5.941033872131505f, 2.490627285651253f},
                         {0, -12.411077166933676f, 30.33818863028232f, -16.548102889244902f},
                         {0, 37.50931341651104f, -88.1789048947664f, 47.37952196281928f},
                         {0, -27.896526289197286f, 65.09189467479366f, -34.87065786149661f},
                         {0, 1.5f, -4.0f, 2.5f}};
  float th2 = lr_mul(th, th);
  b[0] = lr_mul(th, lr_add(r[0][0], lr_mul(th, lr_add(r[0][1], lr_mul(th, lr_add(r[0][2], lr_mul(th, r[0][3])))))));
  for (int i = 1; i < 7; ++i)
    b[i] = lr_mul(th2, lr_add(r[i][1], lr_mul(th, lr_add(r[i][2], lr_mul(th, r[i][3])))));
}

// ---------------------------------------------------------------- controller state
enum { LR_RET_SUCCESS = 0, LR_RET_MAXITERS = 1, LR_RET_DTMIN = 2, LR_RET_UNSTABLE = 3, LR_RET_TAPEFULL = 4,
       LR_RET_PEERTIMEOUT = 5 };

struct LrCtrl {
  float t, dt, dtpropose, qold, q11, tstop, dtmax, dtmin;
  int tdir, iter, maxiters, naccept, nreject, retcode, accepted_prev, first, pow_mode;
};

LR_HD void lr_ctrl_init(LrCtrl& s, float t0, float tend, float dt0, float dtmin, int maxiters,
                        int pow_mode) {
  s.t = t0;
  s.dt = dt0;
  s.dtpropose = dt0;
  s.qold = 1e-4f;
  s.q11 = 1.0f;
  s.tstop = tend;
  s.dtmax = fabsf(lr_sub(tend, t0));
  s.dtmin = dtmin;
  s.tdir = (tend >= t0) ? 1 : -1;
  s.iter = 0;
  s.maxiters = maxiters;
  s.naccept = s.nreject = 0;
  s.retcode = LR_RET_SUCCESS;
  s.accepted_prev = 1;
  s.first = 1;
  s.pow_mode = pow_mode;
}

// loopheader! + check_error!.  Returns 1 when a step attempt must run with s.dt, 0 when the
// segment is over (t reached tstop: retcode stays SUCCESS) or the solve failed (retcode set).
LR_HD int lr_ctrl_header(LrCtrl& s, int u_has_nan) {
  const float td = (float)s.tdir;
  if (!(lr_mul(td, s.t) < lr_mul(td, s.tstop))) return 0;
  const float qmin = lr_div(1.0f, 5.0f), gamma = lr_div(9.0f, 10.0f);
  if (!s.first) {
    if (s.accepted_prev) s.dt = s.dtpropose;
    else s.dt = lr_div(s.dt, lr_pymin(lr_div(1.0f, qmin), lr_div(s.q11, gamma)));
  }
  s.first = 0;
  s.iter += 1;
  s.dt = lr_mul(td, lr_pymin(fabsf(s.dt), s.dtmax));
  s.dt = lr_mul(td, lr_pymax(fabsf(s.dt), s.dtmin));
  const float rem = fabsf(lr_sub(s.tstop, s.t));
  const int clamped = rem <= fabsf(s.dt);
  s.dt = lr_mul(td, lr_pymin(fabsf(s.dt), rem));
  if (s.iter > s.maxiters) { s.retcode = LR_RET_MAXITERS; return 0; }
  if (fabsf(s.dt) <= s.dtmin && !clamped) { s.retcode = LR_RET_DTMIN; return 0; }
  if (u_has_nan) { s.retcode = LR_RET_UNSTABLE; return 0; }
  return 1;
}

// loopfooter!: PI controller, accept test, tstop snapping.  Returns accept (1/0).  On accept
// s.t advances (snapped to tstop within 100 eps) and *dt_taken is the interval actually stored.
LR_HD int lr_ctrl_footer(LrCtrl& s, float EEst, float* dt_taken) {
  const float td = (float)s.tdir;
  const float qmin = lr_div(1.0f, 5.0f), qmax = 10.0f, gamma = lr_div(9.0f, 10.0f);
  const float beta2 = lr_div(2.0f, 25.0f), beta1 = lr_div(7.0f, 50.0f), qoldinit = 1e-4f;
  float q;
  if (EEst == 0.0f) {
    q = lr_div(1.0f, qmax);
  } else {
    s.q11 = lr_fastpow(EEst, beta1, s.pow_mode);
    q = lr_div(s.q11, lr_fastpow(s.qold, beta2, s.pow_mode));
    q = lr_pymax(lr_div(1.0f, qmax), lr_pymin(lr_div(1.0f, qmin), lr_div(q, gamma)));
  }
  const int accept = (EEst <= 1.0f) ? 1 : 0;
  if (accept) {
    s.naccept += 1;
    if (1.0f <= q && q <= 1.0f) q = 1.0f;
    s.qold = lr_pymax(EEst, qoldinit);
    const float dtnew = lr_div(s.dt, q);
    float ttmp = lr_add(s.t, s.dt);
    const float tol = lr_mul(100.0f, lr_spacing(lr_pymax(fabsf(s.t), fabsf(s.tstop))));
    if (fabsf(lr_sub(ttmp, s.tstop)) < tol) ttmp = s.tstop;
    s.dtpropose = lr_mul(td, lr_pymax(lr_pymin(fabsf(dtnew), s.dtmax), s.dtmin));
    if (dt_taken) *dt_taken = lr_sub(ttmp, s.t);
    s.t = ttmp;
  } else {
    s.nreject += 1;
  }
  s.accepted_prev = accept;
  return accept;
}

// ---------------------------------------------------------------- ode_determine_initdt (A.3)
// Part A: from d0 = RMS(u0/sk), d1 = RMS(f0/sk) choose dt0 (signed by tdir).
LR_HD float lr_initdt_a(float d0, float d1, float dtmax, int tdir) {
  float dt0;
  if (d0 < 1e-5f || d1 < 1e-5f) dt0 = 1e-6f;
  else dt0 = lr_mul(0.01f, lr_div(d0, d1));
  dt0 = lr_pymin(dt0, fabsf(dtmax));
  return lr_mul((float)tdir, dt0);
}
// Part B: d2sum_rms = RMS((f1-f0)/sk) (not yet divided by dt0); f_equal = all(f0 == f1).
LR_HD float lr_initdt_b(float dt0_signed, float d1, float d2_rms, int f_equal, float dtmax,
                        float dtmin, int tdir) {
  const float td = (float)tdir;
  const float dt0 = fabsf(dt0_signed);
  dtmax = fabsf(dtmax);
  if (f_equal) return lr_mul(td, lr_pymax(dtmin, lr_mul(100.0f, dt0)));
  const float d2 = lr_div(d2_rms, dt0);
  const float mx = lr_pymax(d1, d2);
  float dt1;
  if (mx <= 1e-15f) dt1 = lr_pymax(1e-6f, lr_mul(dt0, 1e-3f));
  else dt1 = powf(10.0f, lr_div(-(lr_add(2.0f, log10f(mx))), 5.0f));
  // Python min(a, b, c): first minimal element
  float m = lr_mul(100.0f, dt0);
  if (dt1 < m) m = dt1;
  if (dtmax < m) m = dtmax;
  return lr_mul(td, lr_pymax(dtmin, m));
}

// searchsortedfirst-style interval location of the dense forward solution (oracle
// ODESolution.locate): n with ts[n] < tval <= ts[n+1] (mirrored for tdir < 0), clamped to
// [0, nsteps-1].
LR_HD int lr_locate(const float* ts, int nsteps, int tdir, float tval) {
  int lo = 1, hi = nsteps;
  const float d = (float)tdir;
  while (lo < hi) {
    int mid = (lo + hi) / 2;
    if (d * ts[mid] >= d * tval) hi = mid;
    else lo = mid + 1;
  }
  return lo - 1;
}
