// Tensor-core engine of the conv dynamics (BASELINE configs[3] "cifar10", "implicit-GEMM Conv stages"):
//
//   TDChain(Chain(Chain(Conv((3,3), 9 => 64; pad=1, use_bias=false), BatchNorm(64, gelu)),
//                 Chain(Conv((3,3), 65 => 64; pad=1, use_bias=false), BatchNorm(64, gelu)),
//                 Conv((3,3), 65 => 8; pad=1, use_bias=false)))      experiments/src/construct.jl:212-218
//
// Every 3x3 convolution (forward, its transposed twin of the reverse pass, and the weight gradient) is an implicit
// GEMM on tcgen05 in 3xTF32.  No im2col matrix exists anywhere:
//
//   * convtc_pack_kernel applies whatever sits between two convolutions (stage combination of perform_step.jl:11-18,
//     BatchNorm scale / shift + activation, or the BatchNorm pullback) ONCE per element, splits the result into tf32
//     hi / lo parts and writes it
//       (F) channel-interleaved with a zero halo: F[c / 4][guard + b * (W+2)(H+2) + (y+1)(W+2) + (x+1)][c % 4], which
//           is the K-major no-swizzle UMMA layout with a uniform 16-byte row stride: rows = positions, K = channels.
//           A tap (dy, dx) of the convolution is the SAME shared-memory patch read through a descriptor whose start
//           address is moved by (dy (W+2) + dx) rows -- nine taps, one copy of the patch, plain bulk copies;
//       (P) for cotangents, a plain [W, H, C, B] fp32 copy: the weight gradient (K = pixels of an image row) reads its
//           operands -- raw conv outputs, y(t), cotangents -- through tensor-map TMA (K-major SWIZZLE_128B row tiles) and
//           applies BatchNorm + activation, the hi / lo split and the dx shifts in shared memory.
//   * conv_kernel<NOUT, MODE, SRC>: persistent CTAs, M = 128 consecutive (haloed) positions, N = output channels, K = 8
//     input channels per MMA; a group of four M tiles shares every weight stage, accumulators double-buffered in TMEM
//     (2 x 4 x NOUT columns), epilogue = time-channel term + scale + coalesced [W,H,C,B] stores + BatchNorm partial
//     sums (MODE 1: of the raw output; MODE 2: of the BatchNorm pullback that follows a data-gradient convolution).
//     The TDChain time channel (common.jl:19-33: t inside the image, 0 in the padding) never enters the GEMM: its
//     contribution t * sum_{taps inside} w[tap, time, co] depends only on the border class of the pixel.  SRC = 1: the
//     (F) operand is not read from memory at all -- eight producer warps form it in the stage from the stored raw output
//     of the previous layer (ConvTcP::srcZ).
//   * wgrad_kernel<NB>: dW[tap, ci, co] = sum_p X[p + d_tap, ci] Delta[p, co]; see below and lrnde_conv_tc.cu.
#pragma once
#include "lrnde_host.h"

constexpr int kCtGuard = 36;     // positions in front of / behind the packed arrays (>= W + 3)
constexpr int kCtGroup = 512;    // positions per tile group (4 M tiles of 128)

struct ConvTcGeom {
  int Wd = 0, Ht = 0, B = 0, PW = 0, PH = 0, IMG = 0;
  long NP = 0, NPA = 0;   // haloed positions of the batch; allocated positions per channel chunk
  int ngroups = 0;
  void set(int wd, int ht, int b) {
    Wd = wd; Ht = ht; B = b; PW = wd + 2; PH = ht + 2; IMG = PW * PH;
    NP = (long)b * IMG;
    ngroups = (int)((NP + kCtGroup - 1) / kCtGroup);
    NPA = (long)ngroups * kCtGroup + 2 * kCtGuard;
  }
  size_t f_floats(int C) const { return (size_t)(C / 4) * (size_t)NPA * 4; }   // one of the hi / lo arrays
};

// what convtc_pack_kernel computes per element before the hi / lo split
struct ConvTcPackP {
  const float* X; const LinComb* xdesc;     // source [W,H,C,B]: plain array or stage combination
  float* side; int side_to_desc_dst;        // also store the combined source (u_{n+1} / y(t)): plain pointer or xdesc->dst
  const float* in_ab; int in_act;           // v <- act(a[c] v + b[c])  (BatchNorm scale / shift + activation)
  // BatchNorm + activation pullback (X = z, the raw conv output the statistics were taken on):
  //   v <- a (g act'(a z + b) - m1 - (z - mean) invstd m2); coef == nullptr: v <- g act'(z)
  const float* bwd_g; const float* bwd_stat; const float* bwd_coef; int bwd_act;
  float* Fhi; float* Flo;                   // (F) layout (may be null)
  float* Pv;                                // plain fp32 [W,H,C,B] copy of the result: operand of the weight gradient (may be null)
  float* rowsum;                            // [B * Ht][C][3]: (sum, first, last) of every image row of the result (Wd == 32; may be null)
  int C;
  const int* done;
};
void convtc_pack(lrnde_ctx* ctx, const ConvTcGeom& g, const ConvTcPackP& p);

// weight image of one convolution: [stage of 8 input channels][hi | lo][tap][k4][NOUT rows][4] + the time-channel table
size_t convtc_wimg_bytes(int K, int NOUT);
inline int convtc_nout(int cout) { return cout <= 16 ? 16 : 64; }
// w: Lux weight [3,3,CinTot,Cout]; transposed = 0: out channels = Cout, K = first `K` input channels, tsum = table of the
// channel CinTot - 1 when td; transposed = 1: the data-gradient convolution (K = Cout, out channels = first `keep` inputs)
void convtc_wpack(lrnde_ctx* ctx, const float* w, int CinTot, int Cout, int transposed, int keep, int td, uint8_t* img,
                  float* tsum);

struct ConvTcP {
  const float* Fhi; const float* Flo;
  // alternative source of the A operand: a stored [W,H,K,B] array with v <- act(a[c] v + b[c]) applied on the fly
  // (src_ab null: identity scale / shift); the (F) image is not read then
  const float* srcZ; const float* src_ab; int src_act;
  const uint8_t* Wimg; int K;               // input channels (multiple of 8)
  const float* tsum; const LinComb* tdesc;  // time channel (null: none)
  float* Y; const LinComb* ydesc; float out_scale; int Cout;
  float2* stat_part;                        // [convtc_stat_rows(g, Cout)][Cout] (sum, sum of squares) of the raw output, or null
  // data-gradient convolution followed by a BatchNorm + activation pullback: the statistics are the two sums of that
  // pullback instead, (sum ghat, sum ghat xhat) with ghat = g act'(a z + b), xhat = (z - mean) invstd  (z: [W,H,Cout,B])
  const float* bwd_z; const float* bwd_ab; const float* bwd_stat; int bwd_act;
  const int* done;
};
void convtc_conv(lrnde_ctx* ctx, const ConvTcGeom& g, const ConvTcP& p);
inline int convtc_stat_rows(const ConvTcGeom& g, int cout) { return g.ngroups * (cout <= 16 ? 8 : 4); }

// weight gradient on the tensor cores (Wd == 32): dW[tap, ci, co] = sum_{b,y,x} X[x + dx, y + dy, ci, b] Delta[x, y, co, b]
// as GEMMs over the pixels of an image row (K = 32): the operand with more channels is "P" (M = 64 channels x 2 dx shifts
// per MMA, tensor-map TMA delivers the three dx-shifted, zero-padded copies of a row), the other one "Q" (N = its
// channels); the dy taps pair P row yy with Q rows yy - dy.  The image rows of the batch are split over the CTAs;
// part[split][Lux weight layout] is summed by wgrad_reduce_kernel in fixed order.  The time channel's weight gradient
// t * sum_{pixels whose tap stays inside the image} Delta is formed from nine masked sums by convtc_time_wgrad_kernel.
struct ConvTcWgP {
  // input of the convolution [W,H,Cx,B] (time channel excluded) = act(a[c] X + b[c]) of a stored array (x_ab / x_act as in
  // ConvTcPackP::in_ab / in_act; null / ACT_IDENTITY: X itself)
  const float* X; const float* x_ab; int x_act; int Cx;
  const float* D; int Cd;                         // cotangent of its output [W,H,Cd,B], fp32
  int CinTot;                                     // Cx + td: layout of the result
  float* part; size_t block;                      // block = 9 * CinTot * Cd floats per split
  const LinComb* tdesc;                           // time channel (null: none)
  const float* Drowsum;                           // per-row sums of Delta written by the pack kernel (time channel only)
  const int* done;
};
bool convtc_wgrad_ok(const ConvTcGeom& g);         // Wd == 32
int convtc_wgrad_splits(const ConvTcGeom& g);
void convtc_wgrad(lrnde_ctx* ctx, const ConvTcGeom& g, const ConvTcWgP& p);
