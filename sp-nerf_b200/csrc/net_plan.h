// Host+device shared definitions for the fused point-network kernels: the shared-memory map, the
// step list that drives the weight producer and the MMA issuer, the layout of the small fp32
// parameter block, and the per-tile activation save map.
//
// Network being executed: models/spnerf.py:273-369 (SPNeRF.forward) with feat=512, 8 layers,
// skip at 4 (modules/opt.py:43-46 defaults).
#pragma once
#include <stdint.h>
#include "../../include/spnerf_b200.h"

namespace net {

constexpr int kTileM = 128;          // points per tile (= accumulator rows = TMEM lanes)
constexpr int kSlabBytes = 16384;    // 128 rows x 64 halves, 128B-swizzled (sm100.cuh)
constexpr int kActSlabs = 8;         // 512-wide activation tile
constexpr int kSlabInpHi = 8;        // encoded input (fp16 high part); scratch after the skip layer
constexpr int kSlabInpLo = 0;        // encoded input (fp16 residual): aliases activation slab 0, which is
                                     // dead until the first layer's epilogue overwrites it
constexpr int kNumSlabs = 9;
// parameter region: the 16-column "aux" operand + the tiny last-layer weights (resident for the
// whole kernel, read by the epilogue with broadcast shared-memory loads)
constexpr int kOffAux = kNumSlabs * kSlabBytes;      // 128 rows x 16 halves, no-swizzle K-major (4 KB)
constexpr int kOffRgb2 = kOffAux + 4096;             // float4[256]: rgb_from_xyzdir.2 rows (3) per hidden unit
constexpr int kOffSem2 = kOffRgb2 + 4096;            // float4[256][2]: logit_from_label.2 rows (<= 8)
constexpr int kOffWst = kOffSem2 + 8192;
constexpr int kWStageBytes = 16384;  // this CTA's half of one weight item: up to 128 of 256 rows x 64 k
constexpr int kNumWStages = 4;
constexpr int kSmemBars = kOffWst + kNumWStages * kWStageBytes;
constexpr int kOffSun6 = kSmemBars + 256;            // float[256] sun_v_net.6 row
constexpr int kOffBeta2 = kOffSun6 + 1024;           // float[256] beta_from_xyz.2 row
constexpr int kSmemTotal = kOffBeta2 + 1024;
static_assert(kSmemTotal <= 232448, "shared memory budget");
// fp32 block mirrored into the parameter region at kernel start (floats): rgb2 | sem2 | sun6 | beta2
constexpr int kSmallWFloats = 1024 + 2048 + 256 + 256;

constexpr int kFeat = 512;           // widest trunk the shared-memory / tensor-memory map is sized for
constexpr int kHalf = 256;           // ... and its head width; a configuration's own widths are cfg.feat and cfg.feat / 2
// trunk widths the kernels are instantiated for: 512 (modules/opt.py:44 default) and 256 (the SPNeRF class default,
// models/spnerf.py:163); everything below is parameterised by cfg.feat
inline bool feat_supported(int feat) { return feat == 512 || feat == 256; }
constexpr int kMaxSteps = 384;
constexpr int kAuxSlab = 255;        // MmaStep::a_slab value naming the aux operand

// One weight item = one 64-wide K slab of one accumulation chunk (or its 16-wide aux step).
struct __align__(16) MmaStep {
  uint32_t w_off16;   // offset of the packed B tile in the weight blob, in 16-byte units
  uint16_t n;         // B rows (accumulator columns) of this item
  uint16_t tmem_col;  // first accumulator column
  uint8_t a_slab;     // shared-memory slab holding the A operand, or kAuxSlab
  uint8_t ksteps;     // K=16 instructions to issue: 1..4 from this slab, more for an item fused over following slabs
  uint8_t first;      // 1: overwrite the accumulator (first item of a chunk)
  uint8_t last;       // 1: last item of a phase (in ring order) -> both issuers signal the epilogue
  uint16_t bytes16;   // item size in 16-byte units
  uint8_t lane;       // which of the two MMA issuer warps owns this item (its accumulation chunk)
  uint8_t half;       // bit 0: last item of the first accumulator half in a split phase: both issuers signal bar_half when
                      //    they pass it (the epilogue starts on that half while the tensor pipe works on the second one)
                      // bit 1: no later item of the phase reads the A slabs of the first half's columns: both issuers
                      //    signal bar_free, and the epilogue may write that half's new activations in place
};

// Split phases (SPNERF_SPLIT, default on).  A phase that fills the whole accumulator (trunk layers, feats_from_xyz and
// their transposes) is cut into four chunks of a quarter of the columns.  Chunks 0 and 1 -- one per MMA issuer,
// interleaved in the ring -- make up the first accumulator half, chunks 2 and 3 the second.  When the first half has
// retired (bar_half) the epilogue converts it into registers (and saves it to global memory) while the tensor pipe
// works on the second half; its shared-memory image -- the next layer's A operand, in place of the one the MMAs are
// still reading -- is written only after the whole phase has retired.  Every chunk still belongs to ONE issuer, so the
// fp32 summation order, hence the result bits, stay deterministic.
// Measured (round 2, C2 shape): quarter-width chunks (N = 128) run the MMA phase of a 512-wide layer in 11.8 k cycles
// instead of 9.45 k (the A operand is re-read from shared memory twice as often).  The forward's epilogue is short
// (5-6 k cycles), so hiding half of it does not pay for that: the forward keeps whole phases.  The backward's epilogue is
// 11.8 k cycles (it streams the saved activations from HBM), so there the split wins.
#ifndef SPNERF_SPLIT_FWD
#define SPNERF_SPLIT_FWD 0
#endif
#ifndef SPNERF_SPLIT_BWD
#define SPNERF_SPLIT_BWD 1
#endif
constexpr bool kSplitFwd = SPNERF_SPLIT_FWD != 0, kSplitBwd = SPNERF_SPLIT_BWD != 0;

// The step list travels to the kernels as a launch parameter (constant bank): the producer and the
// MMA issuer read one entry per item, and a dependent global load per item would cap the issue rate.
struct StepTable {
  int n;
  int _pad[3];
  MmaStep s[kMaxSteps];
};
// host: the forward (backward = 0) or backward-data (1) step list of a configuration (mlp_pack.cu)
const StepTable* step_table(const SpnerfNetConfig& cfg, int backward);

// The aux operand: per point [1, sun_dir(3), t_emb(<=8), 1, 0, 0, 0].  Multiplying it by a B tile
// that holds [fp16(b), W[:, sun columns], W[:, t columns], b - fp16(b)] folds the bias and the
// per-ray input columns of a layer into its GEMM (models/spnerf.py:351,360 concatenations).
constexpr int kAuxColOne = 0, kAuxColSun = 1, kAuxColT = 4, kAuxColOneLo = 12;
// byte offset of element (row, col) inside a no-swizzle K-major 16-column operand
__host__ __device__ constexpr uint32_t aux_offset(uint32_t row, uint32_t col) {
  return (row >> 3) * 256u + (col >> 3) * 128u + (row & 7u) * 16u + (col & 7u) * 2u;
}

// Offsets (in floats) into the small fp32 parameter block copied by the pack kernel.
struct SmallOffsets {
  int fc_b[8];
  int sigma_b, feats_b;
  int sem0_b, sem2_w, sem2_b;
  int rgb0_b, rgb2_w, rgb2_b;
  int sun0_b, sun0_wsun, sun2_b, sun4_b, sun6_w, sun6_b;
  int beta0_b, beta0_wt, beta2_w, beta2_b;
  int emb;       // (C+1, emb_dim)
  int sky0_w, sky0_b, sky2_w, sky2_b;
  int smallw;    // image of the shared-memory parameter region: rgb2 f4[256] | sem2 f4[256][2] | sun6 | beta2
  int total;
};

// Per-tile activation save area, in 16 KB units (training forward -> backward).  Activations are
// stored row-interleaved (roles::xsave_off): 16-byte chunk (8 columns) x 128 points, 64 columns per
// unit.  The derivative of a sine layer is rebuilt in the backward as +-sqrt(1 - y^2); only its sign
// is saved, one bit per element ("s" regions: uint32 [column / 32][point], 8 KB for 512 columns).
struct SaveMap {
  int inp;         // encoded input (hi)                          1 unit
  int aux;         // [1, sun_dir(3), t_emb(t_dim), 0...]          1 unit (16 columns used)
  int y[8];        // post-activation of trunk layer i            feat / 64 units each (8 at 512)
  int x[8];        // sign bits of cos(sine argument), layer i    1 unit each
  int f;           // feats_from_xyz output                       feat / 64 units
  int sem_x, sem_y, rgb_x, rgb_y, beta_x, beta_y;   // *_y: feat / 128 units, *_x (sign bits): 1 unit (-1 if absent)
  int sun_x[3], sun_y[3];
  int total;
};

// Per-tile gradient save area written by the backward-data kernel, in slabs: every entry is a
// pre-activation gradient tile (fp16, scaled), the A operand of one weight-gradient GEMM.
struct GradMap {
  int G[8];        // trunk layers                         feat / 64 slabs each
  int g_f;         // feats_from_xyz output gradient       feat / 64 slabs
  int G_sem, G_rgb, G_beta;   // hidden layers of the heads feat / 128 slabs each (-1 if absent)
  int G_sun[3];
  int gsmall;      // [g_u(3), g_v, g_sigma_pre, g_beta_pre, 0, 0, g_logit(8), 0...]   1 slab
  int total;
};

struct NetDims {
  int in_dim;      // encoded xyz (3 or 60) + emb_dim
  int in_ksteps;   // K=16 steps covering the columns of the encoded input that live in the input slab (<= 64)
  int n_out;       // 8 (+1 beta) (+C sem)
  int col_beta, col_sem;
};

inline NetDims make_dims(const SpnerfNetConfig& c) {
  NetDims d;
  d.in_dim = (c.mapping ? 60 : 3) + (c.sem ? c.emb_dim : 0);
  d.in_ksteps = ((d.in_dim < 64 ? d.in_dim : 64) + 15) / 16;
  d.col_beta = c.beta ? 8 : -1;
  d.col_sem = c.sem ? 8 + (c.beta ? 1 : 0) : -1;
  d.n_out = 8 + (c.beta ? 1 : 0) + (c.sem ? c.num_sem_classes : 0);
  return d;
}

// Encoded input wider than the 64-column input slab (--mapping with more than 4 semantic classes: 60 + C columns).
// Columns 64.. are label-embedding values; they travel in free columns of the aux operand instead of a second slab
// (no shared memory left for one): per extra column an fp16 high part, the residual, and a second copy of the high
// part, so that the first layer keeps its three-product split precision (hi*W_hi + lo*W_hi + hi*W_lo); the skip
// layer reads the high part only, like the slab columns.  Free aux columns: 13..15, and the transient-embedding
// columns a configuration does not use.  n < 0: does not fit (refused).
struct AuxExtra { int n; int col_hi[4], col_lo[4], col_dup[4]; };
inline AuxExtra make_aux_extra(const SpnerfNetConfig& c) {
  AuxExtra x{};
  const int in_dim = (c.mapping ? 60 : 3) + (c.sem ? c.emb_dim : 0);
  x.n = in_dim > 64 ? in_dim - 64 : 0;
  if (x.n == 0) return x;
  int freec[16], nf = 0;
  for (int k = 13; k < 16; ++k) freec[nf++] = k;
  for (int k = kAuxColT + (c.beta ? c.t_dim : 0); k < kAuxColOneLo; ++k) freec[nf++] = k;
  if (x.n > 4 || 3 * x.n > nf) { x.n = -1; return x; }
  for (int q = 0; q < x.n; ++q) { x.col_hi[q] = freec[q]; x.col_lo[q] = freec[x.n + q]; x.col_dup[q] = freec[2 * x.n + q]; }
  return x;
}

inline SaveMap make_save_map(const SpnerfNetConfig& c) {
  SaveMap m;
  int s = 0;
  const int wide = c.feat / 64, narrow = c.feat / 128;      // units of a trunk-wide / head-wide activation
  m.inp = s++;
  m.aux = s++;
  for (int i = 0; i < 8; ++i) { m.y[i] = s; s += wide; m.x[i] = s; s += 1; }
  m.f = s; s += wide;
  auto four = [&](bool on) { int r = on ? s : -1; if (on) s += narrow; return r; };
  auto one = [&](bool on) { int r = on ? s : -1; if (on) s += 1; return r; };
  m.sem_x = one(c.sem); m.sem_y = four(c.sem);
  m.rgb_x = one(true); m.rgb_y = four(true);
  m.beta_x = one(c.beta); m.beta_y = four(c.beta);
  for (int i = 0; i < 3; ++i) { m.sun_x[i] = one(true); m.sun_y[i] = four(true); }
  m.total = s;
  return m;
}

inline GradMap make_grad_map(const SpnerfNetConfig& c) {
  GradMap m;
  int s = 0;
  const int wide = c.feat / 64, narrow = c.feat / 128;
  for (int i = 0; i < 8; ++i) { m.G[i] = s; s += wide; }
  m.g_f = s; s += wide;
  auto four = [&](bool on) { int r = on ? s : -1; if (on) s += narrow; return r; };
  m.G_sem = four(c.sem); m.G_rgb = four(true); m.G_beta = four(c.beta);
  for (int i = 0; i < 3; ++i) m.G_sun[i] = four(true);
  m.gsmall = s++;
  m.total = s;
  return m;
}

inline SmallOffsets make_small_offsets(const SpnerfNetConfig& c) {
  SmallOffsets o;
  int s = 0;
  auto take = [&](int n) { int r = s; s += (n + 3) & ~3; return r; };   // keep float4 alignment
  // sized for the widest configuration whatever cfg.feat is: the offsets are compile-time facts of the kernels'
  // parameter structs, only the first feat (feat / 2) entries of a block are used by a narrower network
  for (int i = 0; i < 8; ++i) o.fc_b[i] = take(kFeat);
  o.sigma_b = take(1);
  o.feats_b = take(kFeat);
  o.sem0_b = take(kHalf); o.sem2_w = take(8 * kHalf); o.sem2_b = take(8);
  o.rgb0_b = take(kHalf); o.rgb2_w = take(3 * kHalf); o.rgb2_b = take(3);
  o.sun0_b = take(kHalf); o.sun0_wsun = take(3 * kHalf);
  o.sun2_b = take(kHalf); o.sun4_b = take(kHalf); o.sun6_w = take(kHalf); o.sun6_b = take(1);
  o.beta0_b = take(kHalf); o.beta0_wt = take(8 * kHalf); o.beta2_w = take(kHalf); o.beta2_b = take(1);
  o.emb = take(9 * 8);
  o.sky0_w = take(3 * kHalf); o.sky0_b = take(kHalf); o.sky2_w = take(3 * kHalf); o.sky2_b = take(3);
  o.smallw = take(kSmallWFloats);
  o.total = s;
  (void)c;
  return o;
}

}  // namespace net
