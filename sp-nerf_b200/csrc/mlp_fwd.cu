// Fused point-network forward: positional encoding + label embedding + 8x512 sine trunk + sigma /
// albedo / sun-visibility / (beta) / semantic heads for a tile of 128 sample points per CTA,
// activations resident in shared memory across layers, accumulators in tensor memory.
//
// Replaces models/spnerf.py:305-369 (SPNeRF.forward) as driven by models/spnerf.py:85-109
// (flatten, repeat_interleave of per-ray inputs, chunked calls) and modules/rendering.py:147
// (points = origin + direction * depth).
//
// Roles (mlp_roles.cuh): producer warp, MMA warp, 16 epilogue warps (TMEM -> registers, sine /
// heads, fp16 -> shared memory = next layer's A operand).  Biases and the per-ray input columns of
// the sun / beta heads are folded into the GEMMs through the 16-column aux operand, so the epilogue
// is activation-only.  When training, post-activations leave through bulk copies of the shared
// slabs, sine arguments through row-interleaved coalesced stores.
// Phases alternate MMA and epilogue (handshake on two mbarriers); the step list built by
// mlp_pack.cu fixes the order on both sides.
#include <cstdlib>
// Store policy of the activation saves (mlp_roles.cuh stg16): streaming (st.global.cs, evict-first).  The 6.8 GB a
// training forward writes are read back milliseconds later by other kernels; with the default write-back policy they
// push the 5 MB weight blob -- re-read by every tile pair -- and each other through L2.  Alternating A/B on one box:
// default 4.19 ms, write-through 3.96, streaming 3.74.
#ifndef SPNERF_STG_MODE_FWD
#define SPNERF_STG_MODE_FWD 1
#endif
#define SPNERF_STG_MODE SPNERF_STG_MODE_FWD
#include "mlp_roles.cuh"

using namespace roles;

namespace {

struct FwdParams {
  const float* rays; const float* z; const float* xyz; const float* dir_override;
  const int64_t* labels; const float* t_emb; const float* sky;
  int64_t n_rays; int64_t n_points; int n_samples;
  const uint8_t* blob; StepTable tab;
  const float* small; SmallOffsets so; SaveMap sm;
  float* out; uint8_t* saves;
  int mapping, sem, n_classes, emb_dim, beta, t_dim, in_dim, n_out, col_beta, col_sem;
  AuxExtra ax;
  int debug;
  long long* prof;
  int stagger, stagger_groups;      // start delay of cluster c: stagger * (c % groups) / groups cycles
};

// 32 accumulator columns [j0, j0+32) of one chunk (column 0 of the chunk at TMEM address `taddr`)
// for this thread's row:  y = ACT(acc) ; y -> fp16 -> shared slab(s) at column dst_col0 + j
// ACT 0: y = sin(x)      ACT 1: y = sin(30 x)  (first layer, Siren w0 = 30)      ACT 2: y = x  (feats_from_xyz)
// ACT 3: y = max(x, 0)   (the ReLU variant of the network, models/spnerf.py:178; the saved bit is x > 0)
// ssave: sign of the derivative cos(argument), one bit per column (the backward rebuilds
//        |cos| = sqrt(1 - y^2)); the sign is the parity of rint(argument / pi), read off the mantissa
//        LSB after adding 1.5 * 2^23
// ysave: copy of y (fp16) in the row-interleaved layout, which the weight-gradient GEMMs read as a
//        no-swizzle MN-major operand (mlp_wgrad.cu); used where y does not pass through shared memory
// `each(j, y)` is called for every output (tiny last layers).
// BITS: compute and save the derivative sign bits (training); the inference instantiation drops the two
//       instructions per element they cost
// W:    accumulator columns per batch, 32 or 16.  The phases that also feed a tiny last layer through `each`
//       use 16: with 32 accumulator values in registers next to the layer's partial sums the training variant
//       spilled inside the loop (semantic head epilogue 17.6 k cycles against 5.8 k without the saves)
// keep: (split phases) the packed fp16 outputs stay in the caller's registers instead of going to shared memory
template <int ACT, bool TO_SMEM, bool BITS, int W, class Each>
__device__ __forceinline__ void epi_batch(uint32_t taddr, int j0, uint8_t* act, int dst_col0, int row, uint8_t* ssave,
                                          uint8_t* ysave, Each each, uint4* keep = nullptr) {
  uint32_t v[W];
  if (W == 32) tmem_ld32(taddr + j0, reinterpret_cast<uint32_t(&)[32]>(v));
  else tmem_ld16(taddr + j0, reinterpret_cast<uint32_t(&)[16]>(v));
  tmem_wait_ld();
  uint32_t sb = 0;
#pragma unroll
  for (int c = 0; c < W / 8; ++c) {
    const int j = j0 + c * 8;
    float y[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float x = __uint_as_float(v[c * 8 + e]);
      if (ACT == 2) y[e] = x;
      else if (ACT == 3) {
        y[e] = fmaxf(x, 0.f);
        if (BITS) sb = __funnelshift_r(sb, x > 0.f ? 1u : 0u, 1);
      } else {
        const float a = (ACT == 1) ? 30.f * x : x;
        y[e] = __sinf(a);
        if (BITS) sb = __funnelshift_r(sb, __float_as_uint(fmaf(a, 0.318309886f, 12582912.f)), 1);   // bit (c*8+e) <- parity
      }
      each(j + e, y[e]);
    }
    const uint4 yp = make_uint4(pack2(y[0], y[1]), pack2(y[2], y[3]), pack2(y[4], y[5]), pack2(y[6], y[7]));
    if (TO_SMEM) *reinterpret_cast<uint4*>(act + slab_off(dst_col0 + j, row)) = yp;
    if (keep) keep[c] = yp;
    if (ysave) stg16(ysave + xsave_off(j, row), yp);
  }
  if (BITS && ACT != 2) {
    if (W == 32) stg4(ssave + sbit_off(j0, row), sb);
    else *reinterpret_cast<uint16_t*>(ssave + sbit_off(j0, row) + ((j0 & 16) >> 3)) = (uint16_t)(sb >> 16);   // low half: columns 0..15
  }
}

template <int ACT, bool TO_SMEM, int W = 32, class Each>
__device__ __forceinline__ void epi_cols(uint32_t taddr, int j0, int ncols, uint8_t* act, int dst_col0, int row,
                                         uint8_t* xsave, uint8_t* ysave, Each each) {
  if (ACT != 2 && xsave) {
#pragma unroll 1
    for (int jb = j0; jb < j0 + ncols; jb += W) epi_batch<ACT, TO_SMEM, true, W>(taddr, jb, act, dst_col0, row, xsave, ysave, each);
  } else {
#pragma unroll 1
    for (int jb = j0; jb < j0 + ncols; jb += W) epi_batch<ACT, TO_SMEM, false, W>(taddr, jb, act, dst_col0, row, nullptr, ysave, each);
  }
}

struct NoEach { __device__ __forceinline__ void operator()(int, float) const {} };

// Epilogue of a layer that fills the whole accumulator (2 H columns: trunk layers, feats_from_xyz), including the waits
// for the MMA phase.  Split phases (net_plan.h): the thread owns H / 4 columns of each accumulator half.  The first
// half is complete when issuer 0 signals bar_half; it is converted (and saved to global memory) while the tensor pipe
// still works on the second half, but its shared-memory image -- the next layer's A operand, in place of the one the
// MMAs are still reading -- is written only after the whole phase has retired.
template <int ACT, int H, class Sync>
__device__ __forceinline__ void epi_full_layer(Sync& sync, uint32_t taddr, int cg, uint8_t* act, int row, uint8_t* xsave,
                                               uint8_t* ysave, int dbg) {
  if constexpr (kSplitFwd) {
    constexpr int CW = H / 4, NB = CW / 32;       // columns per thread and half; 32-column batches
    uint4 keep[NB][4];
    sync.begin_half();
    const bool bits = ACT != 2 && xsave;
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      if (bits) epi_batch<ACT, false, true, 32>(taddr, cg * CW + 32 * b, act, 0, row, xsave, ysave, NoEach(), keep[b]);
      else epi_batch<ACT, false, false, 32>(taddr, cg * CW + 32 * b, act, 0, row, nullptr, ysave, NoEach(), keep[b]);
    }
    sync.begin();
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
      for (int c = 0; c < 4; ++c) *reinterpret_cast<uint4*>(act + slab_off(cg * CW + 32 * b + 8 * c, row)) = keep[b][c];
    epi_cols<ACT, true>(taddr, H + cg * CW, (dbg & 2) ? 32 : CW, act, 0, row, xsave, ysave, NoEach());
  } else {
    sync.begin();
    epi_cols<ACT, true>(taddr, cg * (H / 2), (dbg & 2) ? 32 : H / 2, act, 0, row, xsave, ysave, NoEach());
  }
}

// sin / cos of the positional encoding (models/spnerf.py:32-37), accurate to ~1 ulp over the whole float range like
// sinf / cosf, but without their Payne-Hanek path (a table in local memory: 160 LDL / STL in this kernel's SASS).
// |a| <= 105615: three-constant Cody-Waite reduction by pi/2 in fp32 FMAs and the cephes minimax polynomials on
// [-pi/4, pi/4].  Beyond (scene-normalised coordinates times 2^9 never get there): reduction in double.
__device__ __forceinline__ float pe_sincos(float a, int quadrant_shift) {
  float r;
  int q;
  if (fabsf(a) > 105615.0f) {
    const double t = (double)a;
    const double k = rint(t * 0.6366197723675814);
    double rr = fma(-k, 1.5707963267948966, t);
    rr = fma(-k, 6.123233995736766e-17, rr);
    r = (float)rr;
    q = (int)fmod(k, 4.0);
  } else {
    const float j = rintf(a * 0.636619772f);
    r = fmaf(j, -1.57079601e+00f, a);
    r = fmaf(j, -3.13916473e-07f, r);
    r = fmaf(j, -5.39030253e-15f, r);
    q = (int)j;
  }
  q += quadrant_shift;      // cos(a) = sin(a + pi/2)
  const float s = r * r;
  float v;
  if (q & 1) {
    v = fmaf(s, 2.443315711809948e-5f, -1.388731625493765e-3f);
    v = fmaf(v, s, 4.166664568298827e-2f);
    v = fmaf(v, s, -0.5f);
    v = fmaf(v, s, 1.0f);
  } else {
    v = fmaf(s, -1.9515295891e-4f, 8.3321608736e-3f);
    v = fmaf(v, s, -1.6666654611e-1f);
    v = fmaf(v * s, r, r);
  }
  return (q & 2) ? -v : v;
}

__device__ __forceinline__ float softplus_ref(float x) { return x > 20.f ? x : log1pf(expf(x)); }   // torch Softplus
__device__ __forceinline__ float sigmoid_ref(float x) { return 1.f / (1.f + expf(-x)); }

// FEAT: trunk width (512 or 256).  The tile geometry follows from it: a column group of the epilogue owns FEAT / 4
// accumulator columns of a trunk layer and FEAT / 8 of a head's hidden layer (FEAT / 2 wide).
template <int FEAT, bool RELU>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) mlp_fwd_kernel(const __grid_constant__ FwdParams p) {
  constexpr int H = FEAT / 2, QW = FEAT / 4, HW = FEAT / 8;
  constexpr int A0 = RELU ? 3 : 1, AH = RELU ? 3 : 0;      // activation of the first trunk layer / of every other hidden layer
  extern __shared__ __align__(1024) uint8_t smem[];
  // timing-experiment toggles and the phase clock log exist only in SPNERF_EXPERIMENTS builds (tools/build_variant.sh)
#ifdef SPNERF_EXPERIMENTS
  const int dbg = p.debug;
  long long* const prof = p.prof;
#else
  constexpr int dbg = 0;
  constexpr long long* prof = nullptr;
#endif
  const Smem sh = carve(smem);
  uint8_t* act = sh.act;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem_base = setup(sh, smem, p.small + p.so.smallw, dbg);
  const int64_t n_tiles = (p.n_points + kTileM - 1) / kTileM;
  const int64_t n_iters = my_pairs((n_tiles + 1) / 2);     // tile pairs of this cluster

  if (warp >= kProducerWarp) ctl_registers();
  if (warp == kProducerWarp) {
    producer_loop(sh, p.blob, p.tab, n_iters, dbg, prof);
  } else if (warp == kIssuerWarp0 || warp == kIssuerWarp1) {
    if (sh.rank == 0) mma_loop(sh, tmem_base, p.tab, warp - kIssuerWarp0, n_iters, dbg, prof);
    else if (warp == kIssuerWarp0) relay_loop(sh, p.tab, n_iters, dbg);
  } else if (warp < 16) {
    epi_registers();
    const int cg = (warp - kEpiWarp0) >> 2;    // column group of this warp
    const int row = (warp & 3) * 32 + lane;    // TMEM lane quarter = warp % 4
    const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const float* S = p.small;
    // shared-memory resident tiny last layers; scratch = input slab, dead once the skip layer has run
    const float4* Wrgb2 = reinterpret_cast<const float4*>(smem + kOffRgb2);
    const float4* Wsem2 = reinterpret_cast<const float4*>(smem + kOffSem2);
    const float* Wsun6 = reinterpret_cast<const float*>(smem + kOffSun6);
    const float* Wbeta2 = reinterpret_cast<const float*>(smem + kOffBeta2);
    float* scratch = reinterpret_cast<float*>(smem + kSlabInpHi * kSlabBytes);
    EpiSync sync(sh, prof);
    if (lane == 0) mbar_wait(sh.bar_par, 0, 31);
    __syncwarp();
    stagger_start(p.stagger, p.stagger_groups);

    for (int64_t it = 0; it < n_iters; ++it) {
      // this CTA's tile of the pair; an odd tile count leaves rank 1 a phantom tile (no rows, no stores)
      const int64_t tile = 2 * ((blockIdx.x >> 1) + it * (gridDim.x >> 1)) + sh.rank;
      const int64_t pt = tile * kTileM + row;
      const bool valid = pt < p.n_points;
      const int64_t ray = valid ? pt / p.n_samples : 0;
      uint8_t* tsave = (p.saves && tile < n_tiles) ? p.saves + (size_t)tile * p.sm.total * kSlabBytes : nullptr;
      auto sv = [&](int slab) -> uint8_t* { return (tsave && slab >= 0) ? tsave + (size_t)slab * kSlabBytes : nullptr; };
      float* orow = p.out + pt * p.n_out;

      // ---- encoded input [PE(xyz) | label embedding] as fp16 hi + residual, and the aux operand ----
      sync.stamp();
      {
        float q[3] = {0.f, 0.f, 0.f}, sun[3] = {0.f, 0.f, 0.f};
        if (valid) {
          const float* r = p.rays + ray * 11;
          sun[0] = r[8]; sun[1] = r[9]; sun[2] = r[10];
          if (p.xyz) { q[0] = p.xyz[pt * 3]; q[1] = p.xyz[pt * 3 + 1]; q[2] = p.xyz[pt * 3 + 2]; }
          else {
            const float zz = p.z[pt];
            const float* dd = p.dir_override ? p.dir_override + ray * 3 : r + 3;
            // separate multiply and add, as torch evaluates o + d * z (modules/rendering.py:147)
            q[0] = __fadd_rn(r[0], __fmul_rn(dd[0], zz));
            q[1] = __fadd_rn(r[1], __fmul_rn(dd[1], zz));
            q[2] = __fadd_rn(r[2], __fmul_rn(dd[2], zz));
          }
        }
        int lab = -1;
        if (p.sem && valid && p.labels) {
          const int64_t l = p.labels[ray];
          // -100 -> padding row (models/spnerf.py:310-315).  Any other label outside [0, C) would index past the
          // embedding block (the reference's nn.Embedding raises): it is treated like the padding row
          lab = (l < 0 || l >= p.n_classes) ? p.n_classes : (int)l;
        }
        const int base = p.mapping ? 60 : 3;
        // this thread fills columns [16*cg, 16*cg+16) of its row
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
          float vv[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int col = cg * 16 + e + u;
            float val = 0.f;
            if (valid) {
              if (col < base) {
                if (p.mapping) {                        // [sin(f x)(3), cos(f x)(3)] per f = 2^k (spnerf.py:32-37)
                  const int k = col / 6, w = col % 6;
                  const int w3 = w % 3;
                  const float a = __fmul_rn((float)(1 << k), w3 == 0 ? q[0] : w3 == 1 ? q[1] : q[2]);
                  val = pe_sincos(a, w < 3 ? 0 : 1);
                } else val = col == 0 ? q[0] : col == 1 ? q[1] : q[2];
              } else if (col < p.in_dim && lab >= 0) {
                val = S[p.so.emb + lab * p.emb_dim + (col - base)];
              }
            }
            vv[u] = val;
          }
          const __half2 h = __floats2half2_rn(vv[0], vv[1]);
          const float2 hf = __half22float2(h);
          hi[e >> 1] = *reinterpret_cast<const uint32_t*>(&h);
          lo[e >> 1] = pack2(vv[0] - hf.x, vv[1] - hf.y);
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const uint32_t off = slab_chunk_offset(row, cg * 2 + c);
          *reinterpret_cast<uint4*>(act + kSlabInpHi * kSlabBytes + off) =
              make_uint4(hi[c * 4], hi[c * 4 + 1], hi[c * 4 + 2], hi[c * 4 + 3]);
          *reinterpret_cast<uint4*>(act + kSlabInpLo * kSlabBytes + off) =
              make_uint4(lo[c * 4], lo[c * 4 + 1], lo[c * 4 + 2], lo[c * 4 + 3]);
          if (tsave) stg16(sv(p.sm.inp) + xsave_off(cg * 16 + c * 8, row),
                           make_uint4(hi[c * 4], hi[c * 4 + 1], hi[c * 4 + 2], hi[c * 4 + 3]));
        }
        if (cg == 1 || (tsave && cg == 2)) {     // aux: [1, sun(3), t_emb, 1(lo), 0...]
          float a[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) a[e] = 0.f;
          if (valid) {
            a[kAuxColOne] = 1.f; a[kAuxColSun] = sun[0]; a[kAuxColSun + 1] = sun[1]; a[kAuxColSun + 2] = sun[2];
            // (every index into a[] is a compile-time constant so that the array stays in registers)
            if (p.beta && p.t_emb) {
#pragma unroll
              for (int e = 0; e < 8; ++e)
                if (e < p.t_dim) a[kAuxColT + e] = p.t_emb[ray * p.t_dim + e];
            }
            // encoded-input columns 64.. (label embedding): high part, residual, high part again (net_plan.h AuxExtra)
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (q < p.ax.n) {
                const float val = lab >= 0 ? S[p.so.emb + lab * p.emb_dim + (64 + q - base)] : 0.f;
                const float hi_f = __half2float(__float2half_rn(val));
#pragma unroll
                for (int e = kAuxColT; e < 16; ++e) {
                  if (e == p.ax.col_hi[q] || e == p.ax.col_dup[q]) a[e] = hi_f;
                  else if (e == p.ax.col_lo[q]) a[e] = val - hi_f;
                }
              }
            }
          }
          if (cg == 1) {                         // operand of this tile's aux steps
            if (valid) a[kAuxColOneLo] = 1.f;
            *reinterpret_cast<uint4*>(sh.aux + aux_offset(row, 0)) =
                make_uint4(pack2(a[0], a[1]), pack2(a[2], a[3]), pack2(a[4], a[5]), pack2(a[6], a[7]));
            *reinterpret_cast<uint4*>(sh.aux + aux_offset(row, 8)) =
                make_uint4(pack2(a[8], a[9]), pack2(a[10], a[11]), pack2(a[12], a[13]), pack2(a[14], a[15]));
          } else {                               // saved copy: operand of the weight-gradient GEMMs (16 columns)
            uint8_t* ax = sv(p.sm.aux);
            stg16(ax + xsave_off(0, row),
                  make_uint4(pack2(a[0], a[1]), pack2(a[2], a[3]), pack2(a[4], a[5]), pack2(a[6], a[7])));
            stg16(ax + xsave_off(8, row),
                  make_uint4(pack2(a[8], a[9]), pack2(a[10], a[11]), pack2(a[12], a[13]), pack2(a[14], a[15])));
          }
        }
        if (valid && cg == 3) {     // sky colour is constant along the ray
          orow[5] = p.sky[ray * 3]; orow[6] = p.sky[ray * 3 + 1]; orow[7] = p.sky[ray * 3 + 2];
        }
      }
      sync.end(true);

      // ---- trunk: layer 0 = sin(30 (W0 x + b0)) (spnerf.py:202, Siren w0 = 30), layers 1..7 = sin(W h + b).  The saved
      // copy of every activation tile leaves from registers during the epilogue (copying part of it out of shared
      // memory during the next MMA phase, as the backward does, made the forward 2 % slower: the copy competes with
      // the MMAs for shared-memory bandwidth) ----
      epi_full_layer<A0, H>(sync, taddr, cg, act, row, sv(p.sm.x[0]), sv(p.sm.y[0]), dbg);
      sync.end(true);
      for (int i = 1; i < 8; ++i) {
        epi_full_layer<AH, H>(sync, taddr, cg, act, row, sv(p.sm.x[i]), sv(p.sm.y[i]), dbg);
        sync.end(true);
      }
      // ---- heads on h: semantic hidden (accumulator columns 0..255) and sigma (256, 257) ----
      sync.begin();
      if (cg == 3) {
        uint32_t v[16];
        tmem_ld16(taddr + H, v);
        tmem_wait_ld();
        const float pre = __uint_as_float(v[0]) + __uint_as_float(v[1]);
        if (valid) orow[3] = softplus_ref(pre);                                   // spnerf.py:333
      }
      if (p.sem) {
        float lg[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) lg[c] = 0.f;
        const bool wide = p.n_classes > 4;
        epi_cols<AH, false, 16>(taddr, cg * HW, HW, act, 0, row, (dbg & 32) ? nullptr : sv(p.sm.sem_x),
                           (dbg & 16) ? nullptr : sv(p.sm.sem_y), [&](int j, float y) {
          const float4 w = Wsem2[j * 2];
          lg[0] = fmaf(w.x, y, lg[0]); lg[1] = fmaf(w.y, y, lg[1]); lg[2] = fmaf(w.z, y, lg[2]); lg[3] = fmaf(w.w, y, lg[3]);
          if (wide) {
            const float4 u = Wsem2[j * 2 + 1];
            lg[4] = fmaf(u.x, y, lg[4]); lg[5] = fmaf(u.y, y, lg[5]); lg[6] = fmaf(u.z, y, lg[6]); lg[7] = fmaf(u.w, y, lg[7]);
          }
        });
        reduce_groups<8>(scratch, lg, cg, row);
        if (cg == 0 && valid)
#pragma unroll
          for (int c = 0; c < 8; ++c)      // static indices keep lg[] in registers
            if (c < p.n_classes) orow[p.col_sem + c] = lg[c] + S[p.so.sem2_b + c];   // spnerf.py:365-367
      }
      sync.end(true);
      // ---- feats_from_xyz: linear, overwrites h ----
      epi_full_layer<2, H>(sync, taddr, cg, act, row, nullptr, sv(p.sm.f), dbg);
      sync.end(true);

      // ---- albedo hidden layer (columns 0..255) + first sun layer or beta hidden layer (256..511) ----
      sync.begin();
      {
        float c3[3] = {0.f, 0.f, 0.f};
        auto rgb_each = [&](int j, float y) {
          const float4 w = Wrgb2[j];
          c3[0] = fmaf(w.x, y, c3[0]); c3[1] = fmaf(w.y, y, c3[1]); c3[2] = fmaf(w.z, y, c3[2]);
        };
        if (!p.beta) {
          // every input of this phase has been consumed: the sun activations (next layer's operand) go to slabs
          // 0..3, the albedo activations to slabs 4..7 (only read back by the debug & 128 copy-out variant)
          epi_cols<AH, true, 16>(taddr, cg * HW, HW, act, H, row, sv(p.sm.rgb_x), sv(p.sm.rgb_y), rgb_each);
          epi_cols<AH, true>(taddr + H, cg * HW, HW, act, 0, row, sv(p.sm.sun_x[0]), sv(p.sm.sun_y[0]),
                            NoEach());
          reduce_groups<3>(scratch, c3, cg, row);
        } else {
          // feats stay live for the sun layer of the next phase: nothing may be written to the slabs
          float bsum[1] = {0.f};
          epi_cols<AH, false, 16>(taddr, cg * HW, HW, act, 0, row, sv(p.sm.rgb_x), sv(p.sm.rgb_y), rgb_each);
          epi_cols<AH, false, 16>(taddr + H, cg * HW, HW, act, 0, row, sv(p.sm.beta_x), sv(p.sm.beta_y),
                             [&](int j, float y) { bsum[0] = fmaf(Wbeta2[j], y, bsum[0]); });
          float r4[4] = {c3[0], c3[1], c3[2], bsum[0]};
          reduce_groups<4>(scratch, r4, cg, row);
          c3[0] = r4[0]; c3[1] = r4[1]; c3[2] = r4[2];
          if (cg == 0 && valid) orow[p.col_beta] = softplus_ref(r4[3] + S[p.so.beta2_b]);        // spnerf.py:359-362
        }
        if (cg == 0 && valid)
          for (int c = 0; c < 3; ++c)                                              // spnerf.py:346-347
            orow[c] = sigmoid_ref(c3[c] + S[p.so.rgb2_b + c]) * 1.002f - 0.001f;
      }
      sync.end(true);
      if (p.beta) {
        sync.begin();
        epi_cols<AH, true>(taddr, cg * HW, HW, act, 0, row, sv(p.sm.sun_x[0]), sv(p.sm.sun_y[0]), NoEach());
        sync.end(true);
      }
      // ---- sun layer 1 ----
      sync.begin();
      epi_cols<AH, true>(taddr, cg * HW, HW, act, 0, row, sv(p.sm.sun_x[1]), sv(p.sm.sun_y[1]), NoEach());
      sync.end(true);
      // ---- sun layer 2 + output unit (256 -> 1, sigmoid) ----
      sync.begin();
      {
        float part[1] = {0.f};
        epi_cols<AH, false, 16>(taddr, cg * HW, HW, act, 0, row, sv(p.sm.sun_x[2]), sv(p.sm.sun_y[2]),
                           [&](int j, float y) { part[0] = fmaf(Wsun6[j], y, part[0]); });
        reduce_groups<1>(scratch, part, cg, row);
        if (cg == 0 && valid) orow[4] = sigmoid_ref(part[0] + S[p.so.sun6_b]);   // spnerf.py:352
      }
      // no signal: the next tile's input phase releases the MMA warp
      sync.end(false);
    }
  }
  teardown(tmem_base);
}

}  // namespace

static long long* g_prof_fwd = nullptr;
extern "C" void spnerf_debug_phase_clocks_fwd(long long* dev_buf256) { g_prof_fwd = dev_buf256; }

extern "C" int spnerf_mlp_fwd(const SpnerfMlpFwd* a, void* stream) {
  if (!a || !a->rays || !a->blob || !a->steps || !a->small || !a->out || !a->sky) return SPNERF_ERR_BAD_ARG;
  if (!a->z && !a->xyz) return SPNERF_ERR_BAD_ARG;
  if (a->n_rays < 0 || a->n_samples < 1) return SPNERF_ERR_BAD_ARG;
  if (!feat_supported(a->cfg.feat) || a->cfg.layers != 8 || a->cfg.skip_layer != 4) return SPNERF_ERR_UNSUPPORTED;
  if (a->cfg.beta && !a->t_emb) return SPNERF_ERR_BAD_ARG;
  if (a->n_rays == 0) return 0;
  FwdParams p;
  p.rays = a->rays; p.z = a->z; p.xyz = a->xyz; p.dir_override = a->dir_override;
  p.labels = a->labels; p.t_emb = a->t_emb; p.sky = a->sky;
  p.n_rays = a->n_rays; p.n_samples = a->n_samples; p.n_points = a->n_rays * a->n_samples;
  p.blob = static_cast<const uint8_t*>(a->blob);
  const StepTable* tab = step_table(a->cfg, 0);
  if (!tab) return SPNERF_ERR_UNSUPPORTED;
  p.tab = *tab;
  p.small = a->small; p.so = make_small_offsets(a->cfg); p.sm = make_save_map(a->cfg);
  p.out = a->out; p.saves = static_cast<uint8_t*>(a->saves);
  const NetDims d = make_dims(a->cfg);
  const AuxExtra ax = make_aux_extra(a->cfg);
  if (ax.n < 0) return SPNERF_ERR_UNSUPPORTED;
  p.ax = ax;
  p.mapping = a->cfg.mapping; p.sem = a->cfg.sem; p.n_classes = a->cfg.num_sem_classes; p.emb_dim = a->cfg.emb_dim;
  p.beta = a->cfg.beta; p.t_dim = a->cfg.t_dim; p.in_dim = d.in_dim; p.n_out = d.n_out;
  p.col_beta = d.col_beta; p.col_sem = d.col_sem;
  p.debug = a->debug_flags;
  p.prof = g_prof_fwd;
  host_stagger(p.stagger, p.stagger_groups);

  void (*kern)(const FwdParams) = a->cfg.feat == 512 ? (a->cfg.relu ? mlp_fwd_kernel<512, true> : mlp_fwd_kernel<512, false>)
                                                     : (a->cfg.relu ? mlp_fwd_kernel<256, true> : mlp_fwd_kernel<256, false>);
  if (cudaError_t e = sm100::set_max_dynamic_smem(reinterpret_cast<const void*>(kern), kSmemTotal); e != cudaSuccess) return -(int)e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t n_pairs = ((p.n_points + kTileM - 1) / kTileM + 1) / 2;
  int64_t clusters = sms / 2;
#ifdef SPNERF_EXPERIMENTS
  if (const char* e = getenv("SPNERF_MAX_CLUSTERS")) { const int v = atoi(e); if (v > 0 && v < clusters) clusters = v; }
#endif
  const unsigned grid = 2u * (unsigned)(n_pairs < clusters ? n_pairs : clusters);
  kern<<<grid, kThreads, kSmemTotal, static_cast<cudaStream_t>(stream)>>>(p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}

SPNERF_DEFINE_WATCHDOG_GETTER(spnerf_watchdog_code_fwd)
