// Fused point-network backward, data path: walks the layers of models/spnerf.py:305-369 in reverse
// for a tile of 128 sample points per CTA, gradient tile resident in shared memory (fp16, scaled
// by a power of two chosen from max|dL/d out|), accumulators in tensor memory.
//
// Replaces the autograd backward of SPNeRF.forward (cuBLAS dgrad GEMMs + elementwise cos kernels
// in the reference).  For every layer the pre-activation gradient G = dL/d(sine argument) is
//   * left in shared memory as the A operand of the next (earlier) layer's GEMM, and
//   * streamed to the gradient save area, where the weight-gradient GEMMs (mlp_wgrad.cu) read it.
// Same roles / handshake as mlp_fwd.cu; step order from build_backward() in mlp_pack.cu.
#include <cstdlib>
// store policy of the gradient tiles (mlp_roles.cuh stg16): streaming, like the forward's saves.  Together with
// L1::no_allocate loads of the saved outputs (below): 4.11 -> 4.00 ms in an alternating A/B (each alone -1.2 %), once
// the weight ring was pinned in L2 (SPNERF_W_POLICY); before that neither moved the kernel.
#ifndef SPNERF_STG_MODE_BWD
#define SPNERF_STG_MODE_BWD 1
#endif
#define SPNERF_STG_MODE SPNERF_STG_MODE_BWD
#include "mlp_roles.cuh"

using namespace roles;

namespace {

struct BwdParams {
  const float* g_out; const float* out; const float* rays; const int64_t* labels; const float* t_emb;
  int64_t n_rays; int64_t n_points; int n_samples;
  const uint8_t* blob; StepTable tab;
  const float* small; SmallOffsets so; SaveMap sm; GradMap gm;
  const uint8_t* saves; uint8_t* gsaves;
  const float* absmax; float* scale_out;
  float* g_emb; float* g_small_bias; float* g_t_emb;
  int sem, n_classes, emb_dim, beta, t_dim, n_out, col_beta, col_sem, debug;
  long long* prof;
  int stagger, stagger_groups;      // start delay of cluster c: stagger * (c % groups) / groups cycles
};

// Streaming 16-byte load of a saved activation chunk.  SPNERF_LDG_MODE: 0 = ld.global.nc (allocates an L1
// line per request; with 227 KB of shared memory carved out only ~28 KB of L1 are left to hold the requests in flight),
// 1 = L1::no_allocate (default), 2 = .cs (evict-first: +3 %), 3 = L1::no_allocate + L2::evict_first policy (-0.9 %)
#ifndef SPNERF_LDG_MODE
#define SPNERF_LDG_MODE 1
#endif
__device__ __forceinline__ uint4 ldg16(const uint8_t* p) {
#if SPNERF_LDG_MODE == 0
  return __ldg(reinterpret_cast<const uint4*>(p));
#else
  uint4 v;
#if SPNERF_LDG_MODE == 1
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
#elif SPNERF_LDG_MODE == 2
  asm volatile("ld.global.cs.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
#else
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(pol));
#endif
  return v;
#endif
}
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
  const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 t = __half22float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}

// derivative of a sine layer from its saved output: |cos| = sqrt(1 - y^2), sign from the saved bit
// RELU: the activation was max(x, 0) and the saved bit is x > 0: the derivative is the bit itself
template <bool RELU = false>
__device__ __forceinline__ float dsin(float y, uint32_t sb, int k) {
  if (RELU) return (float)((sb >> k) & 1u);
  float c;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(c) : "f"(__saturatef(fmaf(-y, y, 1.f))));     // one MUFU
  return __uint_as_float(__float_as_uint(c) ^ ((sb >> k) << 31));
}

// One 16-column batch of a saved sine layer for this thread's row: outputs y (fp16) and derivative signs.
// The epilogue reads them straight from global memory (HBM latency), so every batch is requested one
// batch ahead and the first one before the wait for the MMA phase.  (16 columns, not 32: two batches
// in flight next to the accumulator registers must fit the 96-register budget of 640-thread CTAs.)
struct YBatch { uint4 y[2]; uint32_t sb; };
__device__ __forceinline__ YBatch ybatch_load(const uint8_t* ysave, const uint8_t* ssave, int jb, int row) {
  YBatch b;
  b.y[0] = ldg16(ysave + xsave_off(jb, row));
  b.y[1] = ldg16(ysave + xsave_off(jb + 8, row));
  b.sb = __ldg(reinterpret_cast<const uint32_t*>(ssave + sbit_off(jb, row))) >> (jb & 16);
  return b;
}

// A window of kYWin batches per thread is kept in flight (Little: ~64 KB per SM must be outstanding
// to stream a 128 KB activation tile from HBM in a few microseconds; one batch per thread gave 12 GB/s).
// Three batches, not four: at the 96-register budget of 640-thread CTAs the fourth one spilled (120 B stores /
// 272 B loads per thread -> 28 / 36) and the kernel is 1.7 % faster without it; two measure the same as three.
#ifndef SPNERF_YWIN
#define SPNERF_YWIN 3
#endif
constexpr int kYWin = SPNERF_YWIN;
#ifndef SPNERF_BWD_HMUL
#define SPNERF_BWD_HMUL 1
#endif
#ifndef SPNERF_BWD_EARLY_STS
#define SPNERF_BWD_EARLY_STS 1
#endif
#ifndef SPNERF_BWD_NDIRECT
#define SPNERF_BWD_NDIRECT 3
#endif
struct YWindow { YBatch b[kYWin]; };
// NB: batches the column group owns in this layer (a narrow network's head layers have fewer than the window)
template <int NB>
__device__ __forceinline__ void ywin_load(YWindow& w, const uint8_t* ysave, const uint8_t* ssave, int j0, int row) {
#pragma unroll
  for (int i = 0; i < kYWin; ++i)
    if (i < NB) w.b[i] = ybatch_load(ysave, ssave, j0 + 16 * i, row);
}

// The derivative needs nothing from the accumulator, so the batches that sit in the window while the epilogue warps
// wait for an MMA phase are converted in place during that wait: y (fp16) -> dact (fp16: cos x or 30 cos x, signed).
// MUFU, saturate and sign work leave the exposed part of the epilogue; fp16 rounding of the factor (2^-11 relative)
// is below the error of rebuilding it from the fp16 output.
template <int MODE>
__device__ __forceinline__ void ybatch_to_dact(YBatch& b) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    float y[8], d[8];
    unpack8(b.y[c], y);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      d[e] = dsin<MODE == 4>(y[e], b.sb, c * 8 + e);
      if (MODE == 1) d[e] *= 30.f;
    }
    b.y[c] = make_uint4(pack2(d[0], d[1]), pack2(d[2], d[3]), pack2(d[4], d[5]), pack2(d[6], d[7]));
  }
}
template <int MODE, int NB>
__device__ __forceinline__ void ywin_to_dact(YWindow& w) {
#pragma unroll
  for (int i = 0; i < kYWin; ++i)
    if (i < NB) ybatch_to_dact<MODE>(w.b[i]);
}
template <int NB> struct PreBatches { static constexpr int value = NB < kYWin ? NB : kYWin; };

// G[j] = acc[j] * dact(j) for columns [j0, j0 + 16 NB) of a chunk at TMEM address taddr;
// MODE 0: dact = cos(x) rebuilt from (ysave, ssave)     MODE 1: dact = 30 cos(30 x), same     MODE 2: dact = 1
// MODE 4: dact = [x > 0] (ReLU network; the saved outputs are loaded like the sine layers' but only the bits matter)
// result -> fp16 -> shared slab at column dst_col0 + j (copied to the gradient save area afterwards)
// `win`: batches j0 .. j0 + 16 kYWin, loaded by the caller before it waited for the accumulator
// keep: (split phases) the packed fp16 gradients stay in the caller's registers (2 x uint4 per batch) instead of going
//       to shared memory
// PRE: the first PRE batches of the window already hold the derivative (ywin_to_dact)
template <int MODE, int NB, int PRE = 0>
__device__ __forceinline__ void bwd_columns(uint32_t taddr, int j0, const uint8_t* ysave, const uint8_t* ssave,
                                            uint8_t* act, int dst_col0, int row, uint8_t* gsave, YWindow& win, int nb_run = NB,
                                            uint4* keep = nullptr) {
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    if (b < nb_run) {
      const int jb = j0 + 16 * b;
      YBatch cur = win.b[b % kYWin];
      if (MODE != 2 && MODE != 3 && b + kYWin < NB) win.b[b % kYWin] = ybatch_load(ysave, ssave, jb + 16 * kYWin, row);
      uint32_t v[16];
      tmem_ld16(taddr + jb, v);
      tmem_wait_ld();
      if (MODE == 3) {     // timing experiment: the epilogue's arithmetic without its global loads
        cur.y[0] = make_uint4(v[0] & 0x3bff3bffu, v[1] & 0x3bff3bffu, v[2] & 0x3bff3bffu, v[3] & 0x3bff3bffu);
        cur.y[1] = make_uint4(v[4] & 0x3bff3bffu, v[5] & 0x3bff3bffu, v[6] & 0x3bff3bffu, v[7] & 0x3bff3bffu);
        cur.sb = v[8];
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float g[8];
#if SPNERF_BWD_HMUL
        if (MODE != 2 && b < PRE) {      // factor ready in fp16: convert the accumulator pairwise and multiply as half2
          const uint32_t* d2 = reinterpret_cast<const uint32_t*>(&cur.y[c]);
          uint32_t r[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const uint32_t a2 = pack2(__uint_as_float(v[c * 8 + 2 * e]), __uint_as_float(v[c * 8 + 2 * e + 1]));
            const __half2 pr = __hmul2(*reinterpret_cast<const __half2*>(&a2), *reinterpret_cast<const __half2*>(&d2[e]));
            r[e] = *reinterpret_cast<const uint32_t*>(&pr);
          }
          const uint4 gp = make_uint4(r[0], r[1], r[2], r[3]);
          if (keep) keep[2 * b + c] = gp;
          else *reinterpret_cast<uint4*>(act + slab_off(dst_col0 + jb + c * 8, row)) = gp;
          if (gsave) stg16(gsave + xsave_off(jb + c * 8, row), gp);
          continue;
        }
#endif
        if (MODE != 2) {
          float y[8];
          unpack8(cur.y[c], y);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            if (b < PRE) g[e] = __uint_as_float(v[c * 8 + e]) * y[e];
            else {
              const float d = dsin<MODE == 4>(y[e], cur.sb, c * 8 + e);
              g[e] = __uint_as_float(v[c * 8 + e]) * (MODE == 1 ? 30.f * d : d);
            }
          }
        } else {
#pragma unroll
          for (int e = 0; e < 8; ++e) g[e] = __uint_as_float(v[c * 8 + e]);
        }
        const uint4 gp = make_uint4(pack2(g[0], g[1]), pack2(g[2], g[3]), pack2(g[4], g[5]), pack2(g[6], g[7]));
        if (keep) keep[2 * b + c] = gp;
        else *reinterpret_cast<uint4*>(act + slab_off(dst_col0 + jb + c * 8, row)) = gp;
        if (gsave) stg16(gsave + xsave_off(jb + c * 8, row), gp);
      }
    }
  }
}

// Gradient entering a 256-wide hidden layer from its tiny output layer (CUDA cores):
//   G[j] = coefw(j) * cos(x[j])   for j in [j0, j0+ncols),  coefw(j) = sum_c g_c * W2[c][j] from shared memory
// `each(j, G)` lets the caller fold further per-row reductions (t_emb gradient).
template <bool RELU, class CoefW, class Each>
__device__ __forceinline__ void gen_columns(CoefW coefw, int j0, int ncols, const uint8_t* ysave, const uint8_t* ssave,
                                            uint8_t* act, int dst_col0, int row, Each each) {
  // 16-column batches, the next one requested before the current one is used (the saved activations come
  // straight from HBM and nothing else runs on the SM during these phases)
  YBatch nxt = ybatch_load(ysave, ssave, j0, row);
#pragma unroll 1
  for (int jb = j0; jb < j0 + ncols; jb += 16) {
    const YBatch cur = nxt;
    if (jb + 16 < j0 + ncols) nxt = ybatch_load(ysave, ssave, jb + 16, row);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      float y[8], g[8];
      unpack8(cur.y[c], y);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        g[e] = coefw(jb + c * 8 + e) * dsin<RELU>(y[e], cur.sb, c * 8 + e);
        each(jb + c * 8 + e, g[e]);
      }
      *reinterpret_cast<uint4*>(act + slab_off(dst_col0 + jb + c * 8, row)) =
          make_uint4(pack2(g[0], g[1]), pack2(g[2], g[3]), pack2(g[4], g[5]), pack2(g[6], g[7]));
    }
  }
}
struct NoEachG { __device__ __forceinline__ void operator()(int, float) const {} };

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}

// FEAT: trunk width (512 or 256); column-group widths as in mlp_fwd.cu
template <int FEAT, bool RELU>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1) mlp_bwd_kernel(const __grid_constant__ BwdParams p) {
  constexpr int H = FEAT / 2, QW = FEAT / 4, HW = FEAT / 8;
  constexpr int NBQ = QW / 16, NBH = HW / 16;      // 16-column batches per column group: trunk layer / head hidden layer
  constexpr int SPC = QW / 64;                      // activation slabs per column group of a trunk-wide tile
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

  // power-of-two gradient scale keeping fp16 operands in range
  float scale = 1.f;
  {
    const float am = *p.absmax;
    if (am > 0.f && isfinite(am)) {
      int e = (int)floorf(log2f(64.f / am));
      e = e < -60 ? -60 : (e > 60 ? 60 : e);
      scale = exp2f((float)e);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *p.scale_out = scale;
  }
  const float inv_scale = 1.f / scale;

  if (warp >= kProducerWarp) ctl_registers();
  if (warp == kProducerWarp) {
    producer_loop(sh, p.blob, p.tab, n_iters, dbg, prof);
  } else if (warp == kIssuerWarp0 || warp == kIssuerWarp1) {
    if (sh.rank == 0) mma_loop(sh, tmem_base, p.tab, warp - kIssuerWarp0, n_iters, dbg, prof);
    else if (warp == kIssuerWarp0) relay_loop(sh, p.tab, n_iters, dbg);
  } else if (warp < 16) {
    epi_registers();
    const int cg = (warp - kEpiWarp0) >> 2;
    const int row = (warp & 3) * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const float* S = p.small;
    const float4* Wrgb2 = reinterpret_cast<const float4*>(smem + kOffRgb2);
    const float4* Wsem2 = reinterpret_cast<const float4*>(smem + kOffSem2);
    const float* Wsun6 = reinterpret_cast<const float*>(smem + kOffSun6);
    const float* Wbeta2 = reinterpret_cast<const float*>(smem + kOffBeta2);
    const int wide_cols = (dbg & 2) ? 32 : QW;
    EpiSync sync(sh, prof);
    if (lane == 0) mbar_wait(sh.bar_par, 0, 31);
    __syncwarp();
    stagger_start(p.stagger, p.stagger_groups);

    for (int64_t it = 0; it < n_iters; ++it) {
      // this CTA's tile of the pair; an odd tile count leaves rank 1 a phantom tile: it reads tile 0's
      // saves (finite values, multiplied by zero gradients) and stores nothing
      const int64_t tile = 2 * ((blockIdx.x >> 1) + it * (gridDim.x >> 1)) + sh.rank;
      const bool tile_ok = tile < n_tiles;
      const int64_t pt = tile * kTileM + row;
      const bool valid = pt < p.n_points;
      const int64_t ray = valid ? pt / p.n_samples : 0;
      const uint8_t* tsave = p.saves + (size_t)(tile_ok ? tile : 0) * p.sm.total * kSlabBytes;
      uint8_t* tg = tile_ok ? p.gsaves + (size_t)tile * p.gm.total * kSlabBytes : nullptr;
      auto xs = [&](int slab) { return tsave + (size_t)slab * kSlabBytes; };
      auto gs = [&](int slab) -> uint8_t* { return tg ? tg + (size_t)slab * kSlabBytes : nullptr; };

      // ---- head-level gradients of this row (fp32, unscaled) ----
      sync.stamp();
      float g_u[3] = {0.f, 0.f, 0.f}, g_v = 0.f, g_sp = 0.f, g_bp = 0.f, g_lg[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) g_lg[c] = 0.f;
      {
        // The tile's rows of g_out / out (n_out floats per point, contiguous over the tile) are staged through
        // activation slab 7 with coalesced 16-byte loads: a per-thread row read has a 4 n_out byte stride
        // (44 sectors per warp-level load, 22 loads per thread, all four column groups reading the same rows)
        // and cost ~15 k cycles per tile with the tensor pipe idle.  Slabs 4..7 are free here: every thread has
        // passed the barrier of the previous tile's last begin().
        float* stage_go = reinterpret_cast<float*>(act + 7 * kSlabBytes);
        float* stage_o = stage_go + kTileM * 16;
        const int etid = (int)threadIdx.x - kEpiWarp0 * 32;
        const int64_t p0 = tile * kTileM;
        const int nrow = tile_ok ? (int)min((int64_t)kTileM, p.n_points - p0) : 0;
        const int nflt = nrow * p.n_out;
        const float* src_go = p.g_out + p0 * p.n_out;
        const float* src_o = p.out + p0 * p.n_out;
        if (((p0 * p.n_out) & 3) == 0) {                       // 16-byte aligned tile start (always: 128 * n_out floats)
          for (int i = etid * 4; i < nflt; i += kEpiThreads * 4) {
            if (i + 4 <= nflt) {
              *reinterpret_cast<float4*>(stage_go + i) = ldg4(src_go + i);
              *reinterpret_cast<float4*>(stage_o + i) = ldg4(src_o + i);
            } else {
              for (int e = i; e < nflt; ++e) { stage_go[e] = __ldg(src_go + e); stage_o[e] = __ldg(src_o + e); }
            }
          }
        } else {
          for (int i = etid; i < nflt; i += kEpiThreads) { stage_go[i] = __ldg(src_go + i); stage_o[i] = __ldg(src_o + i); }
        }
        epi_bar_sync();
      }
      if (valid) {
        const float* go = reinterpret_cast<const float*>(act + 7 * kSlabBytes) + row * p.n_out;
        const float* o = go + kTileM * 16;
#pragma unroll
        for (int c = 0; c < 3; ++c) {                  // rgb = sigmoid(u)*1.002 - 0.001  (spnerf.py:346-347)
          const float s = (o[c] + 0.001f) / 1.002f;
          g_u[c] = go[c] * 1.002f * s * (1.f - s);
        }
        g_sp = go[3] * (1.f - expf(-o[3]));            // softplus' = 1 - exp(-softplus)
        g_v = go[4] * o[4] * (1.f - o[4]);             // sigmoid'
        if (p.beta) g_bp = go[p.col_beta] * (1.f - expf(-o[p.col_beta]));
        if (p.sem)
          for (int c = 0; c < p.n_classes; ++c) g_lg[c] = go[p.col_sem + c];
      }
      // small-gradient slab [g_u(3), g_v, g_sigma_pre, g_beta_pre, 0, 0, g_logit(8)] (scaled), the B operand
      // of the tiny last-layer weight gradients; bias gradients of those layers are reduced right here
      if (cg == 0 && tile_ok) {
        uint8_t* d = gs(p.gm.gsmall);
        const float sc = scale;
        stg16(d + xsave_off(0, row),
              make_uint4(pack2(g_u[0] * sc, g_u[1] * sc), pack2(g_u[2] * sc, g_v * sc), pack2(g_sp * sc, g_bp * sc), 0u));
        stg16(d + xsave_off(8, row),
              make_uint4(pack2(g_lg[0] * sc, g_lg[1] * sc), pack2(g_lg[2] * sc, g_lg[3] * sc),
                         pack2(g_lg[4] * sc, g_lg[5] * sc), pack2(g_lg[6] * sc, g_lg[7] * sc)));
        float sums[14] = {g_u[0], g_u[1], g_u[2], g_v, g_sp, g_bp, g_lg[0], g_lg[1], g_lg[2], g_lg[3],
                          g_lg[4], g_lg[5], g_lg[6], g_lg[7]};
#pragma unroll
        for (int k = 0; k < 14; ++k) {
          const float t = warp_sum(sums[k]);
          if (lane == 0 && t != 0.f) atomicAdd(p.g_small_bias + k, t);
        }
      }
      // ---- E_in: G_s3 = g_v * W_sun6 * cos(x_s3) -> slabs 0..3 ----
      {
        const float cv = g_v * scale;
        gen_columns<RELU>([&](int j) { return cv * Wsun6[j]; }, cg * HW, HW, xs(p.sm.sun_y[2]), xs(p.sm.sun_x[2]), act, 0, row, NoEachG());
      }
      sync.end(true);
      copy_slabs_out(act, 0, H / 64, gs(p.gm.G_sun[2]));
      // ---- after sun_v_net.4^T: G_s2 ----
      YWindow win;
      ywin_load<NBH>(win, xs(p.sm.sun_y[1]), xs(p.sm.sun_x[1]), cg * HW, row);
      if (!RELU) ywin_to_dact<0, NBH>(win);
      sync.begin();
      bwd_columns<RELU ? 4 : 0, NBH, RELU ? 0 : PreBatches<NBH>::value>(taddr, cg * HW, xs(p.sm.sun_y[1]), xs(p.sm.sun_x[1]), act, 0, row, nullptr, win);
      sync.end(true);
      copy_slabs_out(act, 0, H / 64, gs(p.gm.G_sun[1]));
      // ---- after sun_v_net.2^T: G_s1 -> slabs 0..3 ; albedo hidden G_r1 -> slabs 4..7 ----
      ywin_load<NBH>(win, xs(p.sm.sun_y[0]), xs(p.sm.sun_x[0]), cg * HW, row);
      if (!RELU) ywin_to_dact<0, NBH>(win);
      sync.begin();
      bwd_columns<RELU ? 4 : 0, NBH, RELU ? 0 : PreBatches<NBH>::value>(taddr, cg * HW, xs(p.sm.sun_y[0]), xs(p.sm.sun_x[0]), act, 0, row, nullptr, win);
      {
        const float c0 = g_u[0] * scale, c1 = g_u[1] * scale, c2 = g_u[2] * scale;
        gen_columns<RELU>([&](int j) { const float4 w = Wrgb2[j]; return fmaf(c0, w.x, fmaf(c1, w.y, c2 * w.z)); }, cg * HW, HW,
                    xs(p.sm.rgb_y), xs(p.sm.rgb_x), act, H, row, NoEachG());
      }
      sync.end(true);
      copy_slabs_out(act, 0, H / 64, gs(p.gm.G_sun[0]));
      copy_slabs_out(act, H / 64, H / 64, gs(p.gm.G_rgb));
      if (p.beta) {
        // ---- beta hidden G_b1 -> slabs 0..3 (the sun/albedo GEMMs have retired); d t_emb on the way ----
        sync.begin();
        float tacc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) tacc[e] = 0.f;
        const float cb = g_bp * scale;
        const float* wt = S + p.so.beta0_wt;
        gen_columns<RELU>([&](int j) { return cb * Wbeta2[j]; }, cg * HW, HW, xs(p.sm.beta_y), xs(p.sm.beta_x), act, 0, row,
                    [&](int j, float g) {
#pragma unroll
                      for (int e = 0; e < 8; ++e) tacc[e] = fmaf(g, __ldg(wt + e * H + j), tacc[e]);
                    });
        if (p.g_t_emb && valid)
          for (int e = 0; e < p.t_dim; ++e) atomicAdd(p.g_t_emb + ray * p.t_dim + e, tacc[e] * inv_scale);
        sync.end(true);
        copy_slabs_out(act, 0, H / 64, gs(p.gm.G_beta));
      }
      // ---- g_f (linear) -> slabs 0..7 ----
      sync.begin();
      bwd_columns<2, NBQ>(taddr, cg * QW, nullptr, nullptr, act, 0, row, cg < 2 ? gs(p.gm.g_f) : nullptr, win, wide_cols / 16);
      sync.end(true);
      copy_slabs_out(act, FEAT / 128, FEAT / 128, gs(p.gm.g_f + FEAT / 128));
      // ---- while g_f * W_feats sits in TMEM: semantic hidden G_sem1 -> slabs 0..3, sigma column -> slab 4 ----
      sync.begin();
      if (p.sem) {
        float cf[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) cf[c] = g_lg[c] * scale;
        const bool wide = p.n_classes > 4;
        gen_columns<RELU>([&](int j) {
          const float4 w = Wsem2[j * 2];
          float a = fmaf(cf[0], w.x, fmaf(cf[1], w.y, fmaf(cf[2], w.z, cf[3] * w.w)));
          if (wide) {
            const float4 u = Wsem2[j * 2 + 1];
            a += fmaf(cf[4], u.x, fmaf(cf[5], u.y, fmaf(cf[6], u.z, cf[7] * u.w)));
          }
          return a;
        }, cg * HW, HW, xs(p.sm.sem_y), xs(p.sm.sem_x), act, 0, row, NoEachG());
      }
      if (cg == 3) {
        *reinterpret_cast<uint4*>(act + (H / 64) * kSlabBytes + slab_chunk_offset(row, 0)) =
            make_uint4(pack2(g_sp * scale, 0.f), 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(act + (H / 64) * kSlabBytes + slab_chunk_offset(row, 1)) = make_uint4(0u, 0u, 0u, 0u);
      }
      sync.end(true);
      if (p.sem) copy_slabs_out(act, 0, H / 64, gs(p.gm.G_sem));
      // ---- G_7 = g_h * cos(x_7), then the trunk ----
      float gemb[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) gemb[e] = 0.f;
      auto emb_phase = [&](bool signal) {      // 16-wide product with the embedding columns of a weight
        sync.begin();
        if (cg == 0) {
          uint32_t v[16];
          tmem_ld16(taddr, v);
          tmem_wait_ld();
#pragma unroll
          for (int e = 0; e < 8; ++e) gemb[e] += __uint_as_float(v[e]);
        }
        if (signal) sync.end(true);
      };
      for (int L = 7; L >= 0; --L) {
        // three of the four column groups store their part of the gradient tile from registers, the last one's is
        // copied out of shared memory during the next MMAs (alternating A/B on one box: 2 groups 4.68 ms, 3 groups
        // 4.60, all four 4.72; debug & 256 / 512 select 2 / 1 groups)
        const int ndirect = (dbg & 256) ? 2 : (dbg & 512) ? 1 : SPNERF_BWD_NDIRECT;
        uint8_t* gdirect = cg < ndirect ? gs(p.gm.G[L]) : nullptr;
        const bool more = (L > 0) || p.sem;
        if constexpr (kSplitBwd) {
          // Split phase (net_plan.h): this thread owns H / 4 columns of each accumulator half.  The first half is
          // turned into gradients (kept in registers, stored to the gradient save area) while the tensor pipe still
          // works on the second half; the shared-memory tile -- the A operand the MMAs are reading -- is touched only
          // after the whole phase has retired.
          constexpr int CW = H / 4, NBC = CW / 16;
          if (CW % 64 != 0) gdirect = gs(p.gm.G[L]);      // narrow network: a group's columns are less than a slab, all direct
          uint4 keep[2 * NBC];
          constexpr int PC = RELU ? 0 : PreBatches<NBC>::value;
          ywin_load<NBC>(win, xs(p.sm.y[L]), xs(p.sm.x[L]), cg * CW, row);
          if (!RELU) { if (L > 0) ywin_to_dact<0, NBC>(win); else ywin_to_dact<1, NBC>(win); }
          sync.begin_half();
          if (L > 0) bwd_columns<RELU ? 4 : 0, NBC, PC>(taddr, cg * CW, xs(p.sm.y[L]), xs(p.sm.x[L]), act, 0, row, gdirect, win, NBC, keep);
          else       bwd_columns<RELU ? 4 : 1, NBC, PC>(taddr, cg * CW, xs(p.sm.y[0]), xs(p.sm.x[0]), act, 0, row, gdirect, win, NBC, keep);
          ywin_load<NBC>(win, xs(p.sm.y[L]), xs(p.sm.x[L]), H + cg * CW, row);
#if SPNERF_BWD_EARLY_STS
          // the first half's gradients go to shared memory as soon as the rest of the phase no longer reads those slabs
          // (half of the shared-memory stores leave the exposed epilogue, which is bound by the SM's store path)
          sync.wait_first_half_free();
#pragma unroll
          for (int q = 0; q < 2 * NBC; ++q) *reinterpret_cast<uint4*>(act + slab_off(cg * CW + 8 * q, row)) = keep[q];
#endif
          if (!RELU) { if (L > 0) ywin_to_dact<0, NBC>(win); else ywin_to_dact<1, NBC>(win); }
          sync.begin();
#if !SPNERF_BWD_EARLY_STS
#pragma unroll
          for (int q = 0; q < 2 * NBC; ++q) *reinterpret_cast<uint4*>(act + slab_off(cg * CW + 8 * q, row)) = keep[q];
#endif
          if (L > 0) bwd_columns<RELU ? 4 : 0, NBC, PC>(taddr, H + cg * CW, xs(p.sm.y[L]), xs(p.sm.x[L]), act, 0, row, gdirect, win);
          else       bwd_columns<RELU ? 4 : 1, NBC, PC>(taddr, H + cg * CW, xs(p.sm.y[0]), xs(p.sm.x[0]), act, 0, row, gdirect, win);
          sync.end(more);
          // column groups >= ndirect: their H / 4 columns of each half leave from shared memory during the next MMAs
          if (ndirect < 4) {
            constexpr int SC = CW / 64 > 0 ? CW / 64 : 1;       // slabs per column group and half
            if constexpr (CW % 64 == 0) {
              copy_slabs_out(act, SC * ndirect, SC * (4 - ndirect), gs(p.gm.G[L] + SC * ndirect));
              copy_slabs_out(act, H / 64 + SC * ndirect, SC * (4 - ndirect), gs(p.gm.G[L] + H / 64 + SC * ndirect));
            }
          }
        } else {
#ifdef SPNERF_EXPERIMENTS
          if (!(dbg & (1024 | 4096)))
#endif
          ywin_load<NBQ>(win, xs(p.sm.y[L]), xs(p.sm.x[L]), cg * QW, row);
          constexpr int PQ = RELU ? 0 : PreBatches<NBQ>::value;
#ifdef SPNERF_EXPERIMENTS
          if (!(dbg & (1024 | 4096)))
#endif
          if (!RELU) { if (L > 0) ywin_to_dact<0, NBQ>(win); else ywin_to_dact<1, NBQ>(win); }
          sync.begin();
#ifdef SPNERF_EXPERIMENTS
          if (dbg & 2048) gdirect = nullptr;
          if (dbg & 1024) bwd_columns<3, NBQ>(taddr, cg * QW, xs(p.sm.y[L]), xs(p.sm.x[L]), act, 0, row, gdirect, win, wide_cols / 16);
          else if (dbg & 4096) bwd_columns<2, NBQ>(taddr, cg * QW, xs(p.sm.y[L]), xs(p.sm.x[L]), act, 0, row, gdirect, win, wide_cols / 16);
          else
#endif
          if (L > 0) bwd_columns<RELU ? 4 : 0, NBQ, PQ>(taddr, cg * QW, xs(p.sm.y[L]), xs(p.sm.x[L]), act, 0, row, gdirect, win, wide_cols / 16);
          else       bwd_columns<RELU ? 4 : 1, NBQ, PQ>(taddr, cg * QW, xs(p.sm.y[0]), xs(p.sm.x[0]), act, 0, row, gdirect, win, wide_cols / 16);
          sync.end(more);
#ifdef SPNERF_EXPERIMENTS
          if (!(dbg & 2048))
#endif
          copy_slabs_out(act, SPC * ndirect, SPC * (4 - ndirect), gs(p.gm.G[L] + SPC * ndirect));
        }
        if (p.sem && L == 4) emb_phase(true);
        if (p.sem && L == 0) emb_phase(false);
      }
      // ---- label-embedding gradient: rows of one warp share a ray (hence a label) when n_samples % 32 == 0 ----
      if (p.sem && cg == 0 && p.g_emb) {
        int lab = -1;
        if (valid && p.labels) { const int64_t l = p.labels[ray]; lab = (l < 0 || l >= p.n_classes) ? -1 : (int)l; }   // padding row / out of range: no grad
        const int lab0 = __shfl_sync(0xffffffffu, lab, 0);
        const bool uniform = __all_sync(0xffffffffu, lab == lab0);
        for (int e = 0; e < p.emb_dim; ++e) {
          const float val = gemb[e] * inv_scale;
          if (uniform) {
            const float t = warp_sum(val);
            if (lane == 0 && lab0 >= 0) atomicAdd(p.g_emb + lab0 * p.emb_dim + e, t);
          } else if (lab >= 0) {
            atomicAdd(p.g_emb + lab * p.emb_dim + e, val);
          }
        }
      }
      // the last epilogue of the tile sends no signal; the next tile's E_in does
      tc_fence_before();
      epi_bar_sync();
    }
  }
  teardown(tmem_base);
}

}  // namespace

static long long* g_prof_bwd = nullptr;
extern "C" void spnerf_debug_phase_clocks_bwd(long long* dev_buf256) { g_prof_bwd = dev_buf256; }

extern "C" int spnerf_mlp_bwd_data(const SpnerfMlpBwd* a, void* stream) {
  if (!a || !a->g_out || !a->out || !a->rays || !a->blob || !a->steps || !a->small || !a->saves || !a->grad_saves ||
      !a->g_absmax || !a->scale_out || !a->g_small_bias)
    return SPNERF_ERR_BAD_ARG;
  if (!feat_supported(a->cfg.feat) || a->cfg.layers != 8 || a->cfg.skip_layer != 4) return SPNERF_ERR_UNSUPPORTED;
  if (a->cfg.sem && (!a->labels || !a->g_emb)) return SPNERF_ERR_BAD_ARG;
  if (a->cfg.beta && !a->t_emb) return SPNERF_ERR_BAD_ARG;
  if (a->n_rays <= 0 || a->n_samples < 1) return a->n_rays == 0 ? 0 : SPNERF_ERR_BAD_ARG;
  BwdParams p;
  p.g_out = a->g_out; p.out = a->out; p.rays = a->rays; p.labels = a->labels; p.t_emb = a->t_emb;
  p.n_rays = a->n_rays; p.n_samples = a->n_samples; p.n_points = a->n_rays * a->n_samples;
  p.blob = static_cast<const uint8_t*>(a->blob);
  const StepTable* tab = step_table(a->cfg, 1);
  if (!tab) return SPNERF_ERR_UNSUPPORTED;
  p.tab = *tab;
  p.small = a->small;
  p.so = make_small_offsets(a->cfg); p.sm = make_save_map(a->cfg); p.gm = make_grad_map(a->cfg);
  p.saves = static_cast<const uint8_t*>(a->saves); p.gsaves = static_cast<uint8_t*>(a->grad_saves);
  p.absmax = a->g_absmax; p.scale_out = a->scale_out;
  p.g_emb = a->g_emb; p.g_small_bias = a->g_small_bias; p.g_t_emb = a->g_t_emb;
  const NetDims d = make_dims(a->cfg);
  p.sem = a->cfg.sem; p.n_classes = a->cfg.num_sem_classes; p.emb_dim = a->cfg.emb_dim; p.beta = a->cfg.beta;
  p.t_dim = a->cfg.t_dim; p.n_out = d.n_out; p.col_beta = d.col_beta; p.col_sem = d.col_sem;
  p.debug = a->debug_flags;
  p.prof = g_prof_bwd;
  host_stagger(p.stagger, p.stagger_groups);
  void (*kern)(const BwdParams) = a->cfg.feat == 512 ? (a->cfg.relu ? mlp_bwd_kernel<512, true> : mlp_bwd_kernel<512, false>)
                                                     : (a->cfg.relu ? mlp_bwd_kernel<256, true> : mlp_bwd_kernel<256, false>);
  if (cudaError_t e = sm100::set_max_dynamic_smem(reinterpret_cast<const void*>(kern), kSmemTotal); e != cudaSuccess) return -(int)e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t n_pairs = ((p.n_points + kTileM - 1) / kTileM + 1) / 2;
  const int64_t clusters = sms / 2;
  kern<<<2u * (unsigned)(n_pairs < clusters ? n_pairs : clusters), kThreads, kSmemTotal,
                   static_cast<cudaStream_t>(stream)>>>(p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}

SPNERF_DEFINE_WATCHDOG_GETTER(spnerf_watchdog_code_bwd)
