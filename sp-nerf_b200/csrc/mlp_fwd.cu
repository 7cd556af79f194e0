// Fused point-network forward: positional encoding + label embedding + 8x512 sine trunk + sigma /
// albedo / sun-visibility / (beta) / semantic heads for a tile of 128 sample points per CTA,
// activations resident in shared memory across layers, accumulators in tensor memory.
//
// Replaces models/spnerf.py:305-369 (SPNeRF.forward) as driven by models/spnerf.py:85-109
// (flatten, repeat_interleave of per-ray inputs, chunked calls) and modules/rendering.py:147
// (points = origin + direction * depth).
//
// Roles (384 threads, one CTA per SM, persistent over tiles):
//   warp 0      weight producer: 1-D bulk copies of pre-packed fp16 B tiles into a 2-stage ring
//   warp 1      MMA issuer (one lane): tcgen05.mma, M=128, N<=256, K=16, fp16 x fp16 -> fp32 in TMEM
//   warps 4-11  epilogue: TMEM -> registers, bias + sine / heads, fp16 -> shared memory (next
//               layer's A operand) and, when training, -> the activation save area
// Phases alternate MMA and epilogue (handshake on two mbarriers); the step list built by
// mlp_pack.cu fixes the order on both sides.
#include "mlp_roles.cuh"

using namespace roles;

namespace {

struct FwdParams {
  const float* rays; const float* z; const float* xyz; const float* dir_override;
  const int64_t* labels; const float* t_emb; const float* sky;
  int64_t n_rays; int64_t n_points; int n_samples;
  const uint8_t* blob; const MmaStep* steps; int n_steps;
  const float* small; SmallOffsets so; SaveMap sm;
  float* out; uint8_t* saves;
  int mapping, sem, n_classes, emb_dim, beta, t_dim, in_dim, n_out, col_beta, col_sem;
  int debug;
};


// Process columns [j0, j0+ncols) of one accumulation chunk (whose column 0 sits at TMEM column tcol0)
// for this thread's row:
//   x = acc + bias[j] (+ extra(j)) ; y = ACT(x) ; y -> fp16 -> shared slab(s) at column dst_col0 + j
// MODE 0: y = sin(x)        save x (fp16 argument) and y
// MODE 1: y = sin(30 x)     save cos(30 x) in the x slot and y           (first layer, Siren w0=30)
// MODE 2: y = x             save y only                                   (feats_from_xyz)
// `each(j, y)` is called for every output (head reductions) ; `extra(j)` adds per-row terms.
template <int MODE, bool TO_SMEM, class Extra, class Each>
__device__ __forceinline__ void epi_columns(uint32_t taddr, int tcol0, int j0, int ncols,
                                            const float* __restrict__ bias,
                                            uint8_t* act, int dst_col0, int row, uint8_t* save_x, uint8_t* save_y,
                                            Extra extra, Each each, int skip = 0) {
  if (skip) ncols = 32;
  for (int jb = j0; jb < j0 + ncols; jb += 32) {
    uint32_t v[32];
    tmem_ld32(taddr + tcol0 + jb, v);
    tmem_wait_ld();
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = jb + c * 8;
      const float4 b0 = ldg4(bias + j), b1 = ldg4(bias + j + 4);
      float x[8], y[8], s[8];
      x[0] = __uint_as_float(v[c * 8 + 0]) + b0.x; x[1] = __uint_as_float(v[c * 8 + 1]) + b0.y;
      x[2] = __uint_as_float(v[c * 8 + 2]) + b0.z; x[3] = __uint_as_float(v[c * 8 + 3]) + b0.w;
      x[4] = __uint_as_float(v[c * 8 + 4]) + b1.x; x[5] = __uint_as_float(v[c * 8 + 5]) + b1.y;
      x[6] = __uint_as_float(v[c * 8 + 6]) + b1.z; x[7] = __uint_as_float(v[c * 8 + 7]) + b1.w;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        x[e] += extra(j + e);
        if (MODE == 0) { y[e] = __sinf(x[e]); s[e] = x[e]; }
        else if (MODE == 1) { const float a = 30.f * x[e]; y[e] = __sinf(a); s[e] = __cosf(a); }
        else { y[e] = x[e]; s[e] = 0.f; }
        each(j + e, y[e]);
      }
      const uint4 yp = make_uint4(pack2(y[0], y[1]), pack2(y[2], y[3]), pack2(y[4], y[5]), pack2(y[6], y[7]));
      const int dc = dst_col0 + j;
      const uint32_t off = (uint32_t)(dc >> 6) * kSlabBytes + slab_chunk_offset(row, (dc & 63) >> 3);
      if (TO_SMEM) *reinterpret_cast<uint4*>(act + off) = yp;
      if (save_y) *reinterpret_cast<uint4*>(save_y + off) = yp;
      if (MODE != 2 && save_x) {
        const uint4 sp = make_uint4(pack2(s[0], s[1]), pack2(s[2], s[3]), pack2(s[4], s[5]), pack2(s[6], s[7]));
        *reinterpret_cast<uint4*>(save_x + off) = sp;
      }
    }
  }
}

struct NoExtra { __device__ __forceinline__ float operator()(int) const { return 0.f; } };
struct NoEach { __device__ __forceinline__ void operator()(int, float) const {} };

__device__ __forceinline__ float softplus_ref(float x) { return x > 20.f ? x : log1pf(expf(x)); }   // torch Softplus
__device__ __forceinline__ float sigmoid_ref(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void __launch_bounds__(kThreads, 1) mlp_fwd_kernel(const __grid_constant__ FwdParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const Smem sh = carve(smem);
  uint8_t* act = sh.act;
  float* scratch = reinterpret_cast<float*>(smem + kSlabInpLo * kSlabBytes);   // free after layer 0
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem_base = setup(sh, smem);
  const int64_t n_tiles = (p.n_points + kTileM - 1) / kTileM;

  if (warp == 0) {
    if (lane == 0) producer_loop(sh, p.blob, p.steps, p.n_steps, n_tiles, p.debug);
  } else if (warp == 1) {
    if (lane == 0) mma_loop(sh, tmem_base, p.steps, p.n_steps, n_tiles, p.debug);
  } else if (warp >= kEpiWarp0) {
    const int grp = (warp - kEpiWarp0) >> 2;   // column half handled by this thread
    const int row = (warp & 3) * 32 + lane;    // TMEM lane quarter = warp % 4
    const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const float* S = p.small;
    EpiSync sync(sh);
    auto phase_begin = [&]() { sync.begin(); };
    auto phase_end = [&](bool signal, uint8_t* save_dst, int slab0, int nslabs) {
      sync.end(signal, save_dst, slab0, nslabs);
    };

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t pt = tile * kTileM + row;
      const bool valid = pt < p.n_points;
      const int64_t ray = valid ? pt / p.n_samples : 0;
      uint8_t* tsave = p.saves ? p.saves + (size_t)tile * p.sm.total * kSlabBytes : nullptr;
      auto sv = [&](int slab) -> uint8_t* { return (tsave && slab >= 0) ? tsave + (size_t)slab * kSlabBytes : nullptr; };
      float* orow = p.out + pt * p.n_out;

      // ---- encoded input: [PE(xyz) | label embedding] as fp16 hi + residual ----
      sync.drain_stores();
      float sun[3] = {0.f, 0.f, 0.f};
      {
        float q[3] = {0.f, 0.f, 0.f};
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
          lab = (l == -100) ? p.n_classes : (int)l;       // padding row (models/spnerf.py:310-315)
        }
        const int base = p.mapping ? 60 : 3;
        // this thread fills columns [32*grp, 32*grp+32) of its row
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          float vv[2];
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const int col = grp * 32 + e + u;
            float val = 0.f;
            if (valid) {
              if (col < base) {
                if (p.mapping) {                        // [sin(f x)(3), cos(f x)(3)] per f = 2^k (spnerf.py:32-37)
                  const int k = col / 6, w = col % 6;
                  const float a = __fmul_rn((float)(1 << k), q[w % 3]);
                  val = (w < 3) ? sinf(a) : cosf(a);
                } else val = q[col];
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
        for (int c = 0; c < 4; ++c) {
          const uint32_t off = slab_chunk_offset(row, grp * 4 + c);
          const uint4 h4 = make_uint4(hi[c * 4], hi[c * 4 + 1], hi[c * 4 + 2], hi[c * 4 + 3]);
          *reinterpret_cast<uint4*>(act + kSlabInpHi * kSlabBytes + off) = h4;
          *reinterpret_cast<uint4*>(act + kSlabInpLo * kSlabBytes + off) =
              make_uint4(lo[c * 4], lo[c * 4 + 1], lo[c * 4 + 2], lo[c * 4 + 3]);
          if (tsave) *reinterpret_cast<uint4*>(sv(p.sm.inp) + off) = h4;
        }
        if (tsave && grp == 0) {     // aux slab: [1, sun(3), t_emb, 0...]: operand of the weight-gradient GEMMs
          float a[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) a[e] = 0.f;
          if (valid) {
            a[0] = 1.f; a[1] = sun[0]; a[2] = sun[1]; a[3] = sun[2];
            if (p.beta && p.t_emb)
              for (int e = 0; e < p.t_dim; ++e) a[4 + e] = p.t_emb[ray * p.t_dim + e];
          }
          uint8_t* ax = sv(p.sm.aux);
          *reinterpret_cast<uint4*>(ax + slab_chunk_offset(row, 0)) =
              make_uint4(pack2(a[0], a[1]), pack2(a[2], a[3]), pack2(a[4], a[5]), pack2(a[6], a[7]));
          *reinterpret_cast<uint4*>(ax + slab_chunk_offset(row, 1)) =
              make_uint4(pack2(a[8], a[9]), pack2(a[10], a[11]), pack2(a[12], a[13]), pack2(a[14], a[15]));
          for (int c = 2; c < 8; ++c) *reinterpret_cast<uint4*>(ax + slab_chunk_offset(row, c)) = make_uint4(0, 0, 0, 0);
        }
        if (valid && grp == 0) {     // sky colour is constant along the ray
          orow[5] = p.sky[ray * 3]; orow[6] = p.sky[ray * 3 + 1]; orow[7] = p.sky[ray * 3 + 2];
        }
      }
      phase_end(true, nullptr, 0, 0);

      // ---- trunk layer 0: sin(30 (W0 x + b0))  (spnerf.py:202, Siren w0=30) ----
      phase_begin();
      epi_columns<1, true>(taddr, 0, grp * kHalf, kHalf, S + p.so.fc_b[0], act, 0, row, sv(p.sm.x[0]), nullptr,
                           NoExtra(), NoEach(), p.debug & 2);
      phase_end(true, sv(p.sm.y[0]), 0, 8);
      // ---- trunk layers 1..7 ----
      for (int i = 1; i < 8; ++i) {
        phase_begin();
        epi_columns<0, true>(taddr, 0, grp * kHalf, kHalf, S + p.so.fc_b[i], act, 0, row, sv(p.sm.x[i]), nullptr,
                             NoExtra(), NoEach(), p.debug & 2);
        phase_end(true, sv(p.sm.y[i]), 0, 8);
      }
      // ---- heads on h: semantic hidden (group 0) and sigma (group 1) ----
      phase_begin();
      if (grp == 0) {
        if (p.sem) {
          float lg[8];
#pragma unroll
          for (int c = 0; c < 8; ++c) lg[c] = 0.f;
          const float* w2 = S + p.so.sem2_w;
          const int C = p.n_classes;
          epi_columns<0, false>(taddr, 0, 0, kHalf, S + p.so.sem0_b, act, 0, row, sv(p.sm.sem_x), sv(p.sm.sem_y),
                                NoExtra(), [&](int j, float y) {
#pragma unroll
                                  for (int c = 0; c < 8; ++c)
                                    if (c < C) lg[c] = fmaf(__ldg(w2 + c * kHalf + j), y, lg[c]);
                                });
          if (valid)
            for (int c = 0; c < C; ++c) orow[p.col_sem + c] = lg[c] + S[p.so.sem2_b + c];
        }
      } else {
        uint32_t v[16];
        tmem_ld16(taddr + kHalf, v);
        tmem_wait_ld();
        const float pre = __uint_as_float(v[0]) + __uint_as_float(v[1]) + S[p.so.sigma_b];
        if (valid) orow[3] = softplus_ref(pre);                                   // spnerf.py:333
      }
      phase_end(true, nullptr, 0, 0);
      // ---- feats_from_xyz: linear, overwrites h ----
      phase_begin();
      epi_columns<2, true>(taddr, 0, grp * kHalf, kHalf, S + p.so.feats_b, act, 0, row, nullptr, nullptr, NoExtra(),
                           NoEach(), p.debug & 2);
      phase_end(true, sv(p.sm.f), 0, 8);

      // ---- albedo head (group 0) + beta head or first sun layer (group 1) ----
      auto sun_extra = [&](int j) {
        const float* w = S + p.so.sun0_wsun;
        return fmaf(__ldg(w + j), sun[0], fmaf(__ldg(w + kHalf + j), sun[1], __ldg(w + 2 * kHalf + j) * sun[2]));
      };
      phase_begin();
      if (grp == 0) {
        float c3[3] = {0.f, 0.f, 0.f};
        const float* w2 = S + p.so.rgb2_w;
        epi_columns<0, false>(taddr, 0, 0, kHalf, S + p.so.rgb0_b, act, 0, row, sv(p.sm.rgb_x), sv(p.sm.rgb_y), NoExtra(),
                              [&](int j, float y) {
#pragma unroll
                                for (int c = 0; c < 3; ++c) c3[c] = fmaf(__ldg(w2 + c * kHalf + j), y, c3[c]);
                              });
        if (valid)
          for (int c = 0; c < 3; ++c)                                              // spnerf.py:346-347
            orow[c] = sigmoid_ref(c3[c] + S[p.so.rgb2_b + c]) * 1.002f - 0.001f;
      } else if (p.beta) {
        float tv[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) tv[e] = (valid && e < p.t_dim && p.t_emb) ? p.t_emb[ray * p.t_dim + e] : 0.f;
        float bsum = 0.f;
        const float* wt = S + p.so.beta0_wt;
        const float* w2 = S + p.so.beta2_w;
        epi_columns<0, false>(taddr, kHalf, 0, kHalf, S + p.so.beta0_b, act, 0, row, sv(p.sm.beta_x), sv(p.sm.beta_y),
                              [&](int j) {
                                float a = 0.f;
#pragma unroll
                                for (int e = 0; e < 8; ++e) a = fmaf(__ldg(wt + e * kHalf + j), tv[e], a);
                                return a;
                              },
                              [&](int j, float y) { bsum = fmaf(__ldg(w2 + j), y, bsum); });
        if (valid) orow[p.col_beta] = softplus_ref(bsum + S[p.so.beta2_b]);        // spnerf.py:359-362
      } else {
        // sun layer 0 from accumulator columns 256..511 -> activation columns 0..255
        epi_columns<0, true>(taddr, kHalf, 0, kHalf, S + p.so.sun0_b, act, 0, row, sv(p.sm.sun_x[0]), nullptr,
                             sun_extra, NoEach());
      }
      // (the rgb/beta groups write no shared memory; sun-0's slab store waits for all reads: all MMAs of
      //  this phase retired before phase_begin returned)
      phase_end(true, (p.beta ? nullptr : sv(p.sm.sun_y[0])), 0, 4);
      if (p.beta) {
        phase_begin();
        epi_columns<0, true>(taddr, 0, grp * 128, 128, S + p.so.sun0_b, act, 0, row, sv(p.sm.sun_x[0]), nullptr,
                             sun_extra, NoEach());
        phase_end(true, sv(p.sm.sun_y[0]), 0, 4);
      }
      // ---- sun layer 1 ----
      phase_begin();
      epi_columns<0, true>(taddr, 0, grp * 128, 128, S + p.so.sun2_b, act, 0, row, sv(p.sm.sun_x[1]), nullptr,
                           NoExtra(), NoEach());
      phase_end(true, sv(p.sm.sun_y[1]), 0, 4);
      // ---- sun layer 2 + output unit (256 -> 1, sigmoid): column halves reduced through scratch ----
      phase_begin();
      {
        float part = 0.f;
        const float* w6 = S + p.so.sun6_w;
        epi_columns<0, false>(taddr, 0, grp * 128, 128, S + p.so.sun4_b, act, 0, row, sv(p.sm.sun_x[2]), sv(p.sm.sun_y[2]),
                              NoExtra(), [&](int j, float y) { part = fmaf(__ldg(w6 + j), y, part); });
        if (grp == 1) scratch[row] = part;
        tc_fence_before();
        epi_bar_sync();
        if (grp == 0 && valid) orow[4] = sigmoid_ref(part + scratch[row] + S[p.so.sun6_b]);   // spnerf.py:352
      }
      // no signal: the next tile's input phase releases the MMA warp (it also protects `scratch`)
      epi_bar_sync();
    }
    sync.finish();
  }
  teardown(tmem_base);
}

}  // namespace

extern "C" int spnerf_mlp_fwd(const SpnerfMlpFwd* a, void* stream) {
  if (!a || !a->rays || !a->blob || !a->steps || !a->small || !a->out || !a->sky) return SPNERF_ERR_BAD_ARG;
  if (!a->z && !a->xyz) return SPNERF_ERR_BAD_ARG;
  if (a->n_rays < 0 || a->n_samples < 1) return SPNERF_ERR_BAD_ARG;
  if (a->cfg.feat != 512 || a->cfg.layers != 8 || a->cfg.skip_layer != 4) return SPNERF_ERR_UNSUPPORTED;
  if (a->cfg.beta && !a->t_emb) return SPNERF_ERR_BAD_ARG;
  if (a->n_rays == 0) return 0;
  FwdParams p;
  p.rays = a->rays; p.z = a->z; p.xyz = a->xyz; p.dir_override = a->dir_override;
  p.labels = a->labels; p.t_emb = a->t_emb; p.sky = a->sky;
  p.n_rays = a->n_rays; p.n_samples = a->n_samples; p.n_points = a->n_rays * a->n_samples;
  p.blob = static_cast<const uint8_t*>(a->blob); p.steps = static_cast<const MmaStep*>(a->steps);
  p.n_steps = a->n_steps;
  p.small = a->small; p.so = make_small_offsets(a->cfg); p.sm = make_save_map(a->cfg);
  p.out = a->out; p.saves = static_cast<uint8_t*>(a->saves);
  const NetDims d = make_dims(a->cfg);
  if (d.in_dim > 64) return SPNERF_ERR_UNSUPPORTED;
  p.mapping = a->cfg.mapping; p.sem = a->cfg.sem; p.n_classes = a->cfg.num_sem_classes; p.emb_dim = a->cfg.emb_dim;
  p.beta = a->cfg.beta; p.t_dim = a->cfg.t_dim; p.in_dim = d.in_dim; p.n_out = d.n_out;
  p.col_beta = d.col_beta; p.col_sem = d.col_sem;
  p.debug = a->debug_flags;

  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(mlp_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemTotal);
    if (e != cudaSuccess) return -(int)e;
    attr_set = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t n_tiles = (p.n_points + kTileM - 1) / kTileM;
  const unsigned grid = (unsigned)(n_tiles < sms ? n_tiles : sms);
  mlp_fwd_kernel<<<grid, kThreads, kSmemTotal, static_cast<cudaStream_t>(stream)>>>(p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}

SPNERF_DEFINE_WATCHDOG_GETTER(spnerf_watchdog_code_fwd)
