// Weight / bias gradients of the point network as tensor-core GEMMs over the point dimension.
//
//   dW[m][n] = sum_points G[point][m] * Y[point][n]
// G (pre-activation gradient tiles, from mlp_bwd.cu) and Y (the forward's saved layer inputs) are
// both stored per 128-point tile as 16-byte chunks (8 features) x 128 points, which is exactly the
// no-swizzle MN-major operand form of tcgen05.mma: no transposition, the tiles are bulk-copied to
// shared memory and multiplied as is.
// Bias gradients fall out of the same GEMMs through the constant-one column of the "aux" slab,
// the sun-direction / transient-embedding input columns through its other columns.
//
// Replaces the wgrad half of autograd through models/spnerf.py:305-369.
//
// Schedule: GEMMs are processed one after another by the whole grid.  Within a GEMM, CTA c takes
// output tile (c mod tiles) and point slice (c div tiles), so the CTAs that share a point slice
// run side by side and hit each other's operands in L2.  Partial tiles go to a workspace; a
// reduce kernel sums the slices, un-scales and scatters into the parameter-shaped gradients.
#include <vector>
#include "sm100.cuh"
#include "net_plan.h"

using namespace sm100;
using namespace net;

namespace {

constexpr int kMaxBSlabs = 5;                  // B slabs per output tile (N <= 320)
constexpr int kStageSlabs = 2 + kMaxBSlabs;    // A: 2 slabs (M = 128)
constexpr int kStageBytes = kStageSlabs * kSlabBytes;
constexpr int kStages = 2;
constexpr int kSmemW = kStages * kStageBytes + 256;
constexpr int kWThreads = 192;                 // warp 0 producer, warp 1 MMA, warps 2-5 epilogue
constexpr int kMaxTiles = 160;
constexpr int kMaxSegs = 400;

struct SlabRef { int16_t from_grads; int16_t slab; };   // slab index inside the per-tile (grad) save area

struct OutTile {                // one 128 x (64*nb) accumulator tile of one GEMM
  SlabRef a;                    // first of 2 consecutive A slabs
  SlabRef b[kMaxBSlabs];
  int nb;
  int gemm;                     // GEMM index (tiles of a GEMM are contiguous)
  int ws_off;                   // float offset of this tile inside one slice's workspace block
};
struct GemmInfo { int tile0, ntiles, nslices, ws_slice_floats; int64_t ws_base; };

struct Segment {                // scatter rule for 64 accumulator columns of one tile
  int tile;                     // OutTile index
  int col0;                     // first accumulator column (multiple of 64)
  int src_col, ncols;           // columns [src_col, src_col+ncols) inside the slab are wanted
  float* dst;                   // dst[m*ld_m + j*ld_j]  for accumulator row m (tile-local + m0) and wanted column j
  int m0, m_valid, ld_m, ld_j;
};

struct WgradParams {
  const uint8_t* saves; const uint8_t* gsaves;
  int save_stride, grad_stride;          // bytes per point tile
  int64_t n_ptiles;
  const OutTile* tiles; const GemmInfo* gemms; int n_gemms;
  float* ws;
};

__global__ void __launch_bounds__(kWThreads, 1) wgrad_kernel(const __grid_constant__ WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* bar_full = bars;        // [2]
  uint64_t* bar_empty = bars + 2;   // [2]
  uint64_t* bar_acc = bars + 4;     // accumulator complete -> epilogue
  uint64_t* bar_drained = bars + 5; // accumulator read out  -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    if (smem_u32(smem) & 1023u) { atomicCAS(&g_watchdog_code, 0u, 901u); __trap(); }
    for (int i = 0; i < 2; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
    mbar_init(bar_acc, 1); mbar_init(bar_drained, 1);
    fence_mbar_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // work of this CTA inside GEMM g: tile index, point-tile range
  auto my_work = [&](int g, int& tile, int64_t& k0, int64_t& k1, int& slice) {
    const GemmInfo gi = p.gemms[g];
    const int c = (int)blockIdx.x;
    if (c >= gi.ntiles * gi.nslices) return false;
    tile = gi.tile0 + c % gi.ntiles;
    slice = c / gi.ntiles;
    k0 = p.n_ptiles * slice / gi.nslices;
    k1 = p.n_ptiles * (slice + 1) / gi.nslices;
    return true;
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (int g = 0; g < p.n_gemms; ++g) {
        int tile, slice; int64_t k0, k1;
        if (!my_work(g, tile, k0, k1, slice)) continue;
        const OutTile t = p.tiles[tile];
        for (int64_t k = k0; k < k1; ++k) {
          mbar_wait(&bar_empty[stage], phase ^ 1, 40);
          uint8_t* dst = smem + stage * kStageBytes;
          mbar_expect_tx(&bar_full[stage], (uint32_t)(2 + t.nb) * kSlabBytes);
          const uint8_t* abase = (t.a.from_grads ? p.gsaves + k * (int64_t)p.grad_stride
                                                 : p.saves + k * (int64_t)p.save_stride);
          bulk_g2s(dst, abase + (size_t)t.a.slab * kSlabBytes, 2 * kSlabBytes, &bar_full[stage]);
          for (int j = 0; j < t.nb; ++j) {
            const uint8_t* bbase = (t.b[j].from_grads ? p.gsaves + k * (int64_t)p.grad_stride
                                                      : p.saves + k * (int64_t)p.save_stride);
            bulk_g2s(dst + (2 + j) * kSlabBytes, bbase + (size_t)t.b[j].slab * kSlabBytes, kSlabBytes,
                     &bar_full[stage]);
          }
          stage ^= 1; if (stage == 0) phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // saved tiles are row-interleaved: 16-byte chunk (8 features) of all 128 points contiguous, i.e. the
      // no-swizzle MN-major operand form: core matrices 128 B apart along K (points), 2048 B along M/N
      constexpr uint64_t tmpl = make_smem_desc_template(128, 2048, kSwizzleNone);
      uint32_t stage = 0, phase = 0, drained_par = 0;
      bool first_gemm = true;
      for (int g = 0; g < p.n_gemms; ++g) {
        int tile, slice; int64_t k0, k1;
        if (!my_work(g, tile, k0, k1, slice)) continue;
        const OutTile t = p.tiles[tile];
        if (!first_gemm) { mbar_wait(bar_drained, drained_par, 41); drained_par ^= 1; tc_fence_after(); }
        first_gemm = false;
        const int n_hi = t.nb > 4 ? 256 : t.nb * 64;        // first instruction: up to 4 slabs
        const int n_lo = t.nb > 4 ? (t.nb - 4) * 64 : 0;    // second: the remaining slab
        for (int64_t k = k0; k < k1; ++k) {
          mbar_wait(&bar_full[stage], phase, 42);
          tc_fence_after();
          const uint32_t a0 = smem_u32(smem + stage * kStageBytes), b0 = a0 + 2 * kSlabBytes;
#pragma unroll
          for (uint32_t s = 0; s < 8; ++s) {                 // 128 points = 8 K-steps of 16 rows
            const uint32_t acc = (k > k0 || s > 0) ? 1u : 0u;
            umma_f16(tmem_base, smem_desc(tmpl, a0 + s * 256), smem_desc(tmpl, b0 + s * 256),
                     make_idesc_f16(128, n_hi, 1, 1), acc);
            if (n_lo)
              umma_f16(tmem_base + 256, smem_desc(tmpl, a0 + s * 256),
                       smem_desc(tmpl, b0 + 4 * kSlabBytes + s * 256), make_idesc_f16(128, n_lo, 1, 1), acc);
          }
          umma_commit(&bar_empty[stage]);
          stage ^= 1; if (stage == 0) phase ^= 1;
        }
        umma_commit(bar_acc);
      }
    }
  } else {
    // epilogue: accumulator -> workspace (plain stores; the reduce kernel sums the slices)
    const int q = warp & 3;                      // TMEM lane quarter of this warp (warps 2,3,4,5 -> 2,3,0,1)
    const int row = q * 32 + lane;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    uint32_t acc_par = 0;
    for (int g = 0; g < p.n_gemms; ++g) {
      int tile, slice; int64_t k0, k1;
      if (!my_work(g, tile, k0, k1, slice)) continue;
      const OutTile t = p.tiles[tile];
      const GemmInfo gi = p.gemms[g];
      mbar_wait(bar_acc, acc_par, 43); acc_par ^= 1;
      tc_fence_after();
      float* dst = p.ws + gi.ws_base + (int64_t)slice * gi.ws_slice_floats + t.ws_off + (size_t)row * (t.nb * 64);
      if (k1 > k0) {
        for (int c0 = 0; c0 < t.nb * 64; c0 += 32) {
          uint32_t v[32];
          tmem_ld32(taddr + c0, v);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<uint4*>(dst + c0 + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        }
      } else {
        for (int c0 = 0; c0 < t.nb * 64; c0 += 4) *reinterpret_cast<uint4*>(dst + c0) = make_uint4(0, 0, 0, 0);
      }
      tc_fence_before();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 64) mbar_arrive(bar_drained);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// out[m][j] = inv_scale * sum over slices of the workspace partials
__global__ void wgrad_reduce_kernel(const Segment* __restrict__ segs, const OutTile* __restrict__ tiles,
                                    const GemmInfo* __restrict__ gemms, const float* __restrict__ ws,
                                    const float* __restrict__ scale) {
  const Segment sg = segs[blockIdx.x];
  const OutTile t = tiles[sg.tile];
  const GemmInfo gi = gemms[t.gemm];
  const float inv = 1.f / *scale;
  const int width = t.nb * 64;
  for (int idx = threadIdx.x; idx < sg.m_valid * sg.ncols; idx += blockDim.x) {
    const int m = idx / sg.ncols, j = idx % sg.ncols;
    const float* src = ws + gi.ws_base + t.ws_off + (size_t)m * width + sg.col0 + sg.src_col + j;
    float acc = 0.f;
    for (int s = 0; s < gi.nslices; ++s) acc += src[(int64_t)s * gi.ws_slice_floats];
    sg.dst[(size_t)(sg.m0 + m) * sg.ld_m + (size_t)j * sg.ld_j] = acc * inv;
  }
}

// ---------------------------------------------------------------------------------------------
// host: the list of GEMMs for a configuration
// ---------------------------------------------------------------------------------------------
struct Plan {
  std::vector<OutTile> tiles;
  std::vector<GemmInfo> gemms;
  std::vector<Segment> segs;
  int64_t ws_floats = 0;
};

// describes where the 64 columns of one B slab go
struct BSlab { SlabRef ref; int src_col, ncols; float* dst; int ld_m, ld_j; };

void add_gemm(Plan& pl, int n_ctas, SlabRef a0, int m_total, const std::vector<BSlab>& bs) {
  GemmInfo gi{};
  gi.tile0 = (int)pl.tiles.size();
  const int m_tiles = m_total / 128;
  std::vector<std::vector<BSlab>> chunks;
  for (size_t i = 0; i < bs.size(); i += kMaxBSlabs)
    chunks.emplace_back(bs.begin() + i, bs.begin() + std::min(bs.size(), i + kMaxBSlabs));
  int ws_off = 0;
  const int g = (int)pl.gemms.size();
  for (int mt = 0; mt < m_tiles; ++mt)
    for (const auto& ch : chunks) {
      OutTile t{};
      t.a = SlabRef{a0.from_grads, (int16_t)(a0.slab + 2 * mt)};
      t.nb = (int)ch.size();
      t.gemm = g;
      t.ws_off = ws_off;
      for (int j = 0; j < t.nb; ++j) {
        t.b[j] = ch[j].ref;
        if (ch[j].dst && ch[j].ncols > 0) {
          Segment s{};
          s.tile = (int)pl.tiles.size(); s.col0 = 64 * j; s.src_col = ch[j].src_col; s.ncols = ch[j].ncols;
          s.dst = ch[j].dst; s.m0 = 128 * mt; s.m_valid = 128; s.ld_m = ch[j].ld_m; s.ld_j = ch[j].ld_j;
          pl.segs.push_back(s);
        }
      }
      ws_off += 128 * 64 * t.nb;
      pl.tiles.push_back(t);
    }
  gi.ntiles = (int)pl.tiles.size() - gi.tile0;
  gi.nslices = std::max(1, n_ctas / gi.ntiles);
  gi.ws_slice_floats = ws_off;
  gi.ws_base = pl.ws_floats;
  pl.ws_floats += (int64_t)ws_off * gi.nslices;
  pl.gemms.push_back(gi);
}

Plan make_plan(const SpnerfNetConfig& c, float* const* G, int n_ctas) {
  Plan pl;
  const SaveMap sm = make_save_map(c);
  const GradMap gm = make_grad_map(c);
  const NetDims d = make_dims(c);
  auto S = [](int slab) { return SlabRef{0, (int16_t)slab}; };
  auto Gr = [](int slab) { return SlabRef{1, (int16_t)slab}; };
  // B slabs of a dense layer input `first..first+n` with destination weight (rows x ld), column offset 0
  auto dense = [&](std::vector<BSlab>& v, int first_slab, int nslabs, float* w, int ld) {
    for (int k = 0; k < nslabs; ++k) v.push_back(BSlab{S(first_slab + k), 0, 64, w ? w + 64 * k : nullptr, ld, 1});
  };
  auto W = [&](int slot) { return G[slot]; };

  // trunk
  for (int L = 0; L < 8; ++L) {
    std::vector<BSlab> bs;
    float* w = W(SPNERF_P_FC_W0 + 2 * L);
    float* bptr = W(SPNERF_P_FC_W0 + 2 * L + 1);
    const int ld = (L == 0) ? d.in_dim : kFeat + (L == c.skip_layer ? d.in_dim : 0);
    if (L == 0) {
      bs.push_back(BSlab{S(sm.inp), 0, d.in_dim, w, ld, 1});
    } else {
      dense(bs, sm.y[L - 1], 8, w, ld);
      if (L == c.skip_layer) bs.push_back(BSlab{S(sm.inp), 0, d.in_dim, w + kFeat, ld, 1});
    }
    bs.push_back(BSlab{S(sm.aux), 0, 1, bptr, 1, 1});
    add_gemm(pl, n_ctas, Gr(gm.G[L]), kFeat, bs);
  }
  {  // feats_from_xyz
    std::vector<BSlab> bs;
    dense(bs, sm.y[7], 8, W(SPNERF_P_FEATS_W), kFeat);
    bs.push_back(BSlab{S(sm.aux), 0, 1, W(SPNERF_P_FEATS_B), 1, 1});
    add_gemm(pl, n_ctas, Gr(gm.g_f), kFeat, bs);
  }
  if (c.sem) {  // logit_from_label.0
    std::vector<BSlab> bs;
    dense(bs, sm.y[7], 8, W(SPNERF_P_SEM0_W), kFeat);
    bs.push_back(BSlab{S(sm.aux), 0, 1, W(SPNERF_P_SEM0_B), 1, 1});
    add_gemm(pl, n_ctas, Gr(gm.G_sem), kHalf, bs);
  }
  {  // rgb_from_xyzdir.0
    std::vector<BSlab> bs;
    dense(bs, sm.f, 8, W(SPNERF_P_RGB0_W), kFeat);
    bs.push_back(BSlab{S(sm.aux), 0, 1, W(SPNERF_P_RGB0_B), 1, 1});
    add_gemm(pl, n_ctas, Gr(gm.G_rgb), kHalf, bs);
  }
  {  // sun_v_net.0: input [feats, sun_dir]; the aux slab is loaded twice (bias column, sun columns)
    std::vector<BSlab> bs;
    dense(bs, sm.f, 8, W(SPNERF_P_SUN0_W), kFeat + 3);
    bs.push_back(BSlab{S(sm.aux), 0, 1, W(SPNERF_P_SUN0_W + 1), 1, 1});
    bs.push_back(BSlab{S(sm.aux), 1, 3, W(SPNERF_P_SUN0_W) + kFeat, kFeat + 3, 1});
    add_gemm(pl, n_ctas, Gr(gm.G_sun[0]), kHalf, bs);
  }
  for (int j = 1; j < 3; ++j) {  // sun_v_net.2 / .4
    std::vector<BSlab> bs;
    dense(bs, sm.sun_y[j - 1], 4, W(SPNERF_P_SUN0_W + 2 * j), kHalf);
    bs.push_back(BSlab{S(sm.aux), 0, 1, W(SPNERF_P_SUN0_W + 2 * j + 1), 1, 1});
    add_gemm(pl, n_ctas, Gr(gm.G_sun[j]), kHalf, bs);
  }
  if (c.beta) {  // beta_from_xyz.0: input [feats, t_emb]
    std::vector<BSlab> bs;
    dense(bs, sm.f, 8, W(SPNERF_P_BETA0_W), kFeat + c.t_dim);
    bs.push_back(BSlab{S(sm.aux), 0, 1, W(SPNERF_P_BETA0_B), 1, 1});
    bs.push_back(BSlab{S(sm.aux), 4, c.t_dim, W(SPNERF_P_BETA0_W) + kFeat, kFeat + c.t_dim, 1});
    add_gemm(pl, n_ctas, Gr(gm.G_beta), kHalf, bs);
  }
  // tiny last layers, transposed: D[hidden j][small column c] -> W2[c][j]
  auto small_head = [&](int a_slab, int m_total, int src_col, int ncols, float* w2, int hidden) {
    std::vector<BSlab> bs;
    bs.push_back(BSlab{Gr(gm.gsmall), src_col, ncols, w2, 1, hidden});
    add_gemm(pl, n_ctas, S(a_slab), m_total, bs);
  };
  small_head(sm.rgb_y, kHalf, 0, 3, W(SPNERF_P_RGB2_W), kHalf);
  small_head(sm.sun_y[2], kHalf, 3, 1, W(SPNERF_P_SUN0_W + 6), kHalf);
  small_head(sm.y[7], kFeat, 4, 1, W(SPNERF_P_SIGMA_W), kFeat);
  if (c.beta) small_head(sm.beta_y, kHalf, 5, 1, W(SPNERF_P_BETA2_W), kHalf);
  if (c.sem) small_head(sm.sem_y, kHalf, 8, c.num_sem_classes, W(SPNERF_P_SEM2_W), kHalf);
  return pl;
}

int n_sms() {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms > 0 ? sms : 148;
}

size_t table_bytes(const Plan& pl) {
  return pl.tiles.size() * sizeof(OutTile) + pl.gemms.size() * sizeof(GemmInfo) + pl.segs.size() * sizeof(Segment) + 64;
}

}  // namespace

extern "C" int64_t spnerf_mlp_wgrad_workspace_bytes(const SpnerfNetConfig* cfg) {
  if (!cfg) return -1;
  float* G[SPNERF_NUM_PARAMS] = {};
  // the plan depends on the SM count only through the slice counts; query-time device = run-time device
  Plan pl = make_plan(*cfg, G, n_sms());
  return (int64_t)pl.ws_floats * 4 + 65536;   // partial tiles + room for the device tables
}

namespace {
struct Tables { OutTile* tiles; GemmInfo* gemms; Segment* segs; };
Tables table_ptrs(const Plan& pl, void* workspace) {
  uint8_t* tab = static_cast<uint8_t*>(workspace) + (size_t)pl.ws_floats * 4;
  Tables t;
  t.tiles = reinterpret_cast<OutTile*>(tab);
  t.gemms = reinterpret_cast<GemmInfo*>(t.tiles + pl.tiles.size());
  t.segs = reinterpret_cast<Segment*>(t.gemms + pl.gemms.size());
  return t;
}
}  // namespace

// Uploads the GEMM / scatter tables to the tail of the workspace.  Call once per (configuration,
// gradient pointers, workspace); synchronises the stream (the tables come from pageable memory).
extern "C" int spnerf_mlp_wgrad_prepare(const SpnerfMlpWgrad* a, void* stream_) {
  if (!a || !a->grads_host || !a->workspace) return SPNERF_ERR_BAD_ARG;
  if (a->cfg.feat != 512 || a->cfg.layers != 8 || a->cfg.skip_layer != 4) return SPNERF_ERR_UNSUPPORTED;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  Plan pl = make_plan(a->cfg, a->grads_host, n_sms());
  const size_t tb = table_bytes(pl);
  if ((int64_t)pl.ws_floats * 4 + (int64_t)tb > a->workspace_bytes || tb > 65536) return SPNERF_ERR_WORKSPACE;
  const Tables t = table_ptrs(pl, a->workspace);
  cudaMemcpyAsync(t.tiles, pl.tiles.data(), pl.tiles.size() * sizeof(OutTile), cudaMemcpyHostToDevice, stream);
  cudaMemcpyAsync(t.gemms, pl.gemms.data(), pl.gemms.size() * sizeof(GemmInfo), cudaMemcpyHostToDevice, stream);
  cudaMemcpyAsync(t.segs, pl.segs.data(), pl.segs.size() * sizeof(Segment), cudaMemcpyHostToDevice, stream);
  cudaError_t e = cudaStreamSynchronize(stream);
  return e == cudaSuccess ? 0 : -(int)e;
}

extern "C" int spnerf_mlp_bwd_weights(const SpnerfMlpWgrad* a, void* stream_) {
  if (!a || !a->saves || !a->grad_saves || !a->scale || !a->grads_host || !a->workspace) return SPNERF_ERR_BAD_ARG;
  if (a->cfg.feat != 512 || a->cfg.layers != 8 || a->cfg.skip_layer != 4) return SPNERF_ERR_UNSUPPORTED;
  if (a->n_points <= 0) return a->n_points == 0 ? 0 : SPNERF_ERR_BAD_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int ctas = n_sms();
  Plan pl = make_plan(a->cfg, a->grads_host, ctas);     // host-side counts only; tables were uploaded by _prepare
  if ((int64_t)pl.ws_floats * 4 + (int64_t)table_bytes(pl) > a->workspace_bytes) return SPNERF_ERR_WORKSPACE;
  const Tables t = table_ptrs(pl, a->workspace);
  WgradParams p;
  p.saves = static_cast<const uint8_t*>(a->saves); p.gsaves = static_cast<const uint8_t*>(a->grad_saves);
  p.save_stride = make_save_map(a->cfg).total * kSlabBytes;
  p.grad_stride = make_grad_map(a->cfg).total * kSlabBytes;
  p.n_ptiles = (a->n_points + kTileM - 1) / kTileM;
  p.tiles = t.tiles; p.gemms = t.gemms; p.n_gemms = (int)pl.gemms.size();
  p.ws = static_cast<float*>(a->workspace);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemW);
    if (e != cudaSuccess) return -(int)e;
    attr_set = true;
  }
  wgrad_kernel<<<ctas, kWThreads, kSmemW, stream>>>(p);
  wgrad_reduce_kernel<<<(unsigned)pl.segs.size(), 256, 0, stream>>>(t.segs, t.tiles, t.gemms, p.ws, a->scale);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}

SPNERF_DEFINE_WATCHDOG_GETTER(spnerf_watchdog_code_wgrad)
