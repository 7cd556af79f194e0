// Host-side plan builder and the device pack kernel for the fused point-network kernels.
//
// The forward / backward-data kernels consume the weights as a linear stream of pre-swizzled fp16
// B tiles ("items"), in exactly the order the MMA issuer walks them, so the weight producer is a
// sequence of 1-D bulk copies.  This file derives that order from the layer graph of
// models/spnerf.py:202-264 and converts the fp32 parameters into it.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <vector>
#include <cstring>
#include "net_plan.h"
#include "sm100.cuh"

using namespace net;

namespace {

struct PackItem {
  const float* src;   // W, row-major (rows x ld)
  int ld;
  int rows, cols;     // valid extent of W
  int row0, col0;     // tile row i / tile col k  <->  W[row0+i][col0+k]  (or W[row0+k][col0+i] if transposed)
  int transpose;
  int mode;           // 0: fp16(w)   1: fp16(w - fp16(w))
  int n;              // tile rows to fill (starting at dst_row0)
  int dst_row0;
  uint32_t dst_off16;
};

__device__ __forceinline__ void pack_item(const PackItem& it, uint8_t* __restrict__ blob) {
  if (it.transpose) {
    // backward tiles: tile (r, k) = W[row0 + k][col0 + r].  A thread-per-chunk gather reads 8 floats a row stride apart
    // (the first version: 130 us per repack, ~8x the traffic of the forward tiles); instead 64 x 64 blocks of W go
    // through shared memory with reads that run along W's rows.
    __shared__ float tile[64][65];
    for (int r0 = blockIdx.y * 64; r0 < it.n; r0 += gridDim.y * 64) {
      for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
        const int k = i >> 6, rr = i & 63;
        const int wr = it.row0 + k, wc = it.col0 + r0 + rr;
        tile[k][rr] = (r0 + rr < it.n && wr < it.rows && wc < it.cols) ? it.src[(size_t)wr * it.ld + wc] : 0.f;
      }
      __syncthreads();
      for (int i = threadIdx.x; i < 64 * 8; i += blockDim.x) {
        const int rr = i & 63, c = i >> 6;
        if (r0 + rr < it.n) {
          __half h[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float w = tile[c * 8 + e][rr];
            const __half hi = __float2half_rn(w);
            h[e] = it.mode == 0 ? hi : __float2half_rn(w - __half2float(hi));
          }
          uint8_t* dst = blob + (size_t)it.dst_off16 * 16 + sm100::slab_chunk_offset(it.dst_row0 + r0 + rr, c);
          *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(h);
        }
      }
      __syncthreads();
    }
    return;
  }
  for (int t = blockIdx.y * blockDim.x + threadIdx.x; t < it.n * 8; t += gridDim.y * blockDim.x) {
    const int r = t >> 3, c = t & 7;   // tile row, 16-byte chunk: 8 threads read 256 contiguous bytes of one row of W
    __half h[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int wr = it.row0 + r, wc = it.col0 + c * 8 + e;
      float w = 0.f;
      if (wr < it.rows && wc < it.cols) w = it.src[(size_t)wr * it.ld + wc];
      __half hi = __float2half_rn(w);
      h[e] = it.mode == 0 ? hi : __float2half_rn(w - __half2float(hi));
    }
    uint8_t* dst = blob + (size_t)it.dst_off16 * 16 + sm100::slab_chunk_offset(it.dst_row0 + r, c);
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(h);
  }
}

struct CopyItem {
  const float* src;
  int dst, rows, cols, ld, transpose;   // dst[r*cols+c] = src[r*ld+c]  (transpose: dst[c*dst_ld+r])
  int dst_ld;
};

__device__ __forceinline__ void small_copy_item(const CopyItem& it, float* __restrict__ small) {
  if (!it.src || blockIdx.y != 0) return;
  const int n = it.rows * it.cols;
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    const int r = t / it.cols, c = t % it.cols;
    const float v = it.src[(size_t)r * it.ld + c];
    small[it.dst + (it.transpose ? c * it.dst_ld + r : t)] = v;
  }
}

// B tile of an aux step (net_plan.h): 16 columns per output unit, no-swizzle K-major.
struct AuxItem {
  const float* bias; int bias_n;          // bias[row0 + i], valid while row0 + i < bias_n
  const float* wx; int wx_ld, wx_rows, wx_col0, wx_ncols, dst_col;   // B[i][dst_col+e] = wx[row0+i][wx_col0+e]
  // single columns of a weight matrix (encoded-input columns beyond the input slab, net_plan.h AuxExtra):
  // B[i][ex_dst[k]] = ex_w[row0+i][ex_src[k]], high part (mode 0) or fp16 residual (mode 1)
  const float* ex_w; int ex_ld, ex_rows, n_ex;
  int8_t ex_src_off[12], ex_dst[12], ex_mode[12];
  int ex_col0;
  int row0, n;
  uint32_t dst_off16;
};

__device__ __forceinline__ void aux_pack_item(const AuxItem& it, uint8_t* __restrict__ blob) {
  if (blockIdx.y != 0) return;
  for (int i = threadIdx.x; i < it.n; i += blockDim.x) {
    __half h[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) h[e] = __float2half_rn(0.f);
    const int r = it.row0 + i;
    if (it.bias && r < it.bias_n) {
      const float b = it.bias[r];
      const __half hi = __float2half_rn(b);
      h[kAuxColOne] = hi;
      h[kAuxColOneLo] = __float2half_rn(b - __half2float(hi));
    }
    if (it.wx && r < it.wx_rows)
      for (int e = 0; e < it.wx_ncols; ++e) h[it.dst_col + e] = __float2half_rn(it.wx[(size_t)r * it.wx_ld + it.wx_col0 + e]);
    if (it.ex_w && r < it.ex_rows)
      for (int k = 0; k < it.n_ex; ++k) {
        const float w = it.ex_w[(size_t)r * it.ex_ld + it.ex_col0 + it.ex_src_off[k]];
        const __half hi = __float2half_rn(w);
        h[it.ex_dst[k]] = it.ex_mode[k] == 0 ? hi : __float2half_rn(w - __half2float(hi));
      }
    uint8_t* dst = blob + (size_t)it.dst_off16 * 16;
    *reinterpret_cast<uint4*>(dst + aux_offset(i, 0)) = *reinterpret_cast<const uint4*>(h);
    *reinterpret_cast<uint4*>(dst + aux_offset(i, 8)) = *reinterpret_cast<const uint4*>(h + 8);
  }
}

// One launch converts everything: blockIdx.x walks [forward weight tiles | backward weight tiles | small fp32
// copies | aux tiles] (three launches less per optimiser step than a kernel per table).
struct PackTables {
  const PackItem* f; int nf;
  const PackItem* b; int nb;
  const CopyItem* c; int nc;
  const AuxItem* a; int na;
};
__global__ void __launch_bounds__(256) pack_all_kernel(const PackTables t, uint8_t* __restrict__ fwd_blob,
                                                       uint8_t* __restrict__ bwd_blob, float* __restrict__ small) {
  int i = blockIdx.x;
  if (i < t.nf) { pack_item(t.f[i], fwd_blob); return; }
  i -= t.nf;
  if (i < t.nb) { pack_item(t.b[i], bwd_blob); return; }
  i -= t.nb;
  if (i < t.nc) { small_copy_item(t.c[i], small); return; }
  i -= t.nc;
  if (i < t.na) aux_pack_item(t.a[i], fwd_blob);
}

struct AuxSpec {
  bool on = false;
  const float* bias = nullptr; int bias_n = 0;
  const float* wx = nullptr; int wx_ld = 0, wx_rows = 0, wx_col0 = 0, wx_ncols = 0, dst_col = 0;
  const float* ex_w = nullptr; int ex_ld = 0, ex_rows = 0, n_ex = 0, ex_col0 = 0;
  int8_t ex_src_off[12] = {}, ex_dst[12] = {}, ex_mode[12] = {};
  void extra(int src_off, int dst, int mode) { ex_src_off[n_ex] = (int8_t)src_off; ex_dst[n_ex] = (int8_t)dst; ex_mode[n_ex] = (int8_t)mode; ++n_ex; }
};
inline AuxSpec aux_bias(const float* b, int n) { AuxSpec a; a.on = true; a.bias = b; a.bias_n = n; return a; }

struct Builder {
  std::vector<MmaStep> steps;
  std::vector<PackItem> items;
  std::vector<AuxItem> aux_items;
  uint32_t off16 = 0;

  // one accumulation chunk: D[:, tmem_col : tmem_col+n] (+)= sum over the listed sources.
  // Each source names the A slab, the K offset inside its weight matrix, the K=16 steps to issue and
  // the pack mode; a source may bring its own weight matrix (several layers feeding one accumulator).
  // Forward (transpose=false): B tile row i <-> W row row0+i, tile col k <-> W col col0+k.
  // Backward (transpose=true): B tile row i <-> W col row0+i (an input unit), col k <-> W row col0+k.
  struct Src { int a_slab, col0, ksteps, mode; const float* W = nullptr; int rows = 0, cols = 0; };
  std::vector<size_t> chunk_begin;      // first step of every chunk of the open phase
  size_t phase_begin = 0;
  void chunk(const float* W, int rows, int cols, int row0, int n, int tmem_col, const std::vector<Src>& srcs,
             bool transpose = false, bool accumulate_all = false, const AuxSpec& aux = AuxSpec()) {
    chunk_begin.push_back(steps.size());
    bool first = !accumulate_all;
    for (const Src& s : srcs) {
      MmaStep st{};
      st.w_off16 = off16;
      st.n = (uint16_t)n;
      st.tmem_col = (uint16_t)tmem_col;
      st.a_slab = (uint8_t)s.a_slab;
      st.ksteps = (uint8_t)s.ksteps;
      st.first = first ? 1 : 0;
      st.last = 0;
      st.bytes16 = (uint16_t)(n * 128 / 16);
      steps.push_back(st);
      PackItem it{};
      const bool own = s.rows > 0;
      it.src = own ? s.W : W; it.ld = own ? s.cols : cols; it.rows = own ? s.rows : rows; it.cols = own ? s.cols : cols;
      it.row0 = transpose ? s.col0 : row0;
      it.col0 = transpose ? row0 : s.col0;
      it.transpose = transpose ? 1 : 0;
      it.mode = s.mode; it.n = n; it.dst_row0 = 0; it.dst_off16 = off16;
      items.push_back(it);
      off16 += (uint32_t)(n * 128 / 16);
      first = false;
    }
    if (aux.on) {      // bias (+ per-ray input columns) as one more K=16 step
      MmaStep st{};
      st.w_off16 = off16;
      st.n = (uint16_t)n;
      st.tmem_col = (uint16_t)tmem_col;
      st.a_slab = (uint8_t)kAuxSlab;
      st.ksteps = 1;
      st.first = first ? 1 : 0;
      st.bytes16 = (uint16_t)(n * 32 / 16);
      steps.push_back(st);
      AuxItem it{};
      it.bias = aux.bias; it.bias_n = aux.bias_n;
      it.wx = aux.wx; it.wx_ld = aux.wx_ld; it.wx_rows = aux.wx_rows; it.wx_col0 = aux.wx_col0;
      it.wx_ncols = aux.wx_ncols; it.dst_col = aux.dst_col;
      it.ex_w = aux.ex_w; it.ex_ld = aux.ex_ld; it.ex_rows = aux.ex_rows; it.n_ex = aux.n_ex; it.ex_col0 = aux.ex_col0;
      for (int k = 0; k < aux.n_ex; ++k) { it.ex_src_off[k] = aux.ex_src_off[k]; it.ex_dst[k] = aux.ex_dst[k]; it.ex_mode[k] = aux.ex_mode[k]; }
      it.row0 = row0; it.n = n; it.dst_off16 = off16;
      aux_items.push_back(it);
      off16 += (uint32_t)(n * 32 / 16);
    }
  }
  // Fuses runs of `per_item` consecutive K-slab steps of the chunk that was just added into one ring item each
  // (the fused B tiles of one CTA, per_item * (n / 2) * 128 bytes, must fit a ring stage).  A narrow product is
  // little tensor work per slab (n = 16: 32 cycles, n = 128: 256 cycles) but every ring item costs the issuer a
  // full wait / issue / commit round trip (~550-750 cycles).  The blob keeps each CTA's half of the fused slabs
  // contiguous: [rank 0: slab 0..S-1][rank 1: slab 0..S-1].  An aux step at the end of the chunk stays as it is.
  void merge_last_chunk(int per_item) {
    const size_t c0 = chunk_begin.back();
    size_t ns = steps.size() - c0;
    const bool has_aux = ns > 0 && steps.back().a_slab == kAuxSlab;
    const MmaStep aux_step = steps.back();
    if (has_aux) --ns;
    if (ns < 2 || per_item < 2) return;
    const size_t i0 = items.size() - ns;                            // aux steps have no PackItem
    std::vector<MmaStep> fused;
    std::vector<PackItem> halves;
    for (size_t g0 = 0; g0 < ns; g0 += (size_t)per_item) {
      const size_t gn = std::min(ns - g0, (size_t)per_item);
      MmaStep m = steps[c0 + g0];
      const uint32_t half16 = (uint32_t)(m.n / 2) * 128 / 16;       // one CTA's half of one slab, 16-byte units
      uint32_t b16 = 0, ks = 0;
      for (size_t k = 0; k < gn; ++k) {
        const MmaStep& st = steps[c0 + g0 + k];
        b16 += st.bytes16; ks += st.ksteps;
        for (int r = 0; r < 2; ++r) {
          PackItem h = items[i0 + g0 + k];
          h.n = m.n / 2;
          if (h.transpose) h.col0 += r * (m.n / 2); else h.row0 += r * (m.n / 2);
          h.dst_row0 = 0;
          h.dst_off16 = m.w_off16 + (uint32_t)r * (uint32_t)gn * half16 + (uint32_t)k * half16;
          halves.push_back(h);
        }
      }
      m.ksteps = (uint8_t)ks;
      m.bytes16 = (uint16_t)b16;
      fused.push_back(m);
    }
    items.resize(i0);
    items.insert(items.end(), halves.begin(), halves.end());
    steps.resize(c0);
    steps.insert(steps.end(), fused.begin(), fused.end());
    if (has_aux) steps.push_back(aux_step);
  }
  // Closes a phase.  Its chunks (independent accumulator column ranges) are dealt to the two issuer
  // lanes, balancing step counts, and the ring order interleaves the lanes so that both issuers
  // always have an item in flight.  The order of the steps inside a chunk is preserved.
  // split = true (four chunks of a quarter of the accumulator each): the first two chunks -- one per issuer, interleaved
  // -- make up the first accumulator half, the last two the second; the last item of the first half carries the `half`
  // mark.  Otherwise every chunk belongs to one issuer and the two lanes are interleaved over the whole phase.
  // early_free_slabs > 0 (split phases of the backward trunk): mark the item after which nothing reads A slabs
  // 0 .. early_free_slabs-1 any more (MmaStep::half bit 1)
  void end_phase(bool split = false, int early_free_slabs = 0) {
    const size_t e = steps.size();
    chunk_begin.push_back(e);
    split = split && chunk_begin.size() == 5;      // the caller built four chunks because its direction splits
    if (split) {
      std::vector<MmaStep> out;
      for (int part = 0; part < 2; ++part) {
        std::vector<MmaStep> ls[2];
        for (int l = 0; l < 2; ++l)
          for (size_t i = chunk_begin[2 * part + l]; i < chunk_begin[2 * part + l + 1]; ++i) {
            MmaStep st = steps[i];
            st.lane = (uint8_t)l;
            ls[l].push_back(st);
          }
        size_t a = 0, b = 0;
        while (a < ls[0].size() || b < ls[1].size()) {
          if (a < ls[0].size()) out.push_back(ls[0][a++]);
          if (b < ls[1].size()) out.push_back(ls[1][b++]);
        }
        if (part == 0) out.back().half = 1;
      }
      if (early_free_slabs > 0) {
        size_t p1 = 0;
        while (!(out[p1].half & 1)) ++p1;
        size_t q = p1 + 1;                                   // first item of the second half
        for (size_t k = out.size(); k-- > p1 + 1;)
          if (out[k].a_slab != kAuxSlab && (int)out[k].a_slab < early_free_slabs) { q = k; break; }
        out[q].half |= 2;
      }
      for (size_t k = 0; k < out.size(); ++k) steps[phase_begin + k] = out[k];
      steps[e - 1].last = 1;
      chunk_begin.clear();
      phase_begin = e;
      return;
    }
    std::vector<MmaStep> lane_steps[2];
    for (size_t c = 0; c + 1 < chunk_begin.size(); ++c) {
      const int lane = lane_steps[0].size() <= lane_steps[1].size() ? 0 : 1;
      for (size_t i = chunk_begin[c]; i < chunk_begin[c + 1]; ++i) {
        MmaStep st = steps[i];
        st.lane = (uint8_t)lane;
        lane_steps[lane].push_back(st);
      }
    }
    size_t o = phase_begin, a = 0, b = 0;
    while (a < lane_steps[0].size() || b < lane_steps[1].size()) {
      if (a < lane_steps[0].size()) steps[o++] = lane_steps[0][a++];
      if (b < lane_steps[1].size()) steps[o++] = lane_steps[1][b++];
    }
    steps[e - 1].last = 1;
    chunk_begin.clear();
    phase_begin = e;
  }
};

// Forward order; must match the phase sequence of mlp_fwd.cu.  Every chunk ends with an aux step
// that adds the layer's bias (and, for the sun / beta heads, the per-ray input columns).
void build_forward(const SpnerfNetConfig& c, const float* const* P, Builder& b) {
  const NetDims d = make_dims(c);
  const int ink = d.in_ksteps;
  const int F = c.feat, H = F / 2, Q = H / 2;      // trunk width, head width, head chunk width (one chunk per issuer)
  const int narrow_merge = F == 512 ? 2 : 4;       // K slabs per ring item of a Q-wide chunk: 16 KB per CTA either way
  // a layer that fills the whole accumulator: two chunks of H columns, or (split phases, net_plan.h) four of Q
  const int parts = kSplitFwd ? 4 : 2, PW = 2 * H / parts;
  auto act8 = [](int ncols) {
    std::vector<Builder::Src> v;
    for (int k = 0; k < ncols / 64; ++k) v.push_back({k, 64 * k, 4, 0});
    return v;
  };
  // layer 0: split-precision product  in_hi*W_hi + in_lo*W_hi + in_hi*W_lo   (models/spnerf.py:202); input columns
  // beyond the slab ride in the aux step with the same three products (net_plan.h AuxExtra)
  const AuxExtra ax = make_aux_extra(c);
  for (int g = 0; g < parts; ++g) {
    AuxSpec a0 = aux_bias(P[SPNERF_P_FC_W0 + 1], F);
    if (ax.n > 0) {
      a0.ex_w = P[SPNERF_P_FC_W0]; a0.ex_ld = d.in_dim; a0.ex_rows = F; a0.ex_col0 = 64;
      for (int q = 0; q < ax.n; ++q) { a0.extra(q, ax.col_hi[q], 0); a0.extra(q, ax.col_lo[q], 0); a0.extra(q, ax.col_dup[q], 1); }
    }
    b.chunk(P[SPNERF_P_FC_W0], F, d.in_dim, g * PW, PW, g * PW,
            {{kSlabInpHi, 0, ink, 0}, {kSlabInpLo, 0, ink, 0}, {kSlabInpHi, 0, ink, 1}}, false, false, a0);
  }
  b.end_phase(true);
  for (int i = 1; i < 8; ++i) {   // models/spnerf.py:203-208, skip concat [h, input] at :327
    const bool skip = (i == c.skip_layer);
    const int cols = F + (skip ? d.in_dim : 0);
    for (int g = 0; g < parts; ++g) {
      auto srcs = act8(F);
      if (skip) srcs.push_back({kSlabInpHi, F, ink, 0});
      AuxSpec ai = aux_bias(P[SPNERF_P_FC_W0 + 2 * i + 1], F);
      if (skip && ax.n > 0) {        // skip concat [h, input]: input columns 64.. through the aux step (high parts)
        ai.ex_w = P[SPNERF_P_FC_W0 + 2 * i]; ai.ex_ld = cols; ai.ex_rows = F; ai.ex_col0 = F + 64;
        for (int q = 0; q < ax.n; ++q) ai.extra(q, ax.col_hi[q], 0);
      }
      b.chunk(P[SPNERF_P_FC_W0 + 2 * i], F, cols, g * PW, PW, g * PW, srcs, false, false, ai);
      // chunks narrower than 256 columns: several K slabs per ring item (16 KB per CTA); the skip layer's trailing
      // input-slab step stays on its own (the activation slabs before it pair up evenly)
      if (256 / PW > 1) b.merge_last_chunk(256 / PW);
    }
    b.end_phase(true);
  }
  // heads reading the trunk output h: semantic hidden (:218-223) and sigma (:212).  The 256-wide hidden layer
  // runs as two 128-wide chunks, one per issuer (a single 256-wide chunk left one issuer alone with 9 items, which
  // it cannot issue at the tensor pipe's pace), two K slabs per ring item; the 1-wide sigma head is one fused item.
  if (c.sem)
    for (int g = 0; g < 2; ++g) {
      b.chunk(P[SPNERF_P_SEM0_W], H, F, g * Q, Q, g * Q, act8(F), false, false,
              aux_bias(P[SPNERF_P_SEM0_B], H));
      b.merge_last_chunk(narrow_merge);
    }
  {
    // sigma: B row 0 = fp16(w), row 1 = residual; the epilogue adds the two accumulator columns
    auto srcs = act8(F);
    const size_t i0 = b.items.size();
    b.chunk(P[SPNERF_P_SIGMA_W], 1, F, 0, 16, H, srcs, false, false, aux_bias(P[SPNERF_P_SIGMA_B], 1));
    b.merge_last_chunk(F / 64);       // PackItems are now per (slab, CTA half): rank 0's half holds rows 0..7
    const size_t i1 = b.items.size();
    std::vector<PackItem> kept;
    for (size_t i = i0; i < i1; ++i) {
      PackItem hi = b.items[i];
      if (hi.row0 != 0) continue;     // rank 1's rows 8..15: never read (accumulator columns 264..271 are unused)
      hi.n = 1;
      PackItem lo = hi;
      lo.mode = 1; lo.dst_row0 = 1;
      kept.push_back(hi);
      kept.push_back(lo);
    }
    b.items.resize(i0);
    b.items.insert(b.items.end(), kept.begin(), kept.end());
  }
  b.end_phase();
  for (int g = 0; g < parts; ++g) {   // feats_from_xyz (:215)
    b.chunk(P[SPNERF_P_FEATS_W], F, F, g * PW, PW, g * PW, act8(F), false, false,
            aux_bias(P[SPNERF_P_FEATS_B], F));
    if (256 / PW > 1) b.merge_last_chunk(256 / PW);
  }
  b.end_phase(true);
  AuxSpec sun0 = aux_bias(P[SPNERF_P_SUN0_W + 1], H);          // + sun direction columns (:351)
  sun0.wx = P[SPNERF_P_SUN0_W]; sun0.wx_ld = F + 3; sun0.wx_rows = H; sun0.wx_col0 = F;
  sun0.wx_ncols = 3; sun0.dst_col = kAuxColSun;
  b.chunk(P[SPNERF_P_RGB0_W], H, F, 0, H, 0, act8(F), false, false,
          aux_bias(P[SPNERF_P_RGB0_B], H));                                           // :226-231
  if (c.beta) {
    AuxSpec beta0 = aux_bias(P[SPNERF_P_BETA0_B], H);          // + transient embedding columns (:360)
    beta0.wx = P[SPNERF_P_BETA0_W]; beta0.wx_ld = F + c.t_dim; beta0.wx_rows = H; beta0.wx_col0 = F;
    beta0.wx_ncols = c.t_dim; beta0.dst_col = kAuxColT;
    b.chunk(P[SPNERF_P_BETA0_W], H, F + c.t_dim, 0, H, H, act8(F), false, false, beta0);   // :258-264
    b.end_phase();
    b.chunk(P[SPNERF_P_SUN0_W], H, F + 3, 0, H, 0, act8(F), false, false, sun0);               // :234-241
    b.end_phase();
  } else {
    b.chunk(P[SPNERF_P_SUN0_W], H, F + 3, 0, H, H, act8(F), false, false, sun0);
    b.end_phase();
  }
  for (int j = 1; j <= 2; ++j) {      // sun layers 1, 2: two 128-wide chunks, one per issuer lane
    for (int g = 0; g < 2; ++g) {
      b.chunk(P[SPNERF_P_SUN0_W + 2 * j], H, H, g * Q, Q, g * Q, act8(H), false, false,
              aux_bias(P[SPNERF_P_SUN0_W + 2 * j + 1], H));
      b.merge_last_chunk(2);          // 128-wide chunks: two K slabs (2 x 8 KB per CTA) per ring item
    }
    b.end_phase();
  }
}

int validate(const SpnerfNetConfig* c) {
  if (!c) return SPNERF_ERR_BAD_ARG;
  if (!feat_supported(c->feat) || c->layers != 8 || c->skip_layer != 4) return SPNERF_ERR_UNSUPPORTED;
  if (c->sem && (c->num_sem_classes < 1 || c->num_sem_classes > 8 || c->emb_dim < 1 || c->emb_dim > 8))
    return SPNERF_ERR_UNSUPPORTED;
  if (c->beta && (c->t_dim < 1 || c->t_dim > 8)) return SPNERF_ERR_UNSUPPORTED;
  if (make_aux_extra(*c).n < 0) return SPNERF_ERR_UNSUPPORTED;      // encoded input beyond the slab + free aux columns
  return 0;
}

}  // namespace

void build_backward(const SpnerfNetConfig& c, const float* const* P, std::vector<MmaStep>& steps,
                    std::vector<PackItem>* items, uint32_t* off16);   // mlp_pack_bwd section below

// cached per configuration; entries are never freed or moved (callers keep the pointer for a launch)
#include <map>
#include <mutex>
#include <array>
#include <memory>
const net::StepTable* net::step_table(const SpnerfNetConfig& cfg, int backward) {
  static std::mutex mu;
  static std::map<std::array<int32_t, 10>, std::unique_ptr<StepTable>> cache;
  std::array<int32_t, 10> key{cfg.feat, cfg.layers, cfg.skip_layer, cfg.mapping, cfg.sem, cfg.num_sem_classes,
                              cfg.emb_dim, cfg.beta, cfg.t_dim, backward};
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second.get();
  const float* P[SPNERF_NUM_PARAMS] = {};
  std::vector<MmaStep> steps;
  if (backward) {
    uint32_t off = 0;
    build_backward(cfg, P, steps, nullptr, &off);
  } else {
    Builder f;
    build_forward(cfg, P, f);
    steps = f.steps;
  }
  if ((int)steps.size() > kMaxSteps) return nullptr;
  std::unique_ptr<StepTable> t(new StepTable());
  std::memset(t.get(), 0, sizeof(StepTable));
  t->n = (int)steps.size();
  std::memcpy(t->s, steps.data(), steps.size() * sizeof(MmaStep));
  const StepTable* r = t.get();
  cache[key] = std::move(t);
  return r;
}

// debug: the step list as 8 int32 per step [n, tmem_col, a_slab, ksteps, first, last, lane, 0]; returns
// the number of steps (or a negative error)
extern "C" int spnerf_debug_step_table(const SpnerfNetConfig* cfg, int backward, int32_t* out, int max_steps) {
  if (!cfg || validate(cfg) != 0) return SPNERF_ERR_UNSUPPORTED;
  const net::StepTable* t = net::step_table(*cfg, backward);
  if (!t) return SPNERF_ERR_UNSUPPORTED;
  for (int i = 0; i < t->n && i < max_steps && out; ++i) {
    const MmaStep& s = t->s[i];
    int32_t* o = out + 8 * i;
    o[0] = s.n; o[1] = s.tmem_col; o[2] = s.a_slab; o[3] = s.ksteps; o[4] = s.first; o[5] = s.last; o[6] = s.lane; o[7] = s.half;
  }
  return t->n;
}

extern "C" int spnerf_net_sizes(const SpnerfNetConfig* cfg, SpnerfNetSizes* s) {
  int rc = validate(cfg);
  if (rc) return rc;
  if (!s) return SPNERF_ERR_BAD_ARG;
  const float* P[SPNERF_NUM_PARAMS] = {};
  Builder f;
  build_forward(*cfg, P, f);
  std::memset(s, 0, sizeof(*s));
  s->fwd_blob_bytes = (int64_t)f.off16 * 16;
  s->fwd_steps = (int32_t)f.steps.size();
  std::vector<MmaStep> bsteps;
  uint32_t boff = 0;
  build_backward(*cfg, P, bsteps, nullptr, &boff);
  s->bwd_blob_bytes = (int64_t)boff * 16;
  s->bwd_steps = (int32_t)bsteps.size();
  s->small_floats = make_small_offsets(*cfg).total;
  s->steps_bytes = (int64_t)kMaxSteps * sizeof(MmaStep);
  s->save_slabs_per_tile = make_save_map(*cfg).total;
  s->grad_slabs_per_tile = make_grad_map(*cfg).total;
  const NetDims d = make_dims(*cfg);
  s->n_out = d.n_out;
  s->in_dim = d.in_dim;
  s->tile_points = kTileM;
  if (s->fwd_steps > kMaxSteps || s->bwd_steps > kMaxSteps) return SPNERF_ERR_UNSUPPORTED;
  return 0;
}

namespace {
struct PackPlan {
  Builder f;
  std::vector<MmaStep> bsteps;
  std::vector<PackItem> bitems;
  uint32_t boff = 0;
  std::vector<CopyItem> cp;
  SmallOffsets o;
  size_t bytes_f, bytes_b, bytes_c, bytes_a;
};

void make_pack_plan(const SpnerfNetConfig* cfg, const float* const* P, PackPlan& pl) {
  const int F = cfg->feat, H = F / 2;
  build_forward(*cfg, P, pl.f);
  build_backward(*cfg, P, pl.bsteps, &pl.bitems, &pl.boff);
  const SmallOffsets o = make_small_offsets(*cfg);
  pl.o = o;
  auto add = [&](int slot, int dst, int rows, int cols, int ld, int col0 = 0, int transpose = 0, int dst_ld = 0) {
    pl.cp.push_back({P[slot] ? P[slot] + col0 : nullptr, dst, rows, cols, ld, transpose, dst_ld ? dst_ld : rows});
  };
  for (int i = 0; i < 8; ++i) add(SPNERF_P_FC_W0 + 2 * i + 1, o.fc_b[i], 1, F, F);
  add(SPNERF_P_SIGMA_B, o.sigma_b, 1, 1, 1);
  add(SPNERF_P_FEATS_B, o.feats_b, 1, F, F);
  if (cfg->sem) {
    add(SPNERF_P_SEM0_B, o.sem0_b, 1, H, H);
    add(SPNERF_P_SEM2_W, o.sem2_w, cfg->num_sem_classes, H, H);
    add(SPNERF_P_SEM2_B, o.sem2_b, 1, cfg->num_sem_classes, cfg->num_sem_classes);
    add(SPNERF_P_SEM_EMB, o.emb, cfg->num_sem_classes + 1, cfg->emb_dim, cfg->emb_dim);
  }
  add(SPNERF_P_RGB0_B, o.rgb0_b, 1, H, H);
  add(SPNERF_P_RGB2_W, o.rgb2_w, 3, H, H);
  add(SPNERF_P_RGB2_B, o.rgb2_b, 1, 3, 3);
  add(SPNERF_P_SUN0_W + 1, o.sun0_b, 1, H, H);
  add(SPNERF_P_SUN0_W, o.sun0_wsun, H, 3, F + 3, F, 1);          // -> [3][256]
  add(SPNERF_P_SUN0_W + 3, o.sun2_b, 1, H, H);
  add(SPNERF_P_SUN0_W + 5, o.sun4_b, 1, H, H);
  add(SPNERF_P_SUN0_W + 6, o.sun6_w, 1, H, H);
  add(SPNERF_P_SUN0_W + 7, o.sun6_b, 1, 1, 1);
  if (cfg->beta) {
    add(SPNERF_P_BETA0_B, o.beta0_b, 1, H, H);
    add(SPNERF_P_BETA0_W, o.beta0_wt, H, cfg->t_dim, F + cfg->t_dim, F, 1);   // -> [t][256]
    add(SPNERF_P_BETA2_W, o.beta2_w, 1, H, H);
    add(SPNERF_P_BETA2_B, o.beta2_b, 1, 1, 1);
  }
  add(SPNERF_P_SKY0_W, o.sky0_w, H, 3, 3, 0, 1);                           // -> [3][256]
  add(SPNERF_P_SKY0_B, o.sky0_b, 1, H, H);
  add(SPNERF_P_SKY2_W, o.sky2_w, 3, H, H);
  add(SPNERF_P_SKY2_B, o.sky2_b, 1, 3, 3);
  // image of the shared-memory parameter region (net_plan.h kOffRgb2..): [j][4] / [j][8] / [j] / [j]
  add(SPNERF_P_RGB2_W, o.smallw, 3, H, H, 0, 1, 4);
  if (cfg->sem) add(SPNERF_P_SEM2_W, o.smallw + 1024, cfg->num_sem_classes, H, H, 0, 1, 8);
  add(SPNERF_P_SUN0_W + 6, o.smallw + 3072, 1, H, H);
  if (cfg->beta) add(SPNERF_P_BETA2_W, o.smallw + 3328, 1, H, H);
  pl.bytes_f = pl.f.items.size() * sizeof(PackItem);
  pl.bytes_b = pl.bitems.size() * sizeof(PackItem);
  pl.bytes_c = pl.cp.size() * sizeof(CopyItem);
  pl.bytes_a = pl.f.aux_items.size() * sizeof(AuxItem);
}
}  // namespace

extern "C" int64_t spnerf_net_pack_workspace_bytes(const SpnerfNetConfig* cfg) {
  if (validate(cfg)) return -1;
  const float* P[SPNERF_NUM_PARAMS] = {};
  PackPlan pl;
  make_pack_plan(cfg, P, pl);
  return (int64_t)(pl.bytes_f + pl.bytes_b + pl.bytes_c + pl.bytes_a + 256);
}

// Uploads the pack tables (they embed the parameter pointers) and the two step tables, and zeroes
// the padding of the operand buffers.  Once per (configuration, parameter pointers, buffers);
// synchronises the stream because the tables come from pageable host memory.
extern "C" int spnerf_net_prepare(const SpnerfNetConfig* cfg, const float* const* P, void* pack_ws,
                                  int64_t pack_ws_bytes, void* fwd_blob, void* bwd_blob, float* small,
                                  void* fwd_steps, void* bwd_steps, void* stream_) {
  int rc = validate(cfg);
  if (rc) return rc;
  if (!P || !pack_ws || !fwd_blob || !bwd_blob || !small || !fwd_steps || !bwd_steps) return SPNERF_ERR_BAD_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  PackPlan pl;
  make_pack_plan(cfg, P, pl);
  if ((int64_t)(pl.bytes_f + pl.bytes_b + pl.bytes_c + pl.bytes_a) > pack_ws_bytes) return SPNERF_ERR_WORKSPACE;
  uint8_t* sc = static_cast<uint8_t*>(pack_ws);
  cudaMemcpyAsync(sc, pl.f.items.data(), pl.bytes_f, cudaMemcpyHostToDevice, stream);
  cudaMemcpyAsync(sc + pl.bytes_f, pl.bitems.data(), pl.bytes_b, cudaMemcpyHostToDevice, stream);
  cudaMemcpyAsync(sc + pl.bytes_f + pl.bytes_b, pl.cp.data(), pl.bytes_c, cudaMemcpyHostToDevice, stream);
  cudaMemcpyAsync(sc + pl.bytes_f + pl.bytes_b + pl.bytes_c, pl.f.aux_items.data(), pl.bytes_a, cudaMemcpyHostToDevice,
                  stream);
  cudaMemcpyAsync(fwd_steps, pl.f.steps.data(), pl.f.steps.size() * sizeof(MmaStep), cudaMemcpyHostToDevice, stream);
  cudaMemcpyAsync(bwd_steps, pl.bsteps.data(), pl.bsteps.size() * sizeof(MmaStep), cudaMemcpyHostToDevice, stream);
  cudaMemsetAsync(fwd_blob, 0, (size_t)pl.f.off16 * 16, stream);
  cudaMemsetAsync(bwd_blob, 0, (size_t)pl.boff * 16, stream);
  cudaMemsetAsync(small, 0, (size_t)pl.o.total * sizeof(float), stream);
  cudaError_t e = cudaStreamSynchronize(stream);
  return e == cudaSuccess ? 0 : -(int)e;
}

// fp32 parameters -> packed operands (one launch, no host synchronisation).
extern "C" int spnerf_net_pack(const SpnerfNetConfig* cfg, const void* pack_ws, void* fwd_blob, void* bwd_blob,
                               float* small, void* stream_) {
  int rc = validate(cfg);
  if (rc) return rc;
  if (!pack_ws || !fwd_blob || !bwd_blob || !small) return SPNERF_ERR_BAD_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const float* P[SPNERF_NUM_PARAMS] = {};
  PackPlan pl;                       // host-side counts only
  make_pack_plan(cfg, P, pl);
  const uint8_t* sc = static_cast<const uint8_t*>(pack_ws);
  PackTables t;
  t.f = reinterpret_cast<const PackItem*>(sc); t.nf = (int)pl.f.items.size();
  t.b = reinterpret_cast<const PackItem*>(sc + pl.bytes_f); t.nb = (int)pl.bitems.size();
  t.c = reinterpret_cast<const CopyItem*>(sc + pl.bytes_f + pl.bytes_b); t.nc = (int)pl.cp.size();
  t.a = reinterpret_cast<const AuxItem*>(sc + pl.bytes_f + pl.bytes_b + pl.bytes_c); t.na = (int)pl.f.aux_items.size();
  pack_all_kernel<<<dim3((unsigned)(t.nf + t.nb + t.nc + t.na), 2), 256, 0, stream>>>(
      t, static_cast<uint8_t*>(fwd_blob), static_cast<uint8_t*>(bwd_blob), small);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}

// ------------------------------------------------------------------------------------------------
// sky colour per ray: sigmoid(W2 relu(W0 s + b0) + b2)   (models/spnerf.py:244-249, :355)
// one warp per ray, 8 hidden units per lane
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void sky_fwd_kernel(const float* __restrict__ small, SmallOffsets o, const float* __restrict__ rays,
                               int64_t n_rays, float* __restrict__ sky, float* __restrict__ hidden, int nh) {
  // nh = hidden units (feat / 2: 256 or 128), 32 per pass of the warp
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < n_rays; r += nwarps) {
    const float sx = rays[r * 11 + 8], sy = rays[r * 11 + 9], sz = rays[r * 11 + 10];
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (u * 32 >= nh) break;
      const int j = u * 32 + lane;
      float h = small[o.sky0_b + j];
      h = fmaf(small[o.sky0_w + j], sx, h);
      h = fmaf(small[o.sky0_w + nh + j], sy, h);
      h = fmaf(small[o.sky0_w + 2 * nh + j], sz, h);
      h = fmaxf(h, 0.f);
      if (hidden) hidden[r * nh + j] = h;
      acc0 = fmaf(small[o.sky2_w + j], h, acc0);
      acc1 = fmaf(small[o.sky2_w + nh + j], h, acc1);
      acc2 = fmaf(small[o.sky2_w + 2 * nh + j], h, acc2);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
      acc0 += __shfl_xor_sync(0xffffffffu, acc0, s);
      acc1 += __shfl_xor_sync(0xffffffffu, acc1, s);
      acc2 += __shfl_xor_sync(0xffffffffu, acc2, s);
    }
    if (lane < 3) {
      const float a = (lane == 0 ? acc0 : lane == 1 ? acc1 : acc2) + small[o.sky2_b + lane];
      sky[r * 3 + lane] = 1.f / (1.f + expf(-a));
    }
  }
}
}  // namespace

extern "C" int spnerf_sky_fwd(const float* small, const SpnerfNetConfig* cfg, const float* rays, int64_t n_rays,
                              float* sky, float* hidden, void* stream) {
  int rc = validate(cfg);
  if (rc) return rc;
  if (!small || !rays || !sky || n_rays < 0) return SPNERF_ERR_BAD_ARG;
  if (n_rays == 0) return 0;
  const int threads = 256;
  const int64_t blocks = (n_rays * 32 + threads - 1) / threads;
  sky_fwd_kernel<<<(unsigned)(blocks > 148 * 16 ? 148 * 16 : blocks), threads, 0, static_cast<cudaStream_t>(stream)>>>(
      small, make_small_offsets(*cfg), rays, n_rays, sky, hidden, cfg->feat / 2);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}

// ------------------------------------------------------------------------------------------------
// backward-data order; must match the phase sequence of mlp_bwd.cu.  Gradient tiles are the A
// operands, transposed weights the B operands: g_in[n] = sum_k G[k] * W[k][n].
// ------------------------------------------------------------------------------------------------
void build_backward(const SpnerfNetConfig& c, const float* const* P, std::vector<MmaStep>& steps,
                    std::vector<PackItem>* items, uint32_t* off16) {
  Builder b;
  const NetDims d = make_dims(c);
  const int base = c.mapping ? 60 : 3;
  const int F = c.feat, H = F / 2, Q = H / 2;      // trunk width, head width, head chunk width
  const int SF = F / 64, SH = H / 64;              // K slabs of a trunk-wide / head-wide gradient tile
  const int parts = kSplitBwd ? 4 : 2, PW = 2 * H / parts;      // chunks of a split-capable full-width phase
  using Src = Builder::Src;
  auto ks = [](int first_slab, int nslabs) {
    std::vector<Src> v;
    for (int k = 0; k < nslabs; ++k) v.push_back(Src{first_slab + k, 64 * k, 4, 0});
    return v;
  };
  // sun_v_net.4 and .2 (256x256): G_s3 -> g_s2 -> g_s1
  for (int j = 2; j >= 1; --j) {      // two 128-wide chunks, one per issuer lane
    for (int g = 0; g < 2; ++g) {
      b.chunk(P[SPNERF_P_SUN0_W + 2 * j], H, H, g * Q, Q, g * Q, ks(0, SH), true);
      b.merge_last_chunk(SH >= 4 ? 2 : SH);
    }
    b.end_phase();
  }
  // g_f = G_s1 * W_sun0[:, :512] + G_r1 * W_rgb0 (+ G_b1 * W_beta0[:, :512] in a second phase)
  for (int g = 0; g < 2; ++g) {
    std::vector<Src> v;
    for (int k = 0; k < SH; ++k) v.push_back(Src{k, 64 * k, 4, 0, P[SPNERF_P_SUN0_W], H, F + 3});
    for (int k = 0; k < SH; ++k) v.push_back(Src{SH + k, 64 * k, 4, 0, P[SPNERF_P_RGB0_W], H, F});
    b.chunk(nullptr, 0, 0, g * H, H, g * H, v, true);
    if (F == 256) b.merge_last_chunk(2);        // 128-wide chunks: two K slabs per ring item
  }
  b.end_phase();
  if (c.beta) {
    for (int g = 0; g < 2; ++g)
      b.chunk(P[SPNERF_P_BETA0_W], H, F + c.t_dim, g * H, H, g * H, ks(0, SH), true, true);
    b.end_phase();
  }
  // g_h = g_f * W_feats (+ G_sem1 * W_sem0 + g_sigma_pre * W_sigma accumulated in the next phase)
  for (int g = 0; g < 2; ++g) {
    b.chunk(P[SPNERF_P_FEATS_W], F, F, g * H, H, g * H, ks(0, SF), true);
    if (F == 256) b.merge_last_chunk(2);
  }
  b.end_phase();
  for (int g = 0; g < parts; ++g) {
    std::vector<Src> v;
    if (c.sem)
      for (int k = 0; k < SH; ++k) v.push_back(Src{k, 64 * k, 4, 0, P[SPNERF_P_SEM0_W], H, F});
    v.push_back(Src{SH, 0, 1, 0, P[SPNERF_P_SIGMA_W], 1, F});       // the sigma column sits in the slab after the semantic hidden gradient
    b.chunk(nullptr, 0, 0, g * PW, PW, g * PW, v, true, true);
    if (c.sem && 256 / PW > 1 && SH % (256 / PW) == 0) b.merge_last_chunk(256 / PW);   // the sigma step stays on its own
  }
  b.end_phase(true, kSplitBwd ? SH : 0);      // the eight phases whose epilogue is the trunk loop of mlp_bwd.cu
  // trunk, layers 7..1; the label-embedding columns of the skip layer and of layer 0 get their own
  // 16-wide mini phases (only when the embedding exists)
  for (int L = 7; L >= 1; --L) {
    const bool skip = (L == c.skip_layer);
    const int cols = F + (skip ? d.in_dim : 0);
    if (skip && c.sem) {
      b.chunk(P[SPNERF_P_FC_W0 + 2 * L], F, cols, F + base, 16, 0, ks(0, SF), true);
      b.merge_last_chunk(SF);
      b.end_phase();
    }
    for (int g = 0; g < parts; ++g) {
      b.chunk(P[SPNERF_P_FC_W0 + 2 * L], F, cols, g * PW, PW, g * PW, ks(0, SF), true);
      if (256 / PW > 1) b.merge_last_chunk(256 / PW);
    }
    b.end_phase(true, kSplitBwd ? SH : 0);
  }
  if (c.sem) {
    b.chunk(P[SPNERF_P_FC_W0], F, d.in_dim, base, 16, 0, ks(0, SF), true);
    b.merge_last_chunk(SF);
    b.end_phase();
  }
  steps = b.steps;
  if (items) *items = b.items;
  *off16 = b.off16;
}

// ------------------------------------------------------------------------------------------------
// sky colour backward (adjoint of sky_fwd_kernel): one warp per ray, 8 hidden units per lane,
// register accumulation over the warp's rays, then shared-memory and global atomics.
// ------------------------------------------------------------------------------------------------
namespace {
__global__ void __launch_bounds__(256) sky_bwd_kernel(const float* __restrict__ small, SmallOffsets o,
                                                      const float* __restrict__ rays, const float* __restrict__ sky,
                                                      const float* __restrict__ hidden, const float* __restrict__ g_sky,
                                                      int64_t n_rays, float* g_w0, float* g_b0, float* g_w2,
                                                      float* g_b2, int nh) {
  __shared__ float acc[3 * kHalf + kHalf + 3 * kHalf + 4];   // w0 [nh][3] | b0 [nh] | w2 [3][nh] | b2  (nh = feat / 2 <= 256)
  for (int i = threadIdx.x; i < 7 * kHalf + 4; i += blockDim.x) acc[i] = 0.f;   // (whole array)
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float aw0[8][3], ab0[8], aw2[3][8], ab2[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    ab0[u] = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) { aw0[u][c] = 0.f; aw2[c][u] = 0.f; }
  }
  for (int64_t r = warp; r < n_rays; r += nwarps) {
    float gp[3], s[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float y = sky[r * 3 + c];
      gp[c] = g_sky[r * 3 + c] * y * (1.f - y);
      s[c] = rays[r * 11 + 8 + c];
      ab2[c] += gp[c];
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (u * 32 >= nh) break;
      const int j = u * 32 + lane;
      const float h = hidden[r * nh + j];
      float gh = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        aw2[c][u] = fmaf(gp[c], h, aw2[c][u]);
        gh = fmaf(gp[c], small[o.sky2_w + c * nh + j], gh);
      }
      gh = h > 0.f ? gh : 0.f;
      ab0[u] += gh;
#pragma unroll
      for (int c = 0; c < 3; ++c) aw0[u][c] = fmaf(gh, s[c], aw0[u][c]);
    }
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    if (u * 32 >= nh) break;
    const int j = u * 32 + lane;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      atomicAdd(&acc[j * 3 + c], aw0[u][c]);
      atomicAdd(&acc[4 * nh + c * nh + j], aw2[c][u]);
    }
    atomicAdd(&acc[3 * nh + j], ab0[u]);
  }
  if (lane == 0)
    for (int c = 0; c < 3; ++c) atomicAdd(&acc[7 * nh + c], ab2[c]);
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * nh; i += blockDim.x) {
    atomicAdd(g_w0 + i, acc[i]);
    atomicAdd(g_w2 + i, acc[4 * nh + i]);
  }
  for (int i = threadIdx.x; i < nh; i += blockDim.x) atomicAdd(g_b0 + i, acc[3 * nh + i]);
  if (threadIdx.x < 3) atomicAdd(g_b2 + threadIdx.x, acc[7 * nh + threadIdx.x]);
}
}  // namespace

extern "C" int spnerf_sky_bwd(const float* small, const SpnerfNetConfig* cfg, const float* rays, const float* sky,
                              const float* hidden, const float* g_sky, int64_t n_rays, float* g_w0, float* g_b0,
                              float* g_w2, float* g_b2, void* stream) {
  int rc = validate(cfg);
  if (rc) return rc;
  if (!small || !rays || !sky || !hidden || !g_sky || !g_w0 || !g_b0 || !g_w2 || !g_b2) return SPNERF_ERR_BAD_ARG;
  if (n_rays <= 0) return n_rays == 0 ? 0 : SPNERF_ERR_BAD_ARG;
  const int64_t blocks = (n_rays + 63) / 64;     // >= 8 rays per warp
  sky_bwd_kernel<<<(unsigned)(blocks > 148 ? 148 : blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      small, make_small_offsets(*cfg), rays, sky, hidden, g_sky, n_rays, g_w0, g_b0, g_w2, g_b2, cfg->feat / 2);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}
