// Volume integration along each ray: alpha compositing with shadow-aware shading, depth and mean
// semantic logits, and its closed-form adjoint.
//
// Replaces models/spnerf.py:109-157 (the ~20 elementwise / cumprod / sum kernels the reference
// launches per pass) and the autograd graph behind them.
//
// One warp per ray.  The ray's network rows (n_samples x n_out fp32, contiguous) are pulled into
// shared memory with 16-byte coalesced loads, each lane then owns a contiguous block of samples:
// transmittance is a warp product-scan (forward), the adjoint a warp reverse sum-scan (backward).
// HBM-bound: forward moves 4*N*(n_out+1 read + 2 write) + 4*(3+1+C) bytes per ray.
#include <cuda_runtime.h>
#include <stdint.h>
#include <algorithm>
#include "../../include/spnerf_b200.h"
#include "sm100.cuh"

namespace {
using sm100::mbar_init; using sm100::mbar_wait; using sm100::mbar_expect_tx; using sm100::bulk_g2s;
using sm100::fence_proxy_async_smem; using sm100::fence_mbar_init;

constexpr int kMaxPerLane = 8;      // samples per lane
constexpr int kSkew = 8;            // floats between the rows of the rays of a group (bank skew)

// Layout of the work: a warp integrates G rays at a time, LPR = 32 / G lanes per ray, each lane a
// contiguous block of samples; transmittance is a product scan over the ray's lanes (forward), the
// adjoint a reverse sum scan (backward).  The rows / depths of the G consecutive rays arrive through
// the bulk-copy engine (one copy per tensor and group) into a 2-deep per-warp ring, so the next group
// is in flight while this one is integrated.  G = 4 (n <= 64) or 2 (n <= 128) need 16-byte multiples
// (n * n_out % 4 == 0, n % 4 == 0); otherwise G = 1 and the warp copies one ray at a time itself.
template <int LPR>
__device__ __forceinline__ float seg_excl_scan_mul(float v, int sl) {
  float inc = v;
#pragma unroll
  for (int s = 1; s < LPR; s <<= 1) {
    const float o = __shfl_up_sync(0xffffffffu, inc, s, LPR);
    if (sl >= s) inc *= o;
  }
  const float ex = __shfl_up_sync(0xffffffffu, inc, 1, LPR);
  return sl == 0 ? 1.f : ex;
}
// inclusive product over the ray's lanes; ex = product over lanes < sl, tot = product over all of them
template <int LPR>
__device__ __forceinline__ void seg_scan_mul(float v, int sl, float& ex, float& tot) {
  float inc = v;
#pragma unroll
  for (int s = 1; s < LPR; s <<= 1) {
    const float o = __shfl_up_sync(0xffffffffu, inc, s, LPR);
    if (sl >= s) inc *= o;
  }
  ex = __shfl_up_sync(0xffffffffu, inc, 1, LPR);
  if (sl == 0) ex = 1.f;
  tot = __shfl_sync(0xffffffffu, inc, LPR - 1, LPR);
}
// inclusive suffix sum; ex = sum over lanes > sl, tot = sum over all of them
template <int LPR>
__device__ __forceinline__ void seg_suffix_sum(float v, int sl, float& ex, float& tot) {
  float inc = v;
#pragma unroll
  for (int s = 1; s < LPR; s <<= 1) {
    const float o = __shfl_down_sync(0xffffffffu, inc, s, LPR);
    if (sl + s < LPR) inc += o;
  }
  ex = __shfl_down_sync(0xffffffffu, inc, 1, LPR);
  if (sl == LPR - 1) ex = 0.f;
  tot = __shfl_sync(0xffffffffu, inc, 0, LPR);
}
template <int LPR>
__device__ __forceinline__ float seg_sum(float v) {
#pragma unroll
  for (int s = LPR / 2; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s, LPR);
  return v;
}
// exclusive suffix sum inside the ray's lanes: result(sl) = sum of v over lanes > sl
template <int LPR>
__device__ __forceinline__ float seg_excl_suffix_sum(float v, int sl) {
  float inc = v;
#pragma unroll
  for (int s = 1; s < LPR; s <<= 1) {
    const float o = __shfl_down_sync(0xffffffffu, inc, s, LPR);
    if (sl + s < LPR) inc += o;
  }
  const float ex = __shfl_down_sync(0xffffffffu, inc, 1, LPR);
  return sl == LPR - 1 ? 0.f : ex;
}

// coalesced copy of `n` floats global -> shared for one warp (16-byte vectors when aligned)
__device__ __forceinline__ void warp_load(float* dst, const float* __restrict__ src, int n, int lane) {
  if ((reinterpret_cast<uintptr_t>(src) & 15) == 0 && (n & 3) == 0) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    for (int i = lane; i < n / 4; i += 32) d4[i] = __ldcs(s4 + i);
  } else {
    for (int i = lane; i < n; i += 32) dst[i] = __ldcs(src + i);
  }
}
__device__ __forceinline__ void warp_store(float* __restrict__ dst, const float* src, int n, int lane) {
  if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0 && (n & 3) == 0) {
    float4* d4 = reinterpret_cast<float4*>(dst);
    const float4* s4 = reinterpret_cast<const float4*>(src);
    for (int i = lane; i < n / 4; i += 32) __stcs(d4 + i, s4[i]);
  } else {
    for (int i = lane; i < n; i += 32) __stcs(dst + i, src[i]);
  }
}

struct FwdP {
  const float* out; const float* z; const float* noise; float noise_std;
  int64_t n_rays; int n; int n_out; int col_sem; int n_sem;
  float* weights; float* trans; float* rgb; float* rgb_raw; float* depth; float* sem;
  float* aux; int* sem_argmax; int col_beta;      // per-ray sums for image export (eval.py:75-101), optional
};

// first maximum, as torch.argmax (eval.py:63)
template <int NS>
__device__ __forceinline__ int argmax_first(const float (&v)[NS > 0 ? NS : 1], int n) {
  int best = 0;
  float bv = v[0];
#pragma unroll
  for (int c = 1; c < NS; ++c)
    if (c < n && v[c] > bv) { bv = v[c]; best = c; }
  return best;
}

template <int G>
__global__ void composite_fwd_kernel(const FwdP p) {
  extern __shared__ __align__(16) float sm[];
  constexpr bool PIPE = G > 1;
  constexpr int LPR = 32 / G;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int sub = lane / LPR, sl = lane % LPR;
  // Lane sl of a ray owns samples sl, sl + LPR, sl + 2 LPR, ...: consecutive lanes read consecutive rows
  // (stride n_out floats), and the rays of a group are skewed by kSkew floats, so the shared-memory
  // reads of a warp fall into distinct banks (a blocked assignment was 8-way conflicted).
  const int row_f = ((p.n * p.n_out + 3) & ~3) + (PIPE ? kSkew : 0);
  const int buf_f = G * (row_f + p.n);                     // one ring slot: rows [G][n][n_out] (+skew) + depths [G][n]
  float* base = sm + (size_t)wib * ((PIPE ? 2 : 1) * buf_f + G * p.n + 4);
  float* wstage = base + (PIPE ? 2 : 1) * buf_f;           // [G][n] weights staging
  uint64_t* bars = reinterpret_cast<uint64_t*>(wstage + G * p.n);
  const int per = (p.n + LPR - 1) / LPR;
  const int64_t warp0 = (int64_t)blockIdx.x * wpb + wib, nw = (int64_t)gridDim.x * wpb;
  const int64_t n_groups = (p.n_rays + G - 1) / G;
  auto prefetch = [&](int64_t g, int slot) {
    const int64_t r0 = g * G;
    const uint32_t cnt = (uint32_t)min((int64_t)G, p.n_rays - r0);
    const uint32_t row_bytes = (uint32_t)(p.n * p.n_out) * 4u, z_bytes = cnt * (uint32_t)p.n * 4u;
    mbar_expect_tx(&bars[slot], cnt * row_bytes + z_bytes);
    for (uint32_t q = 0; q < cnt; ++q)
      bulk_g2s(base + slot * buf_f + q * row_f, p.out + (r0 + q) * p.n * p.n_out, row_bytes, &bars[slot]);
    bulk_g2s(base + slot * buf_f + G * row_f, p.z + r0 * p.n, z_bytes, &bars[slot]);
  };
  if (PIPE) {
    if (lane == 0) {
      mbar_init(&bars[0], 1); mbar_init(&bars[1], 1);
      fence_mbar_init();
      if (warp0 < n_groups) prefetch(warp0, 0);
    }
    __syncwarp();
  }
  int t = 0;
  for (int64_t g = warp0; g < n_groups; g += nw, ++t) {
    float* slot = base + (PIPE ? (t & 1) : 0) * buf_f;
    const int64_t r0 = g * G;
    const int cnt = (int)min((int64_t)G, p.n_rays - r0);
    const int64_t r = r0 + sub;
    const bool active = sub < cnt;
    float* rows = slot + sub * row_f;                      // [n][n_out]
    float* zs = slot + G * row_f + sub * p.n;              // [n]  (depths, then reused for T)
    float* ws = wstage + sub * p.n;
    if (PIPE) {
      if (lane == 0 && g + nw < n_groups) prefetch(g + nw, (t + 1) & 1);
      mbar_wait(&bars[t & 1], (uint32_t)(t >> 1) & 1u, 50);
    } else {
      warp_load(rows, p.out + r * p.n * p.n_out, p.n * p.n_out, lane);
      warp_load(zs, p.z + r * p.n, p.n, lane);
      __syncwarp();
    }
    // lane owns samples sl + LPR * k
    float alpha[kMaxPerLane], keep[kMaxPerLane];
#pragma unroll
    for (int k = 0; k < kMaxPerLane; ++k) {
      const int i = sl + LPR * k;
      alpha[k] = 0.f; keep[k] = 1.f;
      if (k < per && i < p.n) {
        const float delta = (i + 1 < p.n) ? zs[i + 1] - zs[i] : 1e10f;                 // spnerf.py:116-118
        float s = rows[i * p.n_out + 3];
        if (p.noise && active) s += p.noise[r * p.n + i] * p.noise_std;                // :121-122
        alpha[k] = 1.f - expf(-delta * fmaxf(s, 0.f));                                 // :123
        keep[k] = 1.f - alpha[k] + 1e-10f;                                             // :126
      }
    }
    float acc_d = 0.f, acc_c[3] = {0.f, 0.f, 0.f}, acc_s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float acc_a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // albedo(3), sun, sky(3), beta
    float carry = 1.f;                                      // product over the samples of earlier rounds
#pragma unroll
    for (int k = 0; k < kMaxPerLane; ++k) {
      const int i = sl + LPR * k;
      if (k < per) {                                        // warp-uniform
        float ex, tot;
        seg_scan_mul<LPR>(keep[k], sl, ex, tot);            // :127 (exclusive cumprod), one round of LPR samples
        if (i < p.n) {
          const float T = carry * ex;
          const float w = alpha[k] * T;                                                // :128
          const float* o = rows + i * p.n_out;
          const float sv = o[4];
          acc_d = fmaf(w, zs[i], acc_d);                                               // :131
#pragma unroll
          for (int c = 0; c < 3; ++c) acc_c[c] = fmaf(w * o[c], sv + (1.f - sv) * o[5 + c], acc_c[c]);   // :132-133
          for (int c = 0; c < p.n_sem; ++c) acc_s[c] += o[p.col_sem + c];
          if (p.aux) {
#pragma unroll
            for (int c = 0; c < 3; ++c) { acc_a[c] = fmaf(w, o[c], acc_a[c]); acc_a[4 + c] = fmaf(w, o[5 + c], acc_a[4 + c]); }
            acc_a[3] = fmaf(w, sv, acc_a[3]);
            if (p.col_beta >= 0) acc_a[7] = fmaf(w, o[p.col_beta], acc_a[7]);
          }
          ws[i] = w;
          zs[i] = T;   // every depth difference was formed in the first loop (the scans order the loops)
        }
        carry *= tot;
      }
    }
    // The slot was written through the generic proxy (T over the depths) and is refilled by the copy
    // engine two iterations later: order those writes for the async proxy here, BEFORE this
    // iteration's global stores are issued (the fence's membar would otherwise wait for them).
    if (PIPE) fence_proxy_async_smem();
    __syncwarp();
    if (p.weights) warp_store(p.weights + r0 * p.n, wstage, cnt * p.n, lane);
    if (p.trans) warp_store(p.trans + r0 * p.n, slot + G * row_f, cnt * p.n, lane);
    acc_d = seg_sum<LPR>(acc_d);
#pragma unroll
    for (int c = 0; c < 3; ++c) acc_c[c] = seg_sum<LPR>(acc_c[c]);
    for (int c = 0; c < p.n_sem; ++c) acc_s[c] = seg_sum<LPR>(acc_s[c]);
    if (p.aux) {
#pragma unroll
      for (int c = 0; c < 8; ++c) acc_a[c] = seg_sum<LPR>(acc_a[c]);
      if (sl == 0 && active) {
#pragma unroll
        for (int c = 0; c < 8; ++c) p.aux[r * 8 + c] = acc_a[c];
      }
    }
    if (p.sem_argmax && sl == 0 && active) p.sem_argmax[r] = argmax_first<8>(acc_s, p.n_sem);
    if (sl == 0 && active) {
      p.depth[r] = acc_d;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (p.rgb_raw) p.rgb_raw[r * 3 + c] = acc_c[c];
        p.rgb[r * 3 + c] = fminf(fmaxf(acc_c[c], 0.f), 1.f);                           // :134
      }
      for (int c = 0; c < p.n_sem; ++c) p.sem[r * p.n_sem + c] = acc_s[c] / (float)p.n;   // :156 (plain mean)
    }
    __syncwarp();
  }
}

struct BwdP {
  const float* out; const float* z; const float* noise; float noise_std;
  const float* weights; const float* trans; const float* rgb_raw;
  const float* g_rgb; const float* g_depth; const float* g_sem; const float* g_w; const float* g_t;
  const float* g_out_ext;
  int64_t n_rays; int n; int n_out; int col_sem; int n_sem;
  float* g_out; float* g_sky_ray; unsigned int* absmax_bits;
};

template <int G>
__global__ void composite_bwd_kernel(const BwdP p) {
  extern __shared__ __align__(16) float sm[];
  constexpr bool PIPE = G > 1;
  constexpr int LPR = 32 / G;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int sub = lane / LPR, sl = lane % LPR;
  const int row_f = ((p.n * p.n_out + 3) & ~3) + (PIPE ? kSkew : 0);   // skewed per ray, see the forward kernel
  const int ext_f = p.g_out_ext ? row_f : 0;
  const int buf_f = G * (row_f + ext_f + 3 * p.n);          // one ring slot: rows | external gradient rows | z | w | T
  float* base = sm + (size_t)wib * ((PIPE ? 2 : 1) * buf_f + 4);
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + (PIPE ? 2 : 1) * buf_f);
  const int per = (p.n + LPR - 1) / LPR;
  const int64_t warp0 = (int64_t)blockIdx.x * wpb + wib, nw = (int64_t)gridDim.x * wpb;
  const int64_t n_groups = (p.n_rays + G - 1) / G;
  auto prefetch = [&](int64_t g, int slot) {
    const int64_t r0 = g * G;
    const uint32_t cnt = (uint32_t)min((int64_t)G, p.n_rays - r0);
    const uint32_t row_bytes = (uint32_t)(p.n * p.n_out) * 4u, z_bytes = cnt * (uint32_t)p.n * 4u;
    float* b = base + slot * buf_f;
    mbar_expect_tx(&bars[slot], cnt * row_bytes * (p.g_out_ext ? 2u : 1u) + 3u * z_bytes);
    for (uint32_t q = 0; q < cnt; ++q) {
      bulk_g2s(b + q * row_f, p.out + (r0 + q) * p.n * p.n_out, row_bytes, &bars[slot]);
      if (p.g_out_ext) bulk_g2s(b + (G + q) * row_f, p.g_out_ext + (r0 + q) * p.n * p.n_out, row_bytes, &bars[slot]);
    }
    float* zb = b + G * (row_f + ext_f);
    bulk_g2s(zb, p.z + r0 * p.n, z_bytes, &bars[slot]);
    bulk_g2s(zb + G * p.n, p.weights + r0 * p.n, z_bytes, &bars[slot]);
    bulk_g2s(zb + 2 * G * p.n, p.trans + r0 * p.n, z_bytes, &bars[slot]);
  };
  if (PIPE) {
    if (lane == 0) {
      mbar_init(&bars[0], 1); mbar_init(&bars[1], 1);
      fence_mbar_init();
      if (warp0 < n_groups) prefetch(warp0, 0);
    }
    __syncwarp();
  }
  float amax = 0.f;
  int t = 0;
  for (int64_t g = warp0; g < n_groups; g += nw, ++t) {
    float* slot = base + (PIPE ? (t & 1) : 0) * buf_f;
    const int64_t r0 = g * G;
    const int cnt = (int)min((int64_t)G, p.n_rays - r0);
    const int64_t r = r0 + sub;
    const bool active = sub < cnt;
    float* rows = slot + sub * row_f;                       // network rows, overwritten by their gradients
    float* ext = slot + G * row_f + sub * row_f;            // external gradient rows (optional)
    float* zs = slot + G * (row_f + ext_f) + sub * p.n;
    float* ws = zs + G * p.n;
    float* ts = ws + G * p.n;
    if (PIPE) {
      if (lane == 0 && g + nw < n_groups) prefetch(g + nw, (t + 1) & 1);
      mbar_wait(&bars[t & 1], (uint32_t)(t >> 1) & 1u, 51);
    } else {
      warp_load(rows, p.out + r * p.n * p.n_out, p.n * p.n_out, lane);
      if (p.g_out_ext) warp_load(ext, p.g_out_ext + r * p.n * p.n_out, p.n * p.n_out, lane);
      warp_load(zs, p.z + r * p.n, p.n, lane);
      warp_load(ws, p.weights + r * p.n, p.n, lane);
      warp_load(ts, p.trans + r * p.n, p.n, lane);
      __syncwarp();
    }
    float gh[3] = {0.f, 0.f, 0.f};
    if (p.g_rgb && active) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float raw = p.rgb_raw[r * 3 + c];
        gh[c] = (raw >= 0.f && raw <= 1.f) ? p.g_rgb[r * 3 + c] : 0.f;                 // clamp adjoint
      }
    }
    const float gd = (p.g_depth && active) ? p.g_depth[r] : 0.f;
    // lane owns samples sl + LPR * k
    float Gd[kMaxPerLane], Sv[kMaxPerLane];
#pragma unroll
    for (int k = 0; k < kMaxPerLane; ++k) {
      const int i = sl + LPR * k;
      Gd[k] = 0.f; Sv[k] = 0.f;
      if (k < per && i < p.n) {
        const float* o = rows + i * p.n_out;
        const float sv = o[4];
        float gg = gd * zs[i];
#pragma unroll
        for (int c = 0; c < 3; ++c) gg = fmaf(o[c] * (sv + (1.f - sv) * o[5 + c]), gh[c], gg);
        if (p.g_w && active) gg += p.g_w[r * p.n + i];
        Gd[k] = gg;                                                                     // dL/dw_i
        Sv[k] = gg * ws[i] + ((p.g_t && active) ? p.g_t[r * p.n + i] * ts[i] : 0.f);
      }
    }
    float gsky[3] = {0.f, 0.f, 0.f};
    float carry = 0.f;                                       // sum of S over the samples of later rounds
    // rounds backwards: R = sum of S over later samples (exclusive suffix sum along the ray)
#pragma unroll
    for (int k = kMaxPerLane - 1; k >= 0; --k) {
      const int i = sl + LPR * k;
      if (k < per) {                                         // warp-uniform
        float ex, tot;
        seg_suffix_sum<LPR>(Sv[k], sl, ex, tot);
        if (i < p.n) {
          const float R = carry + ex;
          float* o = rows + i * p.n_out;
          const float w = ws[i], T = ts[i];
          const float delta = (i + 1 < p.n) ? zs[i + 1] - zs[i] : 1e10f;
          float s = o[3];
          if (p.noise && active) s += p.noise[r * p.n + i] * p.noise_std;
          const float e = expf(-delta * fmaxf(s, 0.f));          // 1 - alpha
          const float dalpha = Gd[k] * T - R / (e + 1e-10f);
          const float dsigma = (s > 0.f) ? dalpha * delta * e : 0.f;
          const float sv = o[4];
          float go[8];
          float dsv = 0.f;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float irr = sv + (1.f - sv) * o[5 + c];
            go[c] = w * irr * gh[c];                              // d albedo
            const float dirr = w * o[c] * gh[c];
            dsv = fmaf(dirr, 1.f - o[5 + c], dsv);
            go[5 + c] = dirr * (1.f - sv);                        // d sky
          }
          go[3] = dsigma;
          go[4] = dsv;
          const float* e_row = ext + i * p.n_out;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float v = go[c] + (p.g_out_ext ? e_row[c] : 0.f);
            o[c] = v;
            if (active) amax = fmaxf(amax, fabsf(v));
          }
#pragma unroll
          for (int c = 0; c < 3; ++c) gsky[c] += o[5 + c];
          for (int c = 8; c < p.n_out; ++c) {
            float v = p.g_out_ext ? e_row[c] : 0.f;
            if (c >= p.col_sem && c < p.col_sem + p.n_sem && p.g_sem && active)
              v += p.g_sem[r * p.n_sem + (c - p.col_sem)] / (float)p.n;
            o[c] = v;
            if (active) amax = fmaxf(amax, fabsf(v));
          }
        }
        carry += tot;
      }
    }
    if (PIPE) fence_proxy_async_smem();     // see the forward kernel
    __syncwarp();
    for (int q = 0; q < cnt; ++q)
      warp_store(p.g_out + (r0 + q) * p.n * p.n_out, slot + q * row_f, p.n * p.n_out, lane);
    if (p.g_sky_ray) {
#pragma unroll
      for (int c = 0; c < 3; ++c) gsky[c] = seg_sum<LPR>(gsky[c]);
      if (sl < 3 && active) p.g_sky_ray[r * 3 + sl] = sl == 0 ? gsky[0] : sl == 1 ? gsky[1] : gsky[2];
    }
    __syncwarp();
  }
  if (p.absmax_bits) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, s));
    if (lane == 0 && amax > 0.f) atomicMax(p.absmax_bits, __float_as_uint(amax));   // positive floats order as uints
  }
}

// ------------------------------------------------------------------------------------------------
// Specialised kernels for the shapes the renderer actually runs (n = 8 * 32 / G samples, i.e. 64 at
// G = 4 and 128 at G = 2, and the network's four row widths): the same work split and ring as the
// generic kernels above, with every loop bound, row width and column index a compile-time constant,
// so the per-sample shared-memory reads use immediate offsets and all per-lane state stays in
// registers.  The generic kernels spent ~200 instructions per sample (runtime-indexed accumulators in
// local memory, address arithmetic) and were issue/latency bound at 12 % occupancy.
// ------------------------------------------------------------------------------------------------
template <int G, int NO, int NS, bool AUX>
__global__ void __launch_bounds__(128) composite_fwd_fast(const FwdP p) {
  extern __shared__ __align__(16) float sm[];
  constexpr int LPR = 32 / G, N = LPR * 8, ROW = N * NO + kSkew, BUF = G * (ROW + N), CS = NO - NS;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int sub = lane / LPR, sl = lane % LPR;
  float* base = sm + (size_t)wib * (2 * BUF + G * N + 4);
  float* wstage = base + 2 * BUF;
  uint64_t* bars = reinterpret_cast<uint64_t*>(wstage + G * N);
  const int64_t warp0 = (int64_t)blockIdx.x * wpb + wib, nw = (int64_t)gridDim.x * wpb;
  const int64_t n_groups = (p.n_rays + G - 1) / G;
  auto prefetch = [&](int64_t g, int slot) {
    const int64_t r0 = g * G;
    const uint32_t cnt = (uint32_t)min((int64_t)G, p.n_rays - r0);
    constexpr uint32_t row_bytes = (uint32_t)(N * NO) * 4u;
    mbar_expect_tx(&bars[slot], cnt * (row_bytes + N * 4u));
    for (uint32_t q = 0; q < cnt; ++q)
      bulk_g2s(base + slot * BUF + q * ROW, p.out + (r0 + q) * (N * NO), row_bytes, &bars[slot]);
    bulk_g2s(base + slot * BUF + G * ROW, p.z + r0 * N, cnt * N * 4u, &bars[slot]);
  };
  if (lane == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1);
    fence_mbar_init();
    if (warp0 < n_groups) prefetch(warp0, 0);
  }
  __syncwarp();
  int t = 0;
  for (int64_t g = warp0; g < n_groups; g += nw, ++t) {
    float* slot = base + (t & 1) * BUF;
    const int64_t r0 = g * G;
    const int cnt = (int)min((int64_t)G, p.n_rays - r0);
    const int64_t r = r0 + sub;
    const bool active = sub < cnt;
    const float* o = slot + sub * ROW + sl * NO;            // row of sample sl; sample sl + LPR k is LPR k NO floats on
    float* zs = slot + G * ROW + sub * N + sl;              // depths, overwritten by T
    float* ws = wstage + sub * N + sl;
    if (lane == 0 && g + nw < n_groups) prefetch(g + nw, (t + 1) & 1);
    mbar_wait(&bars[t & 1], (uint32_t)(t >> 1) & 1u, 50);
    float alpha[8], keep[8], zv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      zv[k] = zs[LPR * k];
      float delta = zs[LPR * k + 1] - zv[k];                // the word after the ray's last depth is never used:
      if (k == 7 && sl == LPR - 1) delta = 1e10f;           // spnerf.py:116-118
      float s = o[LPR * k * NO + 3];
      if (p.noise && active) s += p.noise[r * N + sl + LPR * k] * p.noise_std;          // :121-122
      alpha[k] = 1.f - __expf(-delta * fmaxf(s, 0.f));                                   // :123
      keep[k] = 1.f - alpha[k] + 1e-10f;                                                 // :126
    }
    float acc_d = 0.f, acc_c[3] = {0.f, 0.f, 0.f}, acc_s[NS > 0 ? NS : 1];
    float acc_a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // albedo(3), sun, sky(3), beta
    acc_s[0] = 0.f;
#pragma unroll
    for (int c = 0; c < NS; ++c) acc_s[c] = 0.f;
    float carry = 1.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float ex, tot;
      seg_scan_mul<LPR>(keep[k], sl, ex, tot);              // :127 (exclusive cumprod), one round of LPR samples
      const float T = carry * ex;
      const float w = alpha[k] * T;                                                      // :128
      const float* ok = o + LPR * k * NO;
      const float sv = ok[4];
      acc_d = fmaf(w, zv[k], acc_d);                                                     // :131
#pragma unroll
      for (int c = 0; c < 3; ++c) acc_c[c] = fmaf(w * ok[c], sv + (1.f - sv) * ok[5 + c], acc_c[c]);   // :132-133
#pragma unroll
      for (int c = 0; c < NS; ++c) acc_s[c] += ok[CS + c];
      if (AUX) {
#pragma unroll
        for (int c = 0; c < 3; ++c) { acc_a[c] = fmaf(w, ok[c], acc_a[c]); acc_a[4 + c] = fmaf(w, ok[5 + c], acc_a[4 + c]); }
        acc_a[3] = fmaf(w, sv, acc_a[3]);
        if (NO - NS == 9) acc_a[7] = fmaf(w, ok[8], acc_a[7]);      // beta column (models/spnerf.py:148-152)
      }
      ws[LPR * k] = w;
      zs[LPR * k] = T;      // every depth difference was formed in the first loop (the scans order the loops)
      carry *= tot;
    }
    fence_proxy_async_smem();     // see the generic kernel
    __syncwarp();
    if (p.weights) {        // both or neither (host check)
      if (cnt == G) {
        const float4* w4 = reinterpret_cast<const float4*>(wstage);
        const float4* t4 = reinterpret_cast<const float4*>(slot + G * ROW);
        float4* gw = reinterpret_cast<float4*>(p.weights + r0 * N);
        float4* gt = reinterpret_cast<float4*>(p.trans + r0 * N);
#pragma unroll
        for (int i = 0; i < G * N / 128; ++i) {
          __stcs(gw + lane + 32 * i, w4[lane + 32 * i]);
          __stcs(gt + lane + 32 * i, t4[lane + 32 * i]);
        }
      } else {
        warp_store(p.weights + r0 * N, wstage, cnt * N, lane);
        warp_store(p.trans + r0 * N, slot + G * ROW, cnt * N, lane);
      }
    }
    if (AUX) {
#pragma unroll
      for (int c = 0; c < 8; ++c) acc_a[c] = seg_sum<LPR>(acc_a[c]);
      if (p.aux && sl == 0 && active) {
        float4* a4 = reinterpret_cast<float4*>(p.aux + r * 8);
        a4[0] = make_float4(acc_a[0], acc_a[1], acc_a[2], acc_a[3]);
        a4[1] = make_float4(acc_a[4], acc_a[5], acc_a[6], acc_a[7]);
      }
    }
    acc_d = seg_sum<LPR>(acc_d);
#pragma unroll
    for (int c = 0; c < 3; ++c) acc_c[c] = seg_sum<LPR>(acc_c[c]);
#pragma unroll
    for (int c = 0; c < NS; ++c) acc_s[c] = seg_sum<LPR>(acc_s[c]);
    if (sl == 0 && active) {
      p.depth[r] = acc_d;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (p.rgb_raw) p.rgb_raw[r * 3 + c] = acc_c[c];
        p.rgb[r * 3 + c] = fminf(fmaxf(acc_c[c], 0.f), 1.f);                             // :134
      }
#pragma unroll
      for (int c = 0; c < NS; ++c) p.sem[r * NS + c] = acc_s[c] / (float)N;              // :156 (plain mean)
      if (AUX && NS > 0 && p.sem_argmax) p.sem_argmax[r] = argmax_first<NS>(acc_s, NS);
    }
    __syncwarp();
  }
}

template <int G, int NO, int NS, bool EXT>
__global__ void __launch_bounds__(64) composite_bwd_fast(const BwdP p) {
  extern __shared__ __align__(16) float sm[];
  constexpr int LPR = 32 / G, N = LPR * 8, ROW = N * NO + kSkew, EXTF = EXT ? ROW : 0, CS = NO - NS;
  constexpr int BUF = G * (ROW + EXTF + 3 * N);             // rows | external gradient rows | z | w | T
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  const int sub = lane / LPR, sl = lane % LPR;
  float* base = sm + (size_t)wib * (2 * BUF + 4);
  uint64_t* bars = reinterpret_cast<uint64_t*>(base + 2 * BUF);
  const int64_t warp0 = (int64_t)blockIdx.x * wpb + wib, nw = (int64_t)gridDim.x * wpb;
  const int64_t n_groups = (p.n_rays + G - 1) / G;
  auto prefetch = [&](int64_t g, int slot) {
    const int64_t r0 = g * G;
    const uint32_t cnt = (uint32_t)min((int64_t)G, p.n_rays - r0);
    constexpr uint32_t row_bytes = (uint32_t)(N * NO) * 4u;
    const uint32_t z_bytes = cnt * N * 4u;
    float* b = base + slot * BUF;
    mbar_expect_tx(&bars[slot], cnt * row_bytes * (EXT ? 2u : 1u) + 3u * z_bytes);
    for (uint32_t q = 0; q < cnt; ++q) {
      bulk_g2s(b + q * ROW, p.out + (r0 + q) * (N * NO), row_bytes, &bars[slot]);
      if (EXT) bulk_g2s(b + (G + q) * ROW, p.g_out_ext + (r0 + q) * (N * NO), row_bytes, &bars[slot]);
    }
    float* zb = b + G * (ROW + EXTF);
    bulk_g2s(zb, p.z + r0 * N, z_bytes, &bars[slot]);
    bulk_g2s(zb + G * N, p.weights + r0 * N, z_bytes, &bars[slot]);
    bulk_g2s(zb + 2 * G * N, p.trans + r0 * N, z_bytes, &bars[slot]);
  };
  if (lane == 0) {
    mbar_init(&bars[0], 1); mbar_init(&bars[1], 1);
    fence_mbar_init();
    if (warp0 < n_groups) prefetch(warp0, 0);
  }
  __syncwarp();
  float amax = 0.f;
  int t = 0;
  for (int64_t g = warp0; g < n_groups; g += nw, ++t) {
    float m = 0.f;                                          // max |gradient| of this step (dropped for padding rays)
    float* slot = base + (t & 1) * BUF;
    const int64_t r0 = g * G;
    const int cnt = (int)min((int64_t)G, p.n_rays - r0);
    const int64_t r = r0 + sub;
    const bool active = sub < cnt;
    float* o = slot + sub * ROW + sl * NO;                  // network row of sample sl, overwritten by its gradient
    const float* e_row = slot + G * ROW + sub * ROW + sl * NO;
    const float* zs = slot + G * (ROW + EXTF) + sub * N + sl;
    const float* ws = zs + G * N;
    const float* ts = ws + G * N;
    if (lane == 0 && g + nw < n_groups) prefetch(g + nw, (t + 1) & 1);
    float gh[3] = {0.f, 0.f, 0.f};
    if (p.g_rgb && active) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float raw = p.rgb_raw[r * 3 + c];
        gh[c] = (raw >= 0.f && raw <= 1.f) ? p.g_rgb[r * 3 + c] : 0.f;                   // clamp adjoint
      }
    }
    const float gd = (p.g_depth && active) ? p.g_depth[r] : 0.f;
    float gsem[NS > 0 ? NS : 1];
#pragma unroll
    for (int c = 0; c < NS; ++c) gsem[c] = (p.g_sem && active) ? p.g_sem[r * NS + c] / (float)N : 0.f;
    mbar_wait(&bars[t & 1], (uint32_t)(t >> 1) & 1u, 51);
    float Gd[8], Sv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float* ok = o + LPR * k * NO;
      const float sv = ok[4];
      float gg = gd * zs[LPR * k];
#pragma unroll
      for (int c = 0; c < 3; ++c) gg = fmaf(ok[c] * (sv + (1.f - sv) * ok[5 + c]), gh[c], gg);
      if (p.g_w && active) gg += p.g_w[r * N + sl + LPR * k];
      Gd[k] = gg;                                                                        // dL/dw_i
      Sv[k] = gg * ws[LPR * k] + ((p.g_t && active) ? p.g_t[r * N + sl + LPR * k] * ts[LPR * k] : 0.f);
    }
    float gsky[3] = {0.f, 0.f, 0.f};
    float carry = 0.f;                                       // sum of S over the samples of later rounds
#pragma unroll
    for (int k = 7; k >= 0; --k) {
      float ex, tot;
      seg_suffix_sum<LPR>(Sv[k], sl, ex, tot);
      const float R = carry + ex;
      float* ok = o + LPR * k * NO;
      const float w = ws[LPR * k], T = ts[LPR * k];
      float delta = zs[LPR * k + 1] - zs[LPR * k];
      if (k == 7 && sl == LPR - 1) delta = 1e10f;
      float s = ok[3];
      if (p.noise && active) s += p.noise[r * N + sl + LPR * k] * p.noise_std;
      const float e = __expf(-delta * fmaxf(s, 0.f));           // 1 - alpha
      const float dalpha = Gd[k] * T - __fdividef(R, e + 1e-10f);
      const float dsigma = (s > 0.f) ? dalpha * delta * e : 0.f;
      const float sv = ok[4];
      float go[8];
      float dsv = 0.f;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float sky = ok[5 + c];
        const float irr = sv + (1.f - sv) * sky;
        go[c] = w * irr * gh[c];                                // d albedo
        const float dirr = w * ok[c] * gh[c];
        dsv = fmaf(dirr, 1.f - sky, dsv);
        go[5 + c] = dirr * (1.f - sv);                          // d sky
      }
      go[3] = dsigma;
      go[4] = dsv;
      const float* ek = e_row + LPR * k * NO;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float v = go[c] + (EXT ? ek[c] : 0.f);
        ok[c] = v;
        m = fmaxf(m, fabsf(v));
        if (c >= 5) gsky[c - 5] += v;
      }
#pragma unroll
      for (int c = 8; c < NO; ++c) {
        float v = EXT ? ek[c] : 0.f;
        if (c >= CS) v += gsem[c - CS < 0 ? 0 : c - CS];
        ok[c] = v;
        m = fmaxf(m, fabsf(v));
      }
      carry += tot;
    }
    if (active) amax = fmaxf(amax, m);
    fence_proxy_async_smem();
    __syncwarp();
    if (cnt == G) {
      // the G gradient rows are ROW floats apart in shared memory, contiguous in global memory
#pragma unroll
      for (int q = 0; q < G; ++q) {
        const float4* s4 = reinterpret_cast<const float4*>(slot + q * ROW);
        float4* d4 = reinterpret_cast<float4*>(p.g_out + (r0 + q) * (N * NO));
#pragma unroll
        for (int i = 0; i < (N * NO / 4 + 31) / 32; ++i)
          if (lane + 32 * i < N * NO / 4) __stcs(d4 + lane + 32 * i, s4[lane + 32 * i]);
      }
    } else {
      for (int q = 0; q < cnt; ++q) warp_store(p.g_out + (r0 + q) * (N * NO), slot + q * ROW, N * NO, lane);
    }
    if (p.g_sky_ray) {
#pragma unroll
      for (int c = 0; c < 3; ++c) gsky[c] = seg_sum<LPR>(gsky[c]);
      if (sl < 3 && active) p.g_sky_ray[r * 3 + sl] = sl == 0 ? gsky[0] : sl == 1 ? gsky[1] : gsky[2];
    }
    __syncwarp();
  }
  if (p.absmax_bits) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, s));
    if (lane == 0 && amax > 0.f) atomicMax(p.absmax_bits, __float_as_uint(amax));   // positive floats order as uints
  }
}

// grid = the blocks that are resident at once (shared memory bound), each warp strides over ray groups
template <class K, class P>
int launch(K kern, const P& p, int64_t n_groups, int wpb, size_t smem, cudaStream_t stream, size_t*) {
  if (smem > 48 * 1024)
    if (cudaError_t e = sm100::set_max_dynamic_smem(reinterpret_cast<const void*>(kern), smem); e != cudaSuccess) return -(int)e;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t need = (n_groups + wpb - 1) / wpb;
  const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(64 / wpb, (224 * 1024) / (smem + 1024)));
  const int64_t cap = (int64_t)sms * per_sm;
  kern<<<(unsigned)(need < cap ? need : cap), wpb * 32, smem, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}

// rays per warp step: 4 (n <= 64), 2 (n <= 128) with bulk copies; 1 without
int pick_group(int n, int n_out, bool aligned) {
  if (!aligned || (n * n_out) % 4 || n % 4) return 1;
  if (n <= 8 * kMaxPerLane) return 4;
  if (n <= 16 * kMaxPerLane) return 2;
  return 1;
}

}  // namespace

extern "C" int spnerf_composite_fwd(const SpnerfCompositeFwd* a, void* stream) {
  if (!a || !a->out || !a->z || !a->rgb || !a->depth) return SPNERF_ERR_BAD_ARG;
  if ((a->weights == nullptr) != (a->transparency == nullptr)) return SPNERF_ERR_BAD_ARG;
  if (a->n_samples < 1 || a->n_samples > 32 * kMaxPerLane || a->n_out < 8 || a->n_sem < 0 || a->n_sem > 8)
    return SPNERF_ERR_UNSUPPORTED;
  if (a->n_sem > 0 && (!a->sem_logits || a->col_sem + a->n_sem > a->n_out)) return SPNERF_ERR_BAD_ARG;
  if (a->sem_argmax && a->n_sem == 0) return SPNERF_ERR_BAD_ARG;
  if (a->ray_aux && (a->col_beta >= a->n_out || (a->col_beta >= 0 && a->col_beta != 8))) return SPNERF_ERR_BAD_ARG;
  if (a->n_rays <= 0) return a->n_rays == 0 ? 0 : SPNERF_ERR_BAD_ARG;
  FwdP p;
  p.out = a->out; p.z = a->z; p.noise = (a->noise && a->noise_std != 0.f) ? a->noise : nullptr;
  p.noise_std = a->noise_std; p.n_rays = a->n_rays; p.n = a->n_samples; p.n_out = a->n_out;
  p.col_sem = a->col_sem; p.n_sem = a->n_sem;
  p.weights = a->weights; p.trans = a->transparency; p.rgb = a->rgb; p.rgb_raw = a->rgb_raw; p.depth = a->depth;
  p.sem = a->sem_logits;
  p.aux = a->ray_aux; p.sem_argmax = a->sem_argmax; p.col_beta = a->ray_aux ? a->col_beta : -1;
  const int row_f = (p.n * p.n_out + 3) & ~3;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const int G = pick_group(p.n, p.n_out, al16(p.out) && al16(p.z));
  const int wpb = 4;
  const int rf = row_f + (G > 1 ? kSkew : 0);
  const size_t smem = (size_t)wpb * ((G > 1 ? 2 : 1) * G * (rf + p.n) + G * p.n + 4) * sizeof(float);
  static size_t cfg[3] = {0, 0, 0};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool beta_ok = !p.aux || ((p.col_beta == 8) == (p.n_out - p.n_sem == 9));
  if (G > 1 && p.n == 8 * (32 / G) && al16(p.weights) && al16(p.trans) && (!p.aux || al16(p.aux)) && beta_ok &&
      (p.n_sem == 0 || p.col_sem == p.n_out - p.n_sem)) {
    static size_t fcfg[16] = {0};
    const int64_t ng = (p.n_rays + G - 1) / G;
    const bool aux = p.aux || p.sem_argmax;
#define SPNERF_FWD_FAST(g_, no_, ns_, slot_)                                                        \
    if (G == g_ && p.n_out == no_ && p.n_sem == ns_) {                                              \
      if (aux) return launch(composite_fwd_fast<g_, no_, ns_, true>, p, ng, wpb, smem, st, &fcfg[2 * slot_]);      \
      return launch(composite_fwd_fast<g_, no_, ns_, false>, p, ng, wpb, smem, st, &fcfg[2 * slot_ + 1]);          \
    }
    SPNERF_FWD_FAST(4, 11, 3, 0) SPNERF_FWD_FAST(4, 8, 0, 1) SPNERF_FWD_FAST(4, 9, 0, 2) SPNERF_FWD_FAST(4, 12, 3, 3)
    SPNERF_FWD_FAST(2, 11, 3, 4) SPNERF_FWD_FAST(2, 8, 0, 5) SPNERF_FWD_FAST(2, 9, 0, 6) SPNERF_FWD_FAST(2, 12, 3, 7)
#undef SPNERF_FWD_FAST
  }
  if (G == 4) return launch(composite_fwd_kernel<4>, p, (p.n_rays + 3) / 4, wpb, smem, st, &cfg[0]);
  if (G == 2) return launch(composite_fwd_kernel<2>, p, (p.n_rays + 1) / 2, wpb, smem, st, &cfg[1]);
  return launch(composite_fwd_kernel<1>, p, p.n_rays, wpb, smem, st, &cfg[2]);
}

extern "C" int spnerf_composite_bwd(const SpnerfCompositeBwd* a, void* stream) {
  if (!a || !a->out || !a->z || !a->weights || !a->transparency || !a->g_out) return SPNERF_ERR_BAD_ARG;
  if (a->g_rgb && !a->rgb_raw) return SPNERF_ERR_BAD_ARG;
  if (a->n_samples < 1 || a->n_samples > 32 * kMaxPerLane || a->n_out < 8 || a->n_sem < 0 || a->n_sem > 8)
    return SPNERF_ERR_UNSUPPORTED;
  if (a->n_rays <= 0) return a->n_rays == 0 ? 0 : SPNERF_ERR_BAD_ARG;
  BwdP p;
  p.out = a->out; p.z = a->z; p.noise = (a->noise && a->noise_std != 0.f) ? a->noise : nullptr;
  p.noise_std = a->noise_std; p.weights = a->weights; p.trans = a->transparency; p.rgb_raw = a->rgb_raw;
  p.g_rgb = a->g_rgb; p.g_depth = a->g_depth; p.g_sem = a->g_sem_logits; p.g_w = a->g_weights;
  p.g_t = a->g_transparency; p.g_out_ext = a->g_out_ext;
  p.n_rays = a->n_rays; p.n = a->n_samples; p.n_out = a->n_out; p.col_sem = a->col_sem; p.n_sem = a->n_sem;
  p.g_out = a->g_out; p.g_sky_ray = a->g_sky_ray; p.absmax_bits = reinterpret_cast<unsigned int*>(a->g_absmax);
  const int row_f = (p.n * p.n_out + 3) & ~3;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  const int G = pick_group(p.n, p.n_out, al16(p.out) && al16(p.z) && al16(p.weights) && al16(p.trans) &&
                                             (!p.g_out_ext || al16(p.g_out_ext)));
  const int slot_f = (row_f + (G > 1 ? kSkew : 0)) * (p.g_out_ext ? 2 : 1) + 3 * p.n;
  const int wpb = G > 1 ? 2 : 4;      // more resident blocks of 2 warps (shared memory bound)
  const size_t smem = (size_t)wpb * ((G > 1 ? 2 : 1) * G * slot_f + 4) * sizeof(float);
  static size_t cfg[3] = {0, 0, 0};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (G > 1 && p.n == 8 * (32 / G) && al16(p.g_out) && (p.n_sem == 0 || p.col_sem == p.n_out - p.n_sem)) {
    static size_t fcfg[16] = {0};
    const int64_t ng = (p.n_rays + G - 1) / G;
#define SPNERF_BWD_FAST(g_, no_, ns_, slot_)                                                        \
    if (G == g_ && p.n_out == no_ && p.n_sem == ns_) {                                              \
      if (p.g_out_ext) return launch(composite_bwd_fast<g_, no_, ns_, true>, p, ng, wpb, smem, st, &fcfg[2 * slot_]);      \
      return launch(composite_bwd_fast<g_, no_, ns_, false>, p, ng, wpb, smem, st, &fcfg[2 * slot_ + 1]);                  \
    }
    SPNERF_BWD_FAST(4, 11, 3, 0) SPNERF_BWD_FAST(4, 8, 0, 1) SPNERF_BWD_FAST(4, 9, 0, 2) SPNERF_BWD_FAST(4, 12, 3, 3)
    SPNERF_BWD_FAST(2, 11, 3, 4) SPNERF_BWD_FAST(2, 8, 0, 5) SPNERF_BWD_FAST(2, 9, 0, 6) SPNERF_BWD_FAST(2, 12, 3, 7)
#undef SPNERF_BWD_FAST
  }
  if (G == 4) return launch(composite_bwd_kernel<4>, p, (p.n_rays + 3) / 4, wpb, smem, st, &cfg[0]);
  if (G == 2) return launch(composite_bwd_kernel<2>, p, (p.n_rays + 1) / 2, wpb, smem, st, &cfg[1]);
  return launch(composite_bwd_kernel<1>, p, p.n_rays, wpb, smem, st, &cfg[2]);
}
