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
#include "../../include/spnerf_b200.h"

namespace {

constexpr int kWarpsPerBlock = 4;
constexpr int kMaxPerLane = 8;      // n_samples <= 256

__device__ __forceinline__ float warp_excl_scan_mul(float v, int lane) {
  float inc = v;
#pragma unroll
  for (int s = 1; s < 32; s <<= 1) {
    const float o = __shfl_up_sync(0xffffffffu, inc, s);
    if (lane >= s) inc *= o;
  }
  const float ex = __shfl_up_sync(0xffffffffu, inc, 1);
  return lane == 0 ? 1.f : ex;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}
// exclusive suffix sum: result(lane) = sum of v over lanes > lane
__device__ __forceinline__ float warp_excl_suffix_sum(float v, int lane) {
  float inc = v;
#pragma unroll
  for (int s = 1; s < 32; s <<= 1) {
    const float o = __shfl_down_sync(0xffffffffu, inc, s);
    if (lane + s < 32) inc += o;
  }
  const float ex = __shfl_down_sync(0xffffffffu, inc, 1);
  return lane == 31 ? 0.f : ex;
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
};

__global__ void __launch_bounds__(kWarpsPerBlock * 32) composite_fwd_kernel(const FwdP p) {
  extern __shared__ __align__(16) float sm[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int row_f = (p.n * p.n_out + 3) & ~3;
  float* rows = sm + (size_t)wib * (row_f + 2 * p.n);     // [n][n_out]
  float* zs = rows + row_f;                               // [n]
  float* ws = zs + p.n;                                   // [n]  (weights, then reused for T)
  const int per = (p.n + 31) / 32;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + wib, nw = (int64_t)gridDim.x * kWarpsPerBlock;
  for (int64_t r = warp0; r < p.n_rays; r += nw) {
    warp_load(rows, p.out + r * p.n * p.n_out, p.n * p.n_out, lane);
    warp_load(zs, p.z + r * p.n, p.n, lane);
    __syncwarp();
    // lane owns samples [i0, i1)
    const int i0 = min(lane * per, p.n), i1 = min(i0 + per, p.n);
    float alpha[kMaxPerLane], tloc[kMaxPerLane];
    float prod = 1.f;
#pragma unroll
    for (int k = 0; k < kMaxPerLane; ++k) {
      const int i = i0 + k;
      if (k < per && i < i1) {
        const float delta = (i + 1 < p.n) ? zs[i + 1] - zs[i] : 1e10f;                 // spnerf.py:116-118
        float s = rows[i * p.n_out + 3];
        if (p.noise) s += p.noise[r * p.n + i] * p.noise_std;                          // :121-122
        alpha[k] = 1.f - expf(-delta * fmaxf(s, 0.f));                                 // :123
        tloc[k] = prod;
        prod *= (1.f - alpha[k] + 1e-10f);                                             // :126
      }
    }
    const float before = warp_excl_scan_mul(prod, lane);                               // :127 (exclusive cumprod)
    float acc_d = 0.f, acc_c[3] = {0.f, 0.f, 0.f}, acc_s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int k = 0; k < kMaxPerLane; ++k) {
      const int i = i0 + k;
      if (k < per && i < i1) {
        const float T = before * tloc[k];
        const float w = alpha[k] * T;                                                  // :128
        const float* o = rows + i * p.n_out;
        const float sv = o[4];
        acc_d = fmaf(w, zs[i], acc_d);                                                 // :131
#pragma unroll
        for (int c = 0; c < 3; ++c) acc_c[c] = fmaf(w * o[c], sv + (1.f - sv) * o[5 + c], acc_c[c]);   // :132-133
        for (int c = 0; c < p.n_sem; ++c) acc_s[c] += o[p.col_sem + c];
        ws[i] = w;
        zs[i] = T;   // z no longer needed by this lane's block; neighbours read zs[i+1] only before this point
      }
    }
    // (the read of zs[i1] by this lane happened in the first loop, before any lane overwrote it:
    //  the scan's shuffles order the two loops across the warp)
    __syncwarp();
    warp_store(p.weights + r * p.n, ws, p.n, lane);
    warp_store(p.trans + r * p.n, zs, p.n, lane);
    acc_d = warp_sum(acc_d);
#pragma unroll
    for (int c = 0; c < 3; ++c) acc_c[c] = warp_sum(acc_c[c]);
    for (int c = 0; c < p.n_sem; ++c) acc_s[c] = warp_sum(acc_s[c]);
    if (lane == 0) {
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

__global__ void __launch_bounds__(kWarpsPerBlock * 32) composite_bwd_kernel(const BwdP p) {
  extern __shared__ __align__(16) float sm[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int row_f = (p.n * p.n_out + 3) & ~3;
  float* rows = sm + (size_t)wib * (2 * row_f + 3 * p.n);   // network rows, overwritten by their gradients
  float* ext = rows + row_f;                                // external gradient rows (optional)
  float* zs = ext + row_f;
  float* ws = zs + p.n;
  float* ts = ws + p.n;
  const int per = (p.n + 31) / 32;
  const int64_t warp0 = (int64_t)blockIdx.x * kWarpsPerBlock + wib, nw = (int64_t)gridDim.x * kWarpsPerBlock;
  float amax = 0.f;
  for (int64_t r = warp0; r < p.n_rays; r += nw) {
    warp_load(rows, p.out + r * p.n * p.n_out, p.n * p.n_out, lane);
    if (p.g_out_ext) warp_load(ext, p.g_out_ext + r * p.n * p.n_out, p.n * p.n_out, lane);
    warp_load(zs, p.z + r * p.n, p.n, lane);
    warp_load(ws, p.weights + r * p.n, p.n, lane);
    warp_load(ts, p.trans + r * p.n, p.n, lane);
    __syncwarp();
    float gh[3] = {0.f, 0.f, 0.f};
    if (p.g_rgb) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float raw = p.rgb_raw[r * 3 + c];
        gh[c] = (raw >= 0.f && raw <= 1.f) ? p.g_rgb[r * 3 + c] : 0.f;                 // clamp adjoint
      }
    }
    const float gd = p.g_depth ? p.g_depth[r] : 0.f;
    const int i0 = min(lane * per, p.n), i1 = min(i0 + per, p.n);
    float G[kMaxPerLane], S_loc = 0.f;
#pragma unroll
    for (int k = 0; k < kMaxPerLane; ++k) {
      const int i = i0 + k;
      G[k] = 0.f;
      if (k < per && i < i1) {
        const float* o = rows + i * p.n_out;
        const float sv = o[4];
        float g = gd * zs[i];
#pragma unroll
        for (int c = 0; c < 3; ++c) g = fmaf(o[c] * (sv + (1.f - sv) * o[5 + c]), gh[c], g);
        if (p.g_w) g += p.g_w[r * p.n + i];
        G[k] = g;                                                                       // dL/dw_i
        S_loc += g * ws[i] + (p.g_t ? p.g_t[r * p.n + i] * ts[i] : 0.f);
      }
    }
    float R = warp_excl_suffix_sum(S_loc, lane);     // sum of S over later lanes' blocks
    float gsky[3] = {0.f, 0.f, 0.f};
    // walk this lane's block backwards so R is the exclusive suffix sum at each sample
#pragma unroll
    for (int k = kMaxPerLane - 1; k >= 0; --k) {
      const int i = i0 + k;
      if (k < per && i < i1) {
        float* o = rows + i * p.n_out;
        const float w = ws[i], T = ts[i];
        const float delta = (i + 1 < p.n) ? zs[i + 1] - zs[i] : 1e10f;
        float s = o[3];
        if (p.noise) s += p.noise[r * p.n + i] * p.noise_std;
        const float e = expf(-delta * fmaxf(s, 0.f));          // 1 - alpha
        const float dalpha = G[k] * T - R / (e + 1e-10f);
        const float dsigma = (s > 0.f) ? dalpha * delta * e : 0.f;
        R += G[k] * w + (p.g_t ? p.g_t[r * p.n + i] * T : 0.f);
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
          amax = fmaxf(amax, fabsf(v));
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) gsky[c] += o[5 + c];
        for (int c = 8; c < p.n_out; ++c) {
          float v = p.g_out_ext ? e_row[c] : 0.f;
          if (c >= p.col_sem && c < p.col_sem + p.n_sem && p.g_sem) v += p.g_sem[r * p.n_sem + (c - p.col_sem)] / (float)p.n;
          o[c] = v;
          amax = fmaxf(amax, fabsf(v));
        }
      }
    }
    __syncwarp();
    warp_store(p.g_out + r * p.n * p.n_out, rows, p.n * p.n_out, lane);
    if (p.g_sky_ray) {
#pragma unroll
      for (int c = 0; c < 3; ++c) gsky[c] = warp_sum(gsky[c]);
      if (lane < 3) p.g_sky_ray[r * 3 + lane] = lane == 0 ? gsky[0] : lane == 1 ? gsky[1] : gsky[2];
    }
    __syncwarp();
  }
  if (p.absmax_bits) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, s));
    if (lane == 0 && amax > 0.f) atomicMax(p.absmax_bits, __float_as_uint(amax));   // positive floats order as uints
  }
}

int launch_dims(int64_t n_rays, int* blocks) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t need = (n_rays + kWarpsPerBlock - 1) / kWarpsPerBlock;
  const int64_t cap = (int64_t)sms * 16;
  *blocks = (int)(need < cap ? need : cap);
  return 0;
}

}  // namespace

extern "C" int spnerf_composite_fwd(const SpnerfCompositeFwd* a, void* stream) {
  if (!a || !a->out || !a->z || !a->weights || !a->transparency || !a->rgb || !a->depth) return SPNERF_ERR_BAD_ARG;
  if (a->n_samples < 1 || a->n_samples > 32 * kMaxPerLane || a->n_out < 8 || a->n_sem < 0 || a->n_sem > 8)
    return SPNERF_ERR_UNSUPPORTED;
  if (a->n_sem > 0 && (!a->sem_logits || a->col_sem + a->n_sem > a->n_out)) return SPNERF_ERR_BAD_ARG;
  if (a->n_rays <= 0) return a->n_rays == 0 ? 0 : SPNERF_ERR_BAD_ARG;
  FwdP p;
  p.out = a->out; p.z = a->z; p.noise = (a->noise && a->noise_std != 0.f) ? a->noise : nullptr;
  p.noise_std = a->noise_std; p.n_rays = a->n_rays; p.n = a->n_samples; p.n_out = a->n_out;
  p.col_sem = a->col_sem; p.n_sem = a->n_sem;
  p.weights = a->weights; p.trans = a->transparency; p.rgb = a->rgb; p.rgb_raw = a->rgb_raw; p.depth = a->depth;
  p.sem = a->sem_logits;
  const int row_f = (p.n * p.n_out + 3) & ~3;
  const size_t smem = (size_t)kWarpsPerBlock * (row_f + 2 * p.n) * sizeof(float);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(composite_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return -(int)e;
    configured = smem;
  }
  int blocks;
  launch_dims(p.n_rays, &blocks);
  composite_fwd_kernel<<<blocks, kWarpsPerBlock * 32, smem, static_cast<cudaStream_t>(stream)>>>(p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
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
  const size_t smem = (size_t)kWarpsPerBlock * (2 * row_f + 3 * p.n) * sizeof(float);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(composite_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return -(int)e;
    configured = smem;
  }
  int blocks;
  launch_dims(p.n_rays, &blocks);
  composite_bwd_kernel<<<blocks, kWarpsPerBlock * 32, smem, static_cast<cudaStream_t>(stream)>>>(p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}
