// Ray samplers: stratified near/far depths and the depth-guided (--guidedsample) resampler.
//
// Replaces modules/rendering.py:128-144 (coarse z), :14-55 sample_pdf, :58-73 sample_3sigma,
// :76-89 compute_samples_around_depth, :92-116 GenerateGuidedSamples and the sort/concat/sort of
// :165-167.  Random numbers are inputs (the reference draws them with torch.rand*): parity tests
// inject the oracle's draws, production passes device-generated uniforms.
//
// Bit-exactness contract: identical sample depths, CDF and searchsorted
// indices to the torch-CPU reference.  That requires reproducing torch's arithmetic:
//  * every elementwise op rounds separately (no FMA contraction): __fmul_rn/__fadd_rn/...,
//  * torch.sum over a contiguous fp32 row = 8-lane vector accumulation (4 accumulators, then
//    lanes in order), see row_sum_torch();
//  * torch.cumsum(fp32) = running sum kept in fp64, rounded to fp32 per element;
//  * torch.linspace / Gaussian-weight tables are supplied by the host (computed by torch).
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/spnerf_b200.h"

namespace {

// ---- device-side uniforms: Philox4x32-10 keyed by the seed, counter = (element index, step) ------------
// `state` = {seed, step, arrivals (low 32 bits)}: every block reads the step when it starts and the block that
// arrives last advances it, so consecutive launches (and replays of a captured graph) draw fresh numbers with no
// host involvement.  Values are k * 2^-24, k < 2^24, like torch.rand's float32 draws.
__device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) { return __umulhi(a, b); }
__device__ __forceinline__ float philox_uniform(uint64_t seed, uint64_t step, uint64_t idx) {
  uint32_t c0 = (uint32_t)idx, c1 = (uint32_t)(idx >> 32), c2 = (uint32_t)step, c3 = (uint32_t)(step >> 32);
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t h0 = mulhi32(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
    const uint32_t h1 = mulhi32(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
    c0 = h1 ^ c1 ^ k0; c1 = l1; c2 = h0 ^ c3 ^ k1; c3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return (float)(c0 >> 8) * 5.9604644775390625e-8f;
}

// ---- coarse: z = lower + (upper-lower)*u over stratified bins of near*(1-t)+far*t -----------------
__global__ void coarse_kernel(const float* __restrict__ rays, const float* __restrict__ t_tab,
                              const float* __restrict__ u, unsigned long long* rng_state, int64_t n_rays, int n,
                              float* __restrict__ z) {
  const int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  unsigned long long seed = 0, step = 0;
  if (rng_state) {
    seed = rng_state[0];
    step = *reinterpret_cast<volatile unsigned long long*>(rng_state + 1);
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int* arrivals = reinterpret_cast<unsigned int*>(rng_state + 2);
      if (atomicAdd(arrivals, 1u) == gridDim.x - 1) { *arrivals = 0u; rng_state[1] = step + 1; }
    }
  }
  // grid-stride: a bounded grid keeps the arrival counter above to a few hundred atomics
  for (int64_t k = idx; k < n_rays * n; k += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = k / n;
    const int i = (int)(k - r * n);
    const float near = rays[r * 11 + 6], far = rays[r * 11 + 7];
    auto zlin = [&](int q) {                                        // rendering.py:133
      const float t = t_tab[q];
      return __fadd_rn(__fmul_rn(near, __fsub_rn(1.f, t)), __fmul_rn(far, t));
    };
    const float zi = zlin(i);
    const float lower = i == 0 ? zi : __fmul_rn(0.5f, __fadd_rn(zlin(i - 1), zi));        // :138-141
    const float upper = i == n - 1 ? zi : __fmul_rn(0.5f, __fadd_rn(zi, zlin(i + 1)));
    const float ui = rng_state ? philox_uniform(seed, step, (uint64_t)k) : u[k];
    z[k] = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), ui));                     // :143-144 (perturb = 1)
  }
}

// ---- torch.sum(row) for a contiguous fp32 row (ATen vectorized inner reduction, 8-float vectors) ----
template <class Get>
__device__ float row_sum_torch(int n, Get x) {
  const int m = n / 8;
  float P[4][8];
#pragma unroll
  for (int k = 0; k < 4; ++k)
#pragma unroll
    for (int l = 0; l < 8; ++l) P[k][l] = 0.f;
  const int full = m / 4;
  for (int i = 0; i < full; ++i)
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
      for (int l = 0; l < 8; ++l) P[k][l] = __fadd_rn(P[k][l], x((4 * i + k) * 8 + l));
  for (int v = 4 * full; v < m; ++v)
#pragma unroll
    for (int l = 0; l < 8; ++l) P[0][l] = __fadd_rn(P[0][l], x(v * 8 + l));
#pragma unroll
  for (int k = 1; k < 4; ++k)
#pragma unroll
    for (int l = 0; l < 8; ++l) P[0][l] = __fadd_rn(P[0][l], P[k][l]);
  float acc = 0.f;
  for (int j = 8 * m; j < n; ++j) acc = __fadd_rn(acc, x(j));
#pragma unroll
  for (int l = 0; l < 8; ++l) acc = __fadd_rn(acc, P[0][l]);
  return acc;
}

struct GuidedP {
  const float* rays; const float* z; const float* weights; const float* depth;
  const int64_t* valid; const float* target_depth; int64_t target_stride; const float* target_std;
  const float* u_pred; const float* u_gt;
  const float* t_tab; const float* gauss_tab;
  int64_t n_rays; int n;
  float* z_unsort; float* z_sorted; int32_t* inds_out;
};

// one thread per ray; per-thread arrays live in shared memory, interleaved by thread
__global__ void guided_kernel(const GuidedP p) {
  extern __shared__ float sm[];
  const int tid = threadIdx.x, bs = blockDim.x, n = p.n;
  float* edges = sm + tid;                   // [n]   element i at edges[i*bs]
  float* cdf = sm + (size_t)n * bs + tid;    // [n]
  float* smp = sm + (size_t)2 * n * bs + tid;   // [n]
  const int64_t r = blockIdx.x * (int64_t)bs + tid;
  if (r >= p.n_rays) return;
  const float near0 = p.rays[6], far0 = p.rays[7];                 // first ray of the batch (rendering.py:95,113)
  const float* zr = p.z + r * n;
  float lo, hi;
  const bool use_gt = p.valid && p.valid[r] > 0;                   // rendering.py:98-114
  const float* u = use_gt ? p.u_gt + r * n : p.u_pred + r * n;
  if (use_gt) {
    const float td = p.target_depth[r * p.target_stride], ts = p.target_std[r];
    lo = __fsub_rn(td, __fmul_rn(3.f, ts));                        // :107-108
    hi = __fadd_rn(td, __fmul_rn(3.f, ts));
  } else {
    const float d = p.depth[r];
    const float* wr = p.weights + r * n;
    const float var = row_sum_torch(n, [&](int i) {                 // :81
      const float df = __fsub_rn(zr[i], d);
      return __fmul_rn(__fmul_rn(df, df), wr[i]);
    });
    const float sd = __fsqrt_rn(var);
    lo = __fsub_rn(d, __fmul_rn(3.f, sd));                          // :83-84
    hi = __fadd_rn(d, __fmul_rn(3.f, sd));
  }
  const float step = __fdiv_rn(__fsub_rn(hi, lo), (float)(n - 1));  // :62
  for (int j = 0; j < n; ++j) {                                      // :64
    const float t = p.t_tab[j];
    float e = __fadd_rn(__fmul_rn(lo, __fsub_rn(1.f, t)), __fmul_rn(hi, t));
    e = fminf(fmaxf(e, near0), far0);
    edges[j * bs] = e;
  }
  // bin weights + eps (:66-71, :27); stored temporarily in cdf[1..n-1]
  for (int j = 0; j < n - 1; ++j) {
    const float factor = __fdiv_rn(__fsub_rn(edges[(j + 1) * bs], edges[j * bs]), step);
    cdf[(j + 1) * bs] = __fadd_rn(__fmul_rn(factor, p.gauss_tab[j]), 1e-5f);
  }
  const float wsum = row_sum_torch(n - 1, [&](int j) { return cdf[(j + 1) * bs]; });   // :28
  double run = 0.0;
  cdf[0] = 0.f;                                                     // :30
  for (int j = 0; j < n - 1; ++j) {                                 // :28-29
    const float pdf = __fdiv_rn(cdf[(j + 1) * bs], wsum);
    run += (double)pdf;
    cdf[(j + 1) * bs] = (float)run;
  }
  for (int k = 0; k < n; ++k) {                                     // :38-54
    const float uk = u[k];
    int lo_i = 0, hi_i = n;                                         // searchsorted(right=True): #entries <= u
    while (lo_i < hi_i) {
      const int mid = (lo_i + hi_i) >> 1;
      if (cdf[mid * bs] <= uk) lo_i = mid + 1; else hi_i = mid;
    }
    const int inds = lo_i;
    if (p.inds_out) p.inds_out[r * n + k] = inds;
    const int below = inds - 1 > 0 ? inds - 1 : 0;
    const int above = inds < n - 1 ? inds : n - 1;
    const float cb = cdf[below * bs], bb = edges[below * bs];
    float denom = __fsub_rn(cdf[above * bs], cb);
    if (denom < 1e-5f) denom = 1.f;
    const float s = __fadd_rn(bb, __fmul_rn(__fdiv_rn(__fsub_rn(uk, cb), denom), __fsub_rn(edges[above * bs], bb)));
    // insertion into the sorted prefix (:165)
    int q = k;
    while (q > 0 && smp[(q - 1) * bs] > s) { smp[q * bs] = smp[(q - 1) * bs]; --q; }
    smp[q * bs] = s;
  }
  float* un = p.z_unsort + r * 2 * n;                               // :166
  float* so = p.z_sorted + r * 2 * n;                               // :167 (merge of two sorted runs)
  int a = 0, b = 0;
  for (int k = 0; k < n; ++k) { un[k] = zr[k]; un[n + k] = smp[k * bs]; }
  for (int k = 0; k < 2 * n; ++k) {
    const bool take_a = b >= n || (a < n && zr[a] <= smp[b * bs]);
    so[k] = take_a ? zr[a++] : smp[(b++) * bs];
  }
}

}  // namespace

extern "C" int spnerf_sample_coarse(const float* rays, const float* t_table, const float* uniforms, int64_t n_rays,
                                    int32_t n_samples, float* z, void* stream) {
  if (!rays || !t_table || !uniforms || !z || n_samples < 2) return SPNERF_ERR_BAD_ARG;
  if (n_rays <= 0) return n_rays == 0 ? 0 : SPNERF_ERR_BAD_ARG;
  const int64_t total = n_rays * n_samples;
  coarse_kernel<<<(unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      rays, t_table, uniforms, nullptr, n_rays, n_samples, z);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}

extern "C" int spnerf_sample_coarse_rng(const float* rays, const float* t_table, uint64_t* rng_state, int64_t n_rays,
                                        int32_t n_samples, float* z, void* stream) {
  if (!rays || !t_table || !rng_state || !z || n_samples < 2) return SPNERF_ERR_BAD_ARG;
  if (n_rays <= 0) return n_rays == 0 ? 0 : SPNERF_ERR_BAD_ARG;
  const int64_t total = n_rays * n_samples;
  const int64_t blocks = (total + 255) / 256;
  coarse_kernel<<<(unsigned)(blocks < 148 * 8 ? blocks : 148 * 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      rays, t_table, nullptr, reinterpret_cast<unsigned long long*>(rng_state), n_rays, n_samples, z);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}

extern "C" int spnerf_sample_guided(const SpnerfGuided* a, void* stream) {
  if (!a || !a->rays || !a->z || !a->weights || !a->depth || !a->u_pred || !a->t_table || !a->gauss_table ||
      !a->z_unsort || !a->z_sorted)
    return SPNERF_ERR_BAD_ARG;
  if (a->valid_depth && (!a->target_depth || !a->target_std || !a->u_gt)) return SPNERF_ERR_BAD_ARG;
  if (a->n_samples < 3 || a->n_samples > 256) return SPNERF_ERR_UNSUPPORTED;
  if (a->n_rays <= 0) return a->n_rays == 0 ? 0 : SPNERF_ERR_BAD_ARG;
  GuidedP p;
  p.rays = a->rays; p.z = a->z; p.weights = a->weights; p.depth = a->depth;
  p.valid = a->valid_depth; p.target_depth = a->target_depth; p.target_stride = a->target_depth_stride;
  p.target_std = a->target_std; p.u_pred = a->u_pred; p.u_gt = a->u_gt; p.t_tab = a->t_table;
  p.gauss_tab = a->gauss_table; p.n_rays = a->n_rays; p.n = a->n_samples;
  p.z_unsort = a->z_unsort; p.z_sorted = a->z_sorted; p.inds_out = a->searchsorted_out;
  const int bs = 32;
  const size_t smem = (size_t)3 * p.n * bs * sizeof(float);
  static size_t configured[64] = {};      // per device
  int dev = 0;
  cudaGetDevice(&dev);
  size_t* conf = (dev >= 0 && dev < 64) ? &configured[dev] : nullptr;
  if (smem > 48 * 1024 && (!conf || smem > *conf)) {
    cudaError_t e = cudaFuncSetAttribute(guided_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return -(int)e;
    if (conf) *conf = smem;
  }
  guided_kernel<<<(unsigned)((p.n_rays + bs - 1) / bs), bs, smem, static_cast<cudaStream_t>(stream)>>>(p);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}
