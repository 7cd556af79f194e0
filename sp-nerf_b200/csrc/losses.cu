// Fused loss reductions + gradients: colour MSE, depth supervision, semantic cross-entropy, solar-correction
// terms, uncertainty-aware colour loss.
//
// Replaces modules/metrics.py:10-24 (uncertainty_aware_loss, solar_correction), :27-45 (SNerfLoss colour term),
// :68-159 (DepthLoss, MSE variants) and :162-183 (SemanticLoss): one warp per ray, block partial sums, the last block to finish adds
// the partials in block order (deterministic scalars).
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/spnerf_b200.h"

namespace {

constexpr int kThreads = 256;
constexpr int kMaxBlocks = 1184;      // 148 SMs x 8
struct Workspace {
  unsigned int ticket;
  unsigned int n_labelled;
  unsigned int _pad[2];
  float partial[kMaxBlocks][8];      // colour, depth, semantic, applied depth rays, labelled rays
};

__global__ void count_labels_kernel(const int64_t* __restrict__ labels, int64_t n, int n_sem, Workspace* ws) {
  unsigned int c = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    c += labels[i] >= 0 && labels[i] < n_sem;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) c += __shfl_xor_sync(0xffffffffu, c, s);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&ws->n_labelled, c);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}

// COUNTED: the labelled rays were counted by count_labels_kernel (large batches).  Otherwise the kernel counts them
// itself: the cross-entropy gradient is written without its 1 / n_labelled and the last block to finish divides it
// (one launch less per training step; the same operations in the same order, so the same bits).
constexpr int64_t kSelfCountMax = 1 << 18;      // gradient elements one block rescales in a few microseconds
template <bool COUNTED>
__global__ void __launch_bounds__(kThreads) losses_kernel(const SpnerfLosses a, Workspace* ws) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + wib, nw = (int64_t)gridDim.x * (kThreads / 32);
  const float inv_b = 1.f / (float)a.n_rays;
  const float lam_d = a.lambda_ds / 3.f;                                              // metrics.py:71
  float n_lab = (COUNTED && a.sem_logits) ? (float)ws->n_labelled : 1.f;
  float s_col = 0.f, s_dep = 0.f, s_sem = 0.f, s_app = 0.f, s_lab = 0.f;   // lane 0 carries the warp's partial sums
  for (int64_t r = warp0; r < a.n_rays; r += nw) {
    if (a.rgb && lane < 3) {                                                           // metrics.py:31,38 MSELoss(mean)
      const float d = a.rgb[r * 3 + lane] - a.rgb_target[r * 3 + lane];
      a.g_rgb[r * 3 + lane] = 2.f * d * inv_b / 3.f;
      float sq = d * d;
      sq += __shfl_down_sync(0x7u, sq, 1) + __shfl_down_sync(0x7u, sq, 2);
      if (lane == 0) s_col += sq;
    }
    if (a.depth) {
      const int64_t ts = a.target_stride > 0 ? a.target_stride : 1;
      const float d = a.depth[r], td = a.target_depth[r * ts], tw = a.target_weight[r * ts];
      bool apply;
      float pstd = 0.f, m1 = 0.f;      // predicted STD and sum_i (z_i - d) w_i (GNLL gradient)
      if (a.use_all_depth) {
        apply = true;                                                                  // metrics.py:140,154-156
      } else {
        const bool valid = a.valid_depth ? a.valid_depth[r] > 0 : true;               // :89
        float v = 0.f;
        for (int i = lane; i < a.n_samples; i += 32) {
          const float dz = a.z[r * a.n_samples + i] - d, w = a.weights[r * a.n_samples + i];
          v = fmaf(dz * dz, w, v);
          m1 = fmaf(dz, w, m1);
        }
        pstd = sqrtf(fmaxf(warp_sum(v), 0.f));                                         // :102
        const float tsd = a.target_std[r];
        apply = valid && (fabsf(d - td) > tsd || pstd > tsd);                          // :78-80,115
      }
      const float e = d - td;
      if (a.gnll) {
        // 0.5 (log var + e^2 / var), var = max(pstd, 1e-6) with an identity gradient through the clamp (torch's
        // GaussianNLLLoss clamps a detached copy); d pstd / d S = 1 / (2 pstd)
        m1 = warp_sum(m1);
        const float var = fmaxf(pstd, 1e-6f);
        const float dvar = apply ? lam_d * inv_b * 0.5f * (1.f / var - e * e / (var * var)) : 0.f;   // d loss / d var
        const float ds = dvar * 0.5f / pstd;                                            // d loss / d S
        for (int i = lane; i < a.n_samples; i += 32) {
          const float dz = a.z[r * a.n_samples + i] - d;
          a.g_weights[r * a.n_samples + i] = apply ? ds * dz * dz : 0.f;
        }
        if (lane == 0) {
          a.g_depth[r] = apply ? lam_d * inv_b * e / var - 2.f * ds * m1 : 0.f;
          if (apply) { s_dep += 0.5f * (logf(var) + e * e / var); s_app += 1.f; }
        }
      } else if (lane == 0) {
        a.g_depth[r] = apply ? lam_d * 2.f * tw * e * inv_b : 0.f;                     // mean over applied of (n_applied/B) tw e^2
        if (apply) { s_dep += tw * e * e; s_app += 1.f; }
      }
    }
    if (a.sem_logits) {                                                                // metrics.py:166,171 CE(ignore_index=-100)
      const int C = a.n_sem;
      int64_t lab = a.labels[r];
      if (lab < 0 || lab >= C) lab = -100;          // out-of-range targets are ignored (torch raises on them)
      float lg = lane < C ? a.sem_logits[r * C + lane] : -INFINITY;
      float m = lg;
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
      const float ex = lane < C ? expf(lg - m) : 0.f;
      const float den = warp_sum(ex);
      if (lane < C) {
        const float sm = ex / den;
        const float t = a.lambda_ss * (sm - (lane == lab ? 1.f : 0.f));
        a.g_sem_logits[r * C + lane] = (lab == -100) ? 0.f : (COUNTED ? t / n_lab : t);
      }
      const float picked = __shfl_sync(0xffffffffu, lg, lab == -100 ? 0 : (int)lab);
      if (lane == 0 && lab != -100) { s_sem += (logf(den) + m) - picked; s_lab += 1.f; }
    }
  }
  __shared__ float red[kThreads / 32][5];
  if (lane == 0) { red[wib][0] = s_col; red[wib][1] = s_dep; red[wib][2] = s_sem; red[wib][3] = s_app; red[wib][4] = s_lab; }
  __syncthreads();
  if (threadIdx.x < 5) {
    float t = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) t += red[w][threadIdx.x];
    ws->partial[blockIdx.x][threadIdx.x] = t;
  }
  __threadfence();
  __syncthreads();
  __shared__ bool is_last;
  if (threadIdx.x == 0) is_last = atomicAdd(&ws->ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (is_last) {      // every block's partial sums and gradient rows are visible (fence before its ticket)
    // column k of the partials by warp k: lanes stride over the blocks, then a fixed shuffle tree (deterministic)
    __shared__ float tot[5];
    __threadfence();
    if (wib < 5) {
      float t = 0.f;
      for (unsigned b = lane; b < gridDim.x; b += 32) t += ws->partial[b][wib];
      t = warp_sum(t);
      if (lane == 0) tot[wib] = t;
    }
    __syncthreads();
    if (!COUNTED && a.sem_logits) {
      n_lab = tot[4];      // exact: integers below 2^24
      if (n_lab > 0.f) {
        // one block touches the whole gradient array: 16-byte accesses, four independent ones in flight per thread
        const int64_t n = a.n_rays * a.n_sem;
        float* g = a.g_sem_logits;
        const bool vec = (n & 3) == 0 && (reinterpret_cast<uintptr_t>(g) & 15) == 0;
        if (vec) {
          float4* g4 = reinterpret_cast<float4*>(g);
          const int64_t n4 = n >> 2;
          for (int64_t i0 = 0; i0 < n4; i0 += 4 * kThreads) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int64_t i = i0 + u * kThreads + threadIdx.x;
              if (i < n4) v[u] = g4[i];
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int64_t i = i0 + u * kThreads + threadIdx.x;
              if (i < n4) g4[i] = make_float4(v[u].x / n_lab, v[u].y / n_lab, v[u].z / n_lab, v[u].w / n_lab);
            }
          }
        } else {
          for (int64_t i = threadIdx.x; i < n; i += kThreads) g[i] = g[i] / n_lab;
        }
      }
    }
    if (threadIdx.x < 4) {
      const float t = tot[threadIdx.x];
      float v = t;
      if (threadIdx.x == 0) v = t * inv_b / 3.f;
      if (threadIdx.x == 1) v = lam_d * t * inv_b;
      if (threadIdx.x == 2) v = a.lambda_ss * t / n_lab;
      a.losses[threadIdx.x] = v;
      a.losses[4 + threadIdx.x] = (threadIdx.x == 0 && a.sem_logits) ? n_lab : 0.f;
      // the workspace cleans itself for the next call (no memsets on the stream): every block has read the label
      // count and taken its ticket by now
      if (threadIdx.x == 0) { ws->ticket = 0; ws->n_labelled = 0; }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Solar-correction terms (metrics.py:17-24) and the uncertainty-aware colour loss (metrics.py:10-14).
// Forward kernels: one warp per ray, deterministic scalars (block partials summed in block order by the
// last block).  Backward kernels: the gradients, scaled by the upstream gradients of the two scalars
// (device pointer; NULL = 1), so the autograd wrapper needs no host synchronisation and no torch arithmetic.
// `sun` / `beta` are read in place from the network's output rows (element (r, i) at p[(r n + i) stride]).
// ------------------------------------------------------------------------------------------------
template <int K>
__device__ __forceinline__ bool block_partials(Workspace* ws, const float (&v)[K]) {
  __shared__ float red[kThreads / 32][4];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  if (lane == 0)
    for (int k = 0; k < K; ++k) red[wib][k] = v[k];
  __syncthreads();
  if (threadIdx.x < K) {
    float t = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) t += red[w][threadIdx.x];
    ws->partial[blockIdx.x][threadIdx.x] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(&ws->ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (is_last) __threadfence();
  return is_last;
}
__device__ __forceinline__ float ordered_total(const Workspace* ws, int k) {
  float t = 0.f;
  for (unsigned b = 0; b < gridDim.x; ++b) t += ws->partial[b][k];
  return t;
}

__global__ void __launch_bounds__(kThreads) solar_fwd_kernel(const SpnerfLossSolar a, Workspace* ws) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, n = a.n_samples;
  const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + wib, nw = (int64_t)gridDim.x * (kThreads / 32);
  float s2 = 0.f, s3 = 0.f;
  for (int64_t r = warp0; r < a.n_rays; r += nw) {
    float q = 0.f, ws_ = 0.f;
    for (int i = lane; i < n; i += 32) {
      const float s = a.sun_sc[(r * n + i) * a.sun_stride];
      const float d = a.transparency_sc[r * n + i] - s;
      q = fmaf(d, d, q);
      ws_ = fmaf(a.weights_sc[r * n + i], s, ws_);
    }
    q = warp_sum(q); ws_ = warp_sum(ws_);
    if (lane == 0) { s2 += q; s3 += 1.f - ws_; }
  }
  const float v[2] = {s2, s3};
  if (block_partials<2>(ws, v)) {
    if (threadIdx.x < 2) a.losses[threadIdx.x] = a.lambda_sc / 3.f * ordered_total(ws, threadIdx.x) / (float)a.n_rays;
    if (threadIdx.x == 0) ws->ticket = 0;
  }
}

__global__ void __launch_bounds__(kThreads) solar_bwd_kernel(const SpnerfLossSolar a) {
  const float up2 = a.upstream ? a.upstream[0] : 1.f, up3 = a.upstream ? a.upstream[1] : 1.f;
  const float c = a.lambda_sc / 3.f / (float)a.n_rays;
  const int64_t total = a.n_rays * a.n_samples;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const float s = a.sun_sc[i * a.sun_stride];
    a.g_sun[i] = c * (up2 * 2.f * (s - a.transparency_sc[i]) - up3 * a.weights_sc[i]);
  }
}

__global__ void __launch_bounds__(kThreads) uncertainty_fwd_kernel(const SpnerfLossUncertainty a, Workspace* ws) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, n = a.n_samples;
  const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + wib, nw = (int64_t)gridDim.x * (kThreads / 32);
  float s_col = 0.f, s_log = 0.f;
  for (int64_t r = warp0; r < a.n_rays; r += nw) {
    float b = 0.f;
    for (int i = lane; i < n; i += 32) b = fmaf(a.weights[r * n + i], a.beta[(r * n + i) * a.beta_stride], b);
    b = warp_sum(b) + a.beta_min;
    if (lane == 0) {
      a.beta_ray[r] = b;
      float q = 0.f;
      for (int c = 0; c < 3; ++c) { const float d = a.rgb[r * 3 + c] - a.rgb_target[r * 3 + c]; q = fmaf(d, d, q); }
      s_col += q / (2.f * b * b);
      s_log += logf(b);
    }
  }
  const float v[2] = {s_col, s_log};
  if (block_partials<2>(ws, v) && threadIdx.x == 0) {
    a.losses[0] = ordered_total(ws, 0) / (3.f * (float)a.n_rays);
    a.losses[1] = (3.f + ordered_total(ws, 1) / (float)a.n_rays) * 0.5f;
    ws->ticket = 0;
  }
}

__global__ void __launch_bounds__(kThreads) uncertainty_bwd_kernel(const SpnerfLossUncertainty a) {
  const float upc = a.upstream ? a.upstream[0] : 1.f, upl = a.upstream ? a.upstream[1] : 1.f;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, n = a.n_samples;
  const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + wib, nw = (int64_t)gridDim.x * (kThreads / 32);
  const float inv_b = 1.f / (float)a.n_rays;
  for (int64_t r = warp0; r < a.n_rays; r += nw) {
    const float b = a.beta_ray[r];
    float q = 0.f;
    for (int c = 0; c < 3; ++c) {
      const float d = a.rgb[r * 3 + c] - a.rgb_target[r * 3 + c];
      q = fmaf(d, d, q);
      if (lane == c) a.g_rgb[r * 3 + c] = upc * d / (b * b) * inv_b / 3.f;
    }
    const float gb = -upc * q / (b * b * b) * inv_b / 3.f + upl * 0.5f * inv_b / b;     // d loss / d beta_ray
    for (int i = lane; i < n; i += 32) {
      a.g_weights[r * n + i] = gb * a.beta[(r * n + i) * a.beta_stride];
      a.g_beta[r * n + i] = gb * a.weights[r * n + i];
    }
  }
}

}  // namespace

static unsigned ray_blocks(int64_t n_rays) {
  const int64_t need = (n_rays + kThreads / 32 - 1) / (kThreads / 32);
  return (unsigned)(need < kMaxBlocks ? need : kMaxBlocks);
}

extern "C" int spnerf_loss_solar(const SpnerfLossSolar* a, int backward, void* stream_) {
  if (!a || a->n_rays <= 0 || a->n_samples < 1 || !a->transparency_sc || !a->weights_sc || !a->sun_sc || a->sun_stride < 1)
    return SPNERF_ERR_BAD_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!backward) {
    if (!a->losses || !a->workspace) return SPNERF_ERR_BAD_ARG;
    Workspace* ws = static_cast<Workspace*>(a->workspace);
    solar_fwd_kernel<<<ray_blocks(a->n_rays), kThreads, 0, stream>>>(*a, ws);
  } else {
    if (!a->g_sun) return SPNERF_ERR_BAD_ARG;
    const int64_t nb = (a->n_rays * a->n_samples + kThreads - 1) / kThreads;
    solar_bwd_kernel<<<(unsigned)(nb < 148 * 16 ? nb : 148 * 16), kThreads, 0, stream>>>(*a);
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}

extern "C" int spnerf_loss_uncertainty(const SpnerfLossUncertainty* a, int backward, void* stream_) {
  if (!a || a->n_rays <= 0 || a->n_samples < 1 || !a->rgb || !a->rgb_target || !a->weights || !a->beta || a->beta_stride < 1 ||
      !a->beta_ray)
    return SPNERF_ERR_BAD_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (!backward) {
    if (!a->losses || !a->workspace) return SPNERF_ERR_BAD_ARG;
    Workspace* ws = static_cast<Workspace*>(a->workspace);
    uncertainty_fwd_kernel<<<ray_blocks(a->n_rays), kThreads, 0, stream>>>(*a, ws);
  } else {
    if (!a->g_rgb || !a->g_weights || !a->g_beta) return SPNERF_ERR_BAD_ARG;
    uncertainty_bwd_kernel<<<ray_blocks(a->n_rays), kThreads, 0, stream>>>(*a);
  }
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}

extern "C" int64_t spnerf_losses_workspace_bytes(void) { return (int64_t)sizeof(Workspace); }

extern "C" int spnerf_losses(const SpnerfLosses* a, void* stream_) {
  if (!a || !a->losses || !a->workspace || a->n_rays <= 0) return SPNERF_ERR_BAD_ARG;
  if (a->rgb && (!a->rgb_target || !a->g_rgb)) return SPNERF_ERR_BAD_ARG;
  if (a->depth && (!a->target_depth || !a->target_weight || !a->g_depth)) return SPNERF_ERR_BAD_ARG;
  if (a->depth && !a->use_all_depth && (!a->z || !a->weights || !a->target_std)) return SPNERF_ERR_BAD_ARG;
  if (a->depth && a->gnll && (a->use_all_depth || !a->g_weights)) return SPNERF_ERR_BAD_ARG;
  if (a->sem_logits && (!a->labels || !a->g_sem_logits || a->n_sem < 1 || a->n_sem > 32)) return SPNERF_ERR_BAD_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  Workspace* ws = static_cast<Workspace*>(a->workspace);
  const bool counted = a->sem_logits && a->n_rays * (int64_t)a->n_sem > kSelfCountMax;
  if (counted) {
    const int64_t nb = (a->n_rays + kThreads - 1) / kThreads;
    count_labels_kernel<<<(unsigned)(nb < 592 ? nb : 592), kThreads, 0, stream>>>(a->labels, a->n_rays, a->n_sem, ws);
  }
  const int64_t need = (a->n_rays + kThreads / 32 - 1) / (kThreads / 32);
  const unsigned blocks = (unsigned)(need < 592 ? need : 592);      // 148 SMs x 4: a short serial tail in the last block
  if (counted) losses_kernel<true><<<blocks, kThreads, 0, stream>>>(*a, ws);
  else losses_kernel<false><<<blocks, kThreads, 0, stream>>>(*a, ws);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}
