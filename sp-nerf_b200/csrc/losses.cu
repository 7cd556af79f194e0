// Fused loss reductions + gradients: colour MSE, depth supervision, semantic cross-entropy.
//
// Replaces modules/metrics.py:27-45 (SNerfLoss colour term), :68-159 (DepthLoss, MSE variants) and
// :162-183 (SemanticLoss): one warp per ray, block partial sums, the last block to finish adds
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
  float partial[kMaxBlocks][4];
};

__global__ void count_labels_kernel(const int64_t* __restrict__ labels, int64_t n, Workspace* ws) {
  unsigned int c = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    c += labels[i] != -100;
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) c += __shfl_xor_sync(0xffffffffu, c, s);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(&ws->n_labelled, c);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
  return v;
}

__global__ void __launch_bounds__(kThreads) losses_kernel(const SpnerfLosses a, Workspace* ws) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp0 = (int64_t)blockIdx.x * (kThreads / 32) + wib, nw = (int64_t)gridDim.x * (kThreads / 32);
  const float inv_b = 1.f / (float)a.n_rays;
  const float lam_d = a.lambda_ds / 3.f;                                              // metrics.py:71
  const float n_lab = a.sem_logits ? (float)ws->n_labelled : 1.f;
  float s_col = 0.f, s_dep = 0.f, s_sem = 0.f, s_app = 0.f;   // lane 0 carries the warp's partial sums
  for (int64_t r = warp0; r < a.n_rays; r += nw) {
    if (a.rgb && lane < 3) {                                                           // metrics.py:31,38 MSELoss(mean)
      const float d = a.rgb[r * 3 + lane] - a.rgb_target[r * 3 + lane];
      a.g_rgb[r * 3 + lane] = 2.f * d * inv_b / 3.f;
      float sq = d * d;
      sq += __shfl_down_sync(0x7u, sq, 1) + __shfl_down_sync(0x7u, sq, 2);
      if (lane == 0) s_col += sq;
    }
    if (a.depth) {
      const float d = a.depth[r], td = a.target_depth[r], tw = a.target_weight[r];
      bool apply;
      if (a.use_all_depth) {
        apply = true;                                                                  // metrics.py:140,154-156
      } else {
        const bool valid = a.valid_depth ? a.valid_depth[r] > 0 : true;               // :89
        float v = 0.f;
        for (int i = lane; i < a.n_samples; i += 32) {
          const float dz = a.z[r * a.n_samples + i] - d;
          v = fmaf(dz * dz, a.weights[r * a.n_samples + i], v);
        }
        const float pstd = sqrtf(warp_sum(v));                                         // :102
        const float tsd = a.target_std[r];
        apply = valid && (fabsf(d - td) > tsd || pstd > tsd);                          // :78-80,115
      }
      if (lane == 0) {
        const float e = d - td;
        a.g_depth[r] = apply ? lam_d * 2.f * tw * e * inv_b : 0.f;                     // mean over applied of (n_applied/B) tw e^2
        if (apply) { s_dep += tw * e * e; s_app += 1.f; }
      }
    }
    if (a.sem_logits) {                                                                // metrics.py:166,171 CE(ignore_index=-100)
      const int64_t lab = a.labels[r];
      const int C = a.n_sem;
      float lg = lane < C ? a.sem_logits[r * C + lane] : -INFINITY;
      float m = lg;
#pragma unroll
      for (int s = 16; s > 0; s >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
      const float ex = lane < C ? expf(lg - m) : 0.f;
      const float den = warp_sum(ex);
      if (lane < C) {
        const float sm = ex / den;
        a.g_sem_logits[r * C + lane] = (lab == -100) ? 0.f : a.lambda_ss * (sm - (lane == lab ? 1.f : 0.f)) / n_lab;
      }
      const float picked = __shfl_sync(0xffffffffu, lg, lab == -100 ? 0 : (int)lab);
      if (lane == 0 && lab != -100) s_sem += (logf(den) + m) - picked;
    }
  }
  __shared__ float red[kThreads / 32][4];
  if (lane == 0) { red[wib][0] = s_col; red[wib][1] = s_dep; red[wib][2] = s_sem; red[wib][3] = s_app; }
  __syncthreads();
  if (threadIdx.x < 4) {
    float t = 0.f;
    for (int w = 0; w < kThreads / 32; ++w) t += red[w][threadIdx.x];
    ws->partial[blockIdx.x][threadIdx.x] = t;
  }
  __threadfence();
  __syncthreads();
  __shared__ bool is_last;
  if (threadIdx.x == 0) is_last = atomicAdd(&ws->ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (is_last && threadIdx.x < 4) {
    __threadfence();
    float t = 0.f;
    for (unsigned b = 0; b < gridDim.x; ++b) t += ws->partial[b][threadIdx.x];
    float v = t;
    if (threadIdx.x == 0) v = t * inv_b / 3.f;
    if (threadIdx.x == 1) v = lam_d * t * inv_b;
    if (threadIdx.x == 2) v = a.lambda_ss * t / n_lab;
    a.losses[threadIdx.x] = v;
    if (threadIdx.x == 0) a.losses[4] = a.sem_logits ? n_lab : 0.f;
  }
}

}  // namespace

extern "C" int64_t spnerf_losses_workspace_bytes(void) { return (int64_t)sizeof(Workspace); }

extern "C" int spnerf_losses(const SpnerfLosses* a, void* stream_) {
  if (!a || !a->losses || !a->workspace || a->n_rays <= 0) return SPNERF_ERR_BAD_ARG;
  if (a->rgb && (!a->rgb_target || !a->g_rgb)) return SPNERF_ERR_BAD_ARG;
  if (a->depth && (!a->target_depth || !a->target_weight || !a->g_depth)) return SPNERF_ERR_BAD_ARG;
  if (a->depth && !a->use_all_depth && (!a->z || !a->weights || !a->target_std)) return SPNERF_ERR_BAD_ARG;
  if (a->sem_logits && (!a->labels || !a->g_sem_logits || a->n_sem < 1 || a->n_sem > 32)) return SPNERF_ERR_BAD_ARG;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  Workspace* ws = static_cast<Workspace*>(a->workspace);
  cudaMemsetAsync(ws, 0, 16, stream);
  cudaMemsetAsync(a->losses, 0, 8 * sizeof(float), stream);
  if (a->sem_logits) {
    const int64_t nb = (a->n_rays + kThreads - 1) / kThreads;
    count_labels_kernel<<<(unsigned)(nb < 592 ? nb : 592), kThreads, 0, stream>>>(a->labels, a->n_rays, ws);
  }
  const int64_t need = (a->n_rays + kThreads / 32 - 1) / (kThreads / 32);
  const unsigned blocks = (unsigned)(need < kMaxBlocks ? need : kMaxBlocks);
  losses_kernel<<<blocks, kThreads, 0, stream>>>(*a, ws);
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? 0 : -(int)e;
}
