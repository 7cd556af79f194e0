// Probe 2: the wgrad ring in miniature.  Both CTAs of a pair issue 3 tensor copies (32 KB each) per iteration into
// a 2-stage ring, all completing on the LEADER's stage barrier (expect_tx by rank 0 for both CTAs' bytes + a second
// plain arrive); rank 0 releases the stage to both CTAs with (remote) arrives.  300 iterations.
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
struct P { int* result; int iters; };
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t r) { uint32_t o; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r)); return o; }
__device__ __forceinline__ bool try_wait(uint32_t bar, uint32_t par) {
  uint32_t ok;
  asm volatile("{\n.reg .pred q;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 q, [%1], %2;\nselp.u32 %0, 1, 0, q;\n}" : "=r"(ok) : "r"(bar), "r"(par) : "memory");
  return ok;
}
__device__ __forceinline__ bool wait(uint32_t bar, uint32_t par, int* res, int code) {
  long long t0 = clock64();
  while (!try_wait(bar, par)) if (clock64() - t0 > 200000000ll) { atomicCAS(res, 0, code); return false; }
  return true;
}
constexpr int kStage = 96 * 1024;
__global__ void __cluster_dims__(2, 1, 1) probe(const __grid_constant__ P p, const __grid_constant__ CUtensorMap tm) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * kStage);      // full[2], empty[2]
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bars[i])), "r"(rank == 0 ? 2 : 1));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[2 + i])));
    }
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  const int warp = threadIdx.x >> 5;
  if (warp == 0 && (threadIdx.x & 31) == 0) {            // producer of both CTAs
    uint32_t stage = 0, phase = 0;
    for (int it = 0; it < p.iters; ++it) {
      if (!wait(s32(&bars[2 + stage]), phase ^ 1, p.result, 10 + rank)) break;
      const uint32_t bar0 = mapa(s32(&bars[stage]), 0);
      if (rank == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bars[stage])), "r"(2 * 3 * 32768) : "memory");
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&bars[stage])) : "memory");
      }
      for (int c = 0; c < 3; ++c)
        asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(s32(smem + stage * kStage + c * 32768)), "l"(&tm), "r"(0), "r"((int)((it * 6 + rank * 3 + c) % 64) * 16), "r"(bar0) : "memory");
      if (++stage == 2) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && (threadIdx.x & 31) == 0 && rank == 0) {   // "issuer": consume and release both CTAs' stages
    uint32_t stage = 0, phase = 0;
    for (int it = 0; it < p.iters; ++it) {
      if (!wait(s32(&bars[stage]), phase, p.result, 20)) break;
      const uint64_t* d = reinterpret_cast<const uint64_t*>(smem + stage * kStage);
      for (int c = 0; c < 3; ++c)
        if (d[c * 4096 + 5] != (uint64_t)(((it * 6 + c) % 64) * 16)) atomicAdd(&p.result[1], 1);
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&bars[2 + stage])) : "memory");
      asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(mapa(s32(&bars[2 + stage]), 1)) : "memory");
      if (++stage == 2) { stage = 0; phase ^= 1; }
    }
    p.result[2] = 1;
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
int main() {
  const int rows = 64 * 16;
  std::vector<uint64_t> h((size_t)rows * 256);
  for (int r = 0; r < rows; ++r) for (int i = 0; i < 256; ++i) h[(size_t)r * 256 + i] = r;
  uint64_t* d; int* res;
  cudaMalloc(&d, h.size() * 8); cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  cudaMalloc(&res, 16); cudaMemset(res, 0, 16);
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
  P p; p.result = res; p.iters = 300;
  alignas(64) CUtensorMap tm;
  const cuuint64_t dims[2] = {256, (cuuint64_t)rows}; const cuuint64_t strides[1] = {2048};
  const cuuint32_t box[2] = {256, 16}; const cuuint32_t es[2] = {1, 1};
  CUresult cr = ((EncodeFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode -> %d\n", (int)cr);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kStage + 64);
  probe<<<148, 64, 2 * kStage + 64>>>(p, tm);
  cudaError_t e = cudaDeviceSynchronize();
  int r[4] = {0, 0, 0, 0}; cudaMemcpy(r, res, 16, cudaMemcpyDeviceToHost);
  printf("sync: %s; timeout code %d, mismatches %d, finished %d\n", cudaGetErrorString(e), r[0], r[1], r[2]);
  return 0;
}
