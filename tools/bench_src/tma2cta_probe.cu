// Probe: cp.async.bulk.tensor.2d.cta_group::2 from both CTAs of a pair completing on the LEADER's mbarrier.
// Tensor map: rows of 2 KB (256 x uint64), box = 16 rows (32 KB).  Each CTA loads its own box; rank 0 waits for
// 64 KB of transaction bytes, tells rank 1, and both check what landed.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o tma2cta_probe tma2cta_probe.cu
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

struct P { alignas(64) CUtensorMap tm; int* result; };

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __cluster_dims__(2, 1, 1) probe(const __grid_constant__ P p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536);       // [0] full (rank 0), [1] go (both)
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[0])));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[1])));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x == 0) {
    uint32_t bar0;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(bar0) : "r"(s32(&bar[0])), "r"(0));
    if (rank == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar[0])), "r"(65536) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(s32(smem)), "l"(&p.tm), "r"(0), "r"((int)(16 * rank + 32 * (blockIdx.x >> 1))), "r"(bar0) : "memory");
    if (rank == 0) {
      uint32_t ok = 0;
      long long t0 = clock64();
      while (!ok && clock64() - t0 < 200000000ll)
        asm volatile("{\n.reg .pred q;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 q, [%1], %2;\nselp.u32 %0, 1, 0, q;\n}" : "=r"(ok) : "r"(s32(&bar[0])), "r"(0) : "memory");
      p.result[0] = ok ? 1 : -1;
      uint32_t go1;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(go1) : "r"(s32(&bar[1])), "r"(1));
      asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(go1) : "memory");
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&bar[1])) : "memory");
    }
  }
  if (threadIdx.x == 0) {
    uint32_t ok = 0;
    long long t0 = clock64();
    while (!ok && clock64() - t0 < 400000000ll)
      asm volatile("{\n.reg .pred q;\nmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 q, [%1], %2;\nselp.u32 %0, 1, 0, q;\n}" : "=r"(ok) : "r"(s32(&bar[1])), "r"(0) : "memory");
  }
  __syncthreads();
  // row r of this CTA's box holds the value (16 * rank + 32 * pair + r) in every uint64
  const uint64_t* d = reinterpret_cast<const uint64_t*>(smem);
  int bad = 0;
  for (int i = threadIdx.x; i < 4096; i += blockDim.x)
    if (d[i] != (uint64_t)(16 * rank + 32 * (blockIdx.x >> 1) + i / 256)) ++bad;
  if (bad) atomicAdd(&p.result[1 + rank], bad);
  asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

int main() {
  const int rows = 64 * 32;
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
  P p; p.result = res;
  const cuuint64_t dims[2] = {256, (cuuint64_t)rows}; const cuuint64_t strides[1] = {2048};
  const cuuint32_t box[2] = {256, 16}; const cuuint32_t es[2] = {1, 1};
  CUresult cr = ((EncodeFn)fn)(&p.tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode -> %d\n", (int)cr);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 + 64);
  probe<<<8, 128, 65536 + 64>>>(p);
  cudaError_t e = cudaDeviceSynchronize();
  int r[4]; cudaMemcpy(r, res, 16, cudaMemcpyDeviceToHost);
  printf("sync: %s; leader wait %d, mismatches rank0 %d rank1 %d\n", cudaGetErrorString(e), r[0], r[1], r[2]);
  return 0;
}
