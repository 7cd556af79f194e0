// Operand feed of the weight-gradient kernel in isolation: per 128-point K tile a CTA receives 32 KB of A' (its own
// rows) and 64 KB of B' (shared with the CTA of the same rank in the pair that owns the other 256 rows of the layer)
// into a 2-stage ring, and a consumer holds each stage for `delay` cycles (2048 = the MMAs of the tile).
//   mode 0: clusters of 2, every CTA fetches its 96 KB itself (what mlp_wgrad.cu does), 148 CTAs
//   mode 1: clusters of 4, each of the two CTAs that share a B' half fetches 32 KB of it and multicasts to both
//   mode 2: clusters of 4, every CTA fetches its 96 KB itself (lock-step through the stage barriers only)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mcast_feed mcast_feed.cu
// run:   ./mcast_feed <mode> <delay cycles> [tiles per CTA] [rounds]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_arrive_remote(uint32_t a) { asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(a) : "memory"); }
__device__ __forceinline__ uint32_t mapa(uint32_t a, uint32_t rank) { uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank)); return r; }
__device__ __forceinline__ bool try_wait(uint64_t* b, uint32_t par) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(par) : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool try_wait_cluster(uint64_t* b, uint32_t par) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(par) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s_mc(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

constexpr uint32_t kA = 32 * 1024, kB = 64 * 1024, kStage = kA + kB, kPiece = 16 * 1024;

// a_base: private stream of this CTA (tiles * 32 KB); b_base: stream of this CTA's B' half (tiles * 64 KB), the same
// pointer for the two CTAs that share it
template <int MODE>
__global__ void __launch_bounds__(64, 1) feed_kernel(const uint8_t* __restrict__ a_all, const uint8_t* __restrict__ b_all,
                                                     int tiles, int delay, long long* __restrict__ stats, int rounds) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + 2 * kStage);
  uint64_t* empty = full + 2;
  const uint32_t rank = cluster_rank();
  constexpr int CS = MODE == 0 ? 2 : 4;
  const int cluster = blockIdx.x / CS;
  // the CTA of rank r shares B' with rank r^2 (mode 1, 2) or with the same rank of the neighbouring cluster (mode 0)
  const int group = MODE == 0 ? cluster / 2 : cluster;                 // which B' stream pair
  const int half = MODE == 0 ? (int)rank : (int)(rank & 1);            // which 64 KB half of B' this CTA consumes
  const uint8_t* a_base = a_all + (size_t)blockIdx.x * tiles * kA;
  const uint8_t* b_base = b_all + ((size_t)group * 2 + half) * tiles * kB;
  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], MODE == 1 ? 2 : 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    for (int t = 0; t < tiles * rounds; ++t) {
      const int s = t & 1;
      const int ta = t % tiles;
      if (t >= 2) { const uint32_t par = ((t >> 1) - 1) & 1; while (!(MODE == 1 ? try_wait_cluster(&empty[s], par) : try_wait(&empty[s], par))) {} }
      uint8_t* st = smem + s * kStage;
      mbar_expect_tx(&full[s], kStage);
      for (uint32_t o = 0; o < kA; o += kPiece) bulk_g2s(st + o, a_base + (size_t)ta * kA + o, kPiece, &full[s]);
      if (MODE == 1) {
        const uint32_t q = rank >> 1;      // which 32 KB of the shared half this CTA fetches for both
        const uint16_t mask = (uint16_t)((1u << (rank & 1)) | (1u << ((rank & 1) + 2)));
        for (uint32_t o = 0; o < kB / 2; o += kPiece)
          bulk_g2s_mc(st + kA + q * (kB / 2) + o, b_base + (size_t)ta * kB + q * (kB / 2) + o, kPiece, &full[s], mask);
      } else {
        for (uint32_t o = 0; o < kB; o += kPiece) bulk_g2s(st + kA + o, b_base + (size_t)ta * kB + o, kPiece, &full[s]);
      }
    }
  } else if (warp == 1 && lane == 0) {
    const long long t0 = clock64();
    long long waited = 0;
    uint32_t peer_empty0 = 0, peer_empty1 = 0;
    if (MODE == 1) { peer_empty0 = mapa(smem_u32(&empty[0]), rank ^ 2); peer_empty1 = mapa(smem_u32(&empty[1]), rank ^ 2); }
    for (int t = 0; t < tiles * rounds; ++t) {
      const int s = t & 1;
      const long long w0 = clock64();
      while (!try_wait(&full[s], (t >> 1) & 1)) {}
      const long long w1 = clock64();
      waited += w1 - w0;
      while (clock64() - w1 < delay) {}
      mbar_arrive(&empty[s]);
      if (MODE == 1) mbar_arrive_remote(s ? peer_empty1 : peer_empty0);
    }
    const long long t1 = clock64();
    stats[blockIdx.x * 2] = t1 - t0;
    stats[blockIdx.x * 2 + 1] = waited;
  }
  cluster_sync();
}

template <int MODE>
static void run(int delay, int tiles, int rounds) {
  constexpr int CS = MODE == 0 ? 2 : 4;
  const int smem = 2 * kStage + 64;
  CK(cudaFuncSetAttribute(feed_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(64, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = CS; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  cfg.gridDim = dim3(148 / CS * CS, 1, 1);
  int nclusters = 0;
  CK(cudaOccupancyMaxActiveClusters(&nclusters, feed_kernel<MODE>, &cfg));
  const int ctas = nclusters * CS;
  cfg.gridDim = dim3(ctas, 1, 1);
  const int groups = MODE == 0 ? (nclusters + 1) / 2 : nclusters;
  uint8_t *a, *b; long long* stats;
  CK(cudaMalloc(&a, (size_t)ctas * tiles * kA));
  CK(cudaMalloc(&b, (size_t)groups * 2 * tiles * kB));
  CK(cudaMalloc(&stats, ctas * 2 * sizeof(long long)));
  CK(cudaMemset(a, 1, (size_t)ctas * tiles * kA));
  CK(cudaMemset(b, 2, (size_t)groups * 2 * tiles * kB));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    // cold L2 for the streams: write a 256 MB scratch buffer in between
    static uint8_t* scratch = nullptr;
    if (!scratch) CK(cudaMalloc(&scratch, 256u << 20));
    CK(cudaMemset(scratch, rep, 256u << 20));
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, feed_kernel<MODE>, (const uint8_t*)a, (const uint8_t*)b, tiles, delay, stats, rounds));
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best) best = ms;
  }
  long long* h = (long long*)malloc(ctas * 2 * sizeof(long long));
  CK(cudaMemcpy(h, stats, ctas * 2 * sizeof(long long), cudaMemcpyDeviceToHost));
  double cyc = 0, wait = 0, cmax = 0;
  for (int i = 0; i < ctas; ++i) { cyc += h[2 * i]; wait += h[2 * i + 1]; if (h[2 * i] > cmax) cmax = h[2 * i]; }
  const double delivered = (double)ctas * tiles * rounds * kStage;
  tiles *= rounds;
  printf("mode %d delay %d: %d clusters of %d = %d CTAs, %d tiles each: %.3f ms, %.0f GB/s into shared memory, "
         "cycles per tile avg %.0f (max CTA %.0f), waiting for operands %.0f per tile; tiles per ms (all CTAs) %.0f\n",
         MODE, delay, nclusters, CS, ctas, tiles, best, delivered / best / 1e6, cyc / ctas / tiles, cmax / tiles,
         wait / ctas / tiles, (double)ctas * tiles / best);
  cudaFree(a); cudaFree(b); cudaFree(stats); free(h);
}

int main(int argc, char** argv) {
  // rounds > 1 walks the same `tiles` again: with a working set below the 126 MB L2 (tiles <= 8) the feed comes from L2
  const int mode = argc > 1 ? atoi(argv[1]) : 0, delay = argc > 2 ? atoi(argv[2]) : 0, tiles = argc > 3 ? atoi(argv[3]) : 160;
  const int rounds = argc > 4 ? atoi(argv[4]) : 1;
  if (mode == 0) run<0>(delay, tiles, rounds);
  else if (mode == 1) run<1>(delay, tiles, rounds);
  else run<2>(delay, tiles, rounds);
  return 0;
}
