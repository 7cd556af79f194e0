// Micro-benchmark of the hand-off latencies that bound the weight ring of the fused MLP kernels
// (single CTA pair, cycles from clock64 on the issuing SM).  Not part of the library.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../sp-nerf_b200/csrc/sm100.cuh"
using namespace sm100;

struct Out { long long v[64]; };

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) hop_kernel(const uint8_t* src, Out* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar[8];
  __shared__ uint32_t tmem_slot;
  const uint32_t rank = cluster_ctarank();
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) mbar_init(&bar[i], 1);
    fence_mbar_init();
  }
  if (warp == 0) { tmem_alloc2(&tmem_slot, 512); tmem_relinquish2(); }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  long long r[64];
  for (int i = 0; i < 64; ++i) r[i] = 0;
  if (warp == 0) {
    // 0: clock overhead
    long long t0 = clock64(); long long t1 = clock64(); r[0] = t1 - t0;
    // 1: arrive + try_wait (same thread, all lanes)
    uint32_t par = 0;
    t0 = clock64();
    if (elect_one()) mbar_arrive(&bar[0]);
    __syncwarp();
    mbar_wait(&bar[0], par); par ^= 1;
    r[1] = clock64() - t0;
    // 2: try_wait on an already completed phase
    if (elect_one()) mbar_arrive(&bar[0]);
    __syncwarp();
    for (int i = 0; i < 2000; ++i) __nanosleep(1);
    t0 = clock64();
    mbar_wait(&bar[0], par); par ^= 1;
    r[2] = clock64() - t0;
    // 3/4/5: bulk copy 16 KB / 32 KB / 1 KB global(L2-warm) -> shared, issue to observed
    for (int rep = 0; rep < 3; ++rep) {
      const uint32_t bytes = rep == 0 ? 16384u : (rep == 1 ? 32768u : 1024u);
      uint32_t p1 = 0;
      for (int w = 0; w < 3; ++w) {       // warm L2, measure the last
        t0 = clock64();
        if (elect_one()) { mbar_expect_tx(&bar[1], bytes); bulk_g2s(smem, src, bytes, &bar[1]); }
        __syncwarp();
        mbar_wait(&bar[1], p1); p1 ^= 1;
        r[3 + rep] = clock64() - t0;
      }
      if (p1) { if (elect_one()) mbar_arrive(&bar[1]); __syncwarp(); mbar_wait(&bar[1], p1); p1 ^= 1; }
    }
    // 6: commit with nothing pending (cta_group::1)
    {
      uint32_t p2 = 0;
      t0 = clock64();
      if (elect_one()) umma_commit(&bar[2]);
      __syncwarp();
      mbar_wait(&bar[2], p2); p2 ^= 1;
      r[6] = clock64() - t0;
      // 7: 4 MMAs (M=128,N=256,K=16) + commit -> observed (cta_group::1)
      constexpr uint64_t tmpl = make_smem_desc_template(16, 1024, kSwizzle128B);
      const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 16384;
      for (int w = 0; w < 2; ++w) {
        t0 = clock64();
        if (elect_one()) {
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem, smem_desc(tmpl, a0 + k * 32), smem_desc(tmpl, b0 + k * 32), make_idesc_f16(128, 256, 0, 0), 1u);
          umma_commit(&bar[2]);
        }
        __syncwarp();
        r[8] = clock64() - t0;     // issue cost
        mbar_wait(&bar[2], p2); p2 ^= 1;
        r[7] = clock64() - t0;
      }
      // 9: 16 MMAs + commit
      t0 = clock64();
      if (elect_one()) {
        for (int k = 0; k < 16; ++k)
          umma_f16(tmem, smem_desc(tmpl, a0 + (k & 3) * 32), smem_desc(tmpl, b0 + (k & 3) * 32), make_idesc_f16(128, 256, 0, 0), 1u);
        umma_commit(&bar[2]);
      }
      __syncwarp();
      r[10] = clock64() - t0;
      mbar_wait(&bar[2], p2); p2 ^= 1;
      r[9] = clock64() - t0;
    }
  }
  cluster_sync_all();
  // pair hops: rank 0 warp 0 <-> rank 1 warp 0
  if (warp == 0) {
    uint32_t p3 = 0;
    const uint32_t remote = mapa_shared(smem_u32(&bar[3]), rank ^ 1);
    // ping-pong 8 times: rank 0 arrives on peer's bar[3]; peer waits then arrives back
    long long t0 = clock64();
    for (int i = 0; i < 8; ++i) {
      if (rank == 0) {
        if (elect_one()) mbar_arrive_remote(remote);
        __syncwarp();
        mbar_wait_cluster(&bar[3], p3); p3 ^= 1;
      } else {
        mbar_wait_cluster(&bar[3], p3); p3 ^= 1;
        if (elect_one()) mbar_arrive_remote(remote);
        __syncwarp();
      }
    }
    r[11] = (clock64() - t0) / 8;      // round trip (2 hops)
    // multicast commit with nothing pending: leader commits to both, both wait
    uint32_t p4 = 0;
    cluster_sync_all();
    t0 = clock64();
    if (rank == 0 && elect_one()) umma2_commit(&bar[4], 3);
    __syncwarp();
    mbar_wait(&bar[4], p4); p4 ^= 1;
    r[12] = clock64() - t0;
    // pair MMA x4 + multicast commit
    cluster_sync_all();
    constexpr uint64_t tmpl = make_smem_desc_template(16, 1024, kSwizzle128B);
    const uint32_t a0 = smem_u32(smem), b0 = smem_u32(smem) + 16384;
    t0 = clock64();
    if (rank == 0 && elect_one()) {
      for (int k = 0; k < 4; ++k)
        umma2_f16(tmem, smem_desc(tmpl, a0 + k * 32), smem_desc(tmpl, b0 + k * 32), make_idesc_f16(256, 256, 0, 0), 1u);
      umma2_commit(&bar[4], 3);
    }
    __syncwarp();
    r[14] = clock64() - t0;
    mbar_wait(&bar[4], p4); p4 ^= 1;
    r[13] = clock64() - t0;
  } else {
    cluster_sync_all();
    cluster_sync_all();
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc2(tmem, 512);
  if (threadIdx.x == 0)
    for (int i = 0; i < 64; ++i) out[rank].v[i] = r[i];
}

int main() {
  uint8_t* src; Out* out;
  cudaMalloc(&src, 1 << 20); cudaMemset(src, 0, 1 << 20);
  cudaMalloc(&out, 2 * sizeof(Out));
  cudaFuncSetAttribute(hop_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  for (int it = 0; it < 2; ++it) hop_kernel<<<2, 128, 65536>>>(src, out);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  Out h[2];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[] = {"clock64 pair", "arrive+wait same thread", "wait on completed phase", "bulk 16KB L2->smem", "bulk 32KB",
                         "bulk 1KB", "commit (nothing pending)", "4 MMA + commit -> seen", "  issue part", "16 MMA + commit -> seen",
                         "  issue part", "remote arrive ping-pong round trip", "multicast commit (nothing pending)",
                         "4 pair-MMA + multicast commit -> seen", "  issue part"};
  for (int i = 0; i < 15; ++i) printf("%-40s rank0 %6lld  rank1 %6lld\n", names[i], h[0].v[i], h[1].v[i]);
  return 0;
}
